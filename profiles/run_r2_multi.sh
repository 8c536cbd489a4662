#!/bin/bash
# usage: bash profiles/run_r2_multi.sh N   (inside `gpurun --gpus N`): multi-device tests + the scaling bench at N ranks
N=$1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2m_topo_$N.txt 2>&1
python -m pytest tests -m gpu -q --timeout 900 -k "multi_device or number_of_engines or phred_entry_on_a_multi or two_gpu" > gpurun_out/r2m_tests_$N.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2m_tests_$N.log
tail -3 gpurun_out/r2m_tests_$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --methods es --single-process > gpurun_out/r2m_bench_$N.json 2> gpurun_out/r2m_bench_$N.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2m_bench_$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2m_bench_$N.json").read().strip().splitlines()[-1])
print("N=$N value %.4g e2e %.4g (ceil %.1f GB/s, frac %.3f) e2e_phred %.4g (frac %.3f) e2e_fp64 %.4g single-process %s" % (d["value"], d["e2e"]["value"], d["e2e"]["host_copy_ceiling_gbs"], d["e2e"]["frac_of_host_copy_ceiling"], d["e2e_phred"]["value"], d["e2e_phred"]["frac_of_host_copy_ceiling"], d["e2e_fp64_full"]["value"], d.get("e2e_single_process")))
PY
