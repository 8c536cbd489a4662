# 2 GPUs: the command line with -device 0,1 against -device 0 (same bytes), then bench.py on two ranks.
python -m pytest tests -m gpu -q --timeout 900 -k "cuda_build" 2>&1 | tail -2
python - <<'PY'
import os, subprocess, sys, time
sys.path.insert(0, os.getcwd())
from famseq_b200 import synth
ped = synth.ped40(); ped.write("/tmp/p40.ped")
pl, fl = synth.synth_pl(ped, 40000, seed=11)
synth.write_vcf("/tmp/p40.vcf", ped, pl, fl)
outs = {}
for dev in ("0", "0,1", "all"):
    t0 = time.perf_counter()
    r = subprocess.run(["famseq_b200/bin/FamSeq", "vcf", "-vcfFile", "/tmp/p40.vcf", "-pedFile", "/tmp/p40.ped", "-method", "3", "-numBurnIn", "200",
                        "-numRep", "2000", "-device", dev, "-output", f"/tmp/o_{dev}.vcf"], capture_output=True, text=True, env=dict(os.environ, FAMSEQ_STATS="1"))
    outs[dev] = open(f"/tmp/o_{dev}.vcf").read()
    print("device", dev, "rc", r.returncode, "wall", round(time.perf_counter() - t0, 2), r.stderr.strip().splitlines()[-1][:300])
print("identical:", outs["0"] == outs["0,1"] == outs["all"], len(outs["0"]))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_2gpu_m.json 2> gpurun_out/bench_2gpu_m.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_2gpu_m.json").read().strip().splitlines()[-1])
print("ES", d["value"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], "n_gpus", d["n_gpus"])
for k,m in d["methods"].items(): print(k, m["value"], m["ms_per_step"], m.get("kernel"))
PY
