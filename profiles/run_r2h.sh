#!/bin/bash
# round 2 record set: full GPU suite, smoke, both bench arms, launch list, ncu captures of every timed kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2h_tests.log
tail -6 gpurun_out/r2h_tests.log
python __graft_entry__.py smoke > gpurun_out/r2h_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2h_smoke.log
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2h_bench_ref.json 2> gpurun_out/r2h_bench.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench.json 2>> gpurun_out/r2h_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2h_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2h_launches.csv \
  python bench.py --steps 3 --warmup 3 --methods es,es14,bn,mcmc --bn-variants 20000 --mcmc-variants 37888 --no-cpu-baseline > gpurun_out/r2h_ncu_bench.log 2>&1
bash profiles/ncu_capture_r2.sh r2h es bn mcmc es14 > gpurun_out/r2h_ncu.log 2>&1
tail -8 gpurun_out/r2h_ncu.log
