#!/bin/bash
# round 2, first GPU pass: parity tests, smoke, bench, ncu launch list + full captures
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
python __graft_entry__.py smoke > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2a_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2a_bench.err
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2a_bench_ref.json 2>> gpurun_out/r2a_bench.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv \
  python bench.py --steps 3 --warmup 3 --methods es --no-cpu-baseline > gpurun_out/r2a_ncu_bench.log 2>&1
bash profiles/ncu_capture_r2.sh r2a es bn > gpurun_out/r2a_ncu.log 2>&1
tail -3 gpurun_out/r2a_ncu.log
