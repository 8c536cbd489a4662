#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2b_tests.log
tail -15 gpurun_out/r2b_tests.log
python bench.py --steps 20 --warmup 5 --methods es,cli --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2b_bench.err
bash profiles/ncu_capture_r2.sh r2b es > gpurun_out/r2b_ncu.log 2>&1
