#!/bin/bash
# validation of the tree with the tile pipelines (generated ES kernel, compact nuclear kernel): full GPU suite, smoke, both
# bench arms (default flags like the driver), launch list, ncu of the three trio layouts and of the generated ES kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2w_tests.log
tail -5 gpurun_out/r2w_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2w_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2w_smoke.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2w_bench_ref.json 2> gpurun_out/r2w_bench.err; echo "ref rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2w_bench.json 2>> gpurun_out/r2w_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2w_bench.err
bash profiles/ncu_capture_r2.sh r2w es es14 > gpurun_out/r2w_ncu.log 2>&1
ls gpurun_out | grep r2w
