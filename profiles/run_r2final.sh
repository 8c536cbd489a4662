#!/bin/bash
# last validation of the committed tree: full GPU suite, smoke, both bench arms (the driver's flags)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2final_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2final_tests.log
tail -4 gpurun_out/r2final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2final_smoke.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2final_bench_ref.json 2> gpurun_out/r2final_bench.err; echo "ref rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2final_bench.json 2>> gpurun_out/r2final_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2final_bench.err
