#!/bin/bash
mkdir -p gpurun_out
{
python profiles/mcmc_time.py ped40 500000
FAMSEQ_JIT_COLD=1 python profiles/mcmc_time.py ped40 200000
for preg in 0 16 24 36; do FAMSEQ_JIT_PREG=$preg python profiles/mcmc_time.py ped40 500000; done
FAMSEQ_JIT_TB=384 FAMSEQ_JIT_PREG=8 python profiles/mcmc_time.py ped40 500000
FAMSEQ_JIT_TB=320 FAMSEQ_JIT_PREG=16 python profiles/mcmc_time.py ped40 500000
FAMSEQ_JIT_TB=128 FAMSEQ_JIT_BLOCKS=2 FAMSEQ_JIT_PREG=30 python profiles/mcmc_time.py ped40 500000
python profiles/mcmc_time.py ped40 100000 1000 10000 flat
python profiles/mcmc_time.py ped40 100000 1000 10000 partial
python profiles/mcmc_time.py ped14 500000
python profiles/mcmc_time.py trio 2000000 100 1000
} > gpurun_out/r2d_mcmc_sweep.log 2>&1
cat gpurun_out/r2d_mcmc_sweep.log | cut -c 1-200
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2d_tests.log
tail -8 gpurun_out/r2d_tests.log
python bench.py --steps 10 --warmup 3 --methods es,cli --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2d_bench.err
bash profiles/ncu_capture_r2.sh r2d mcmc > gpurun_out/r2d_ncu.log 2>&1; tail -2 gpurun_out/r2d_ncu.log
