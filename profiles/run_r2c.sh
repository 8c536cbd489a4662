#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c_tests.log
tail -12 gpurun_out/r2c_tests.log
{
python profiles/mcmc_time.py ped40 500000
for preg in 16 24 32 38; do FAMSEQ_JIT_PREG=$preg python profiles/mcmc_time.py ped40 500000; done
FAMSEQ_JIT_TB=384 FAMSEQ_JIT_PREG=10 python profiles/mcmc_time.py ped40 500000
FAMSEQ_JIT_TB=512 FAMSEQ_JIT_PREG=0 python profiles/mcmc_time.py ped40 500000
FAMSEQ_JIT_TB=128 FAMSEQ_JIT_BLOCKS=2 FAMSEQ_JIT_PREG=32 python profiles/mcmc_time.py ped40 500000
python profiles/mcmc_time.py ped40 100000 1000 10000 flat
python profiles/mcmc_time.py ped40 100000 1000 10000 partial
FAMSEQ_MCMC_JIT=0 python profiles/mcmc_time.py ped40 100000 1000 10000 partial
FAMSEQ_MCMC_JIT=0 python profiles/mcmc_time.py ped40 100000 1000 10000 flat
python profiles/mcmc_time.py ped14 500000
python profiles/mcmc_time.py trio 2000000 100 1000
} > gpurun_out/r2c_mcmc_sweep.log 2>&1
cat gpurun_out/r2c_mcmc_sweep.log | cut -c 1-220
for c in 2 3 5; do python profiles/es_time.py nuclear $c 4000000; done >> gpurun_out/r2c_mcmc_sweep.log 2>&1
FAMSEQ_ES_JIT=1 python profiles/es_time.py ped14 1000000 >> gpurun_out/r2c_mcmc_sweep.log 2>&1
tail -4 gpurun_out/r2c_mcmc_sweep.log
python bench.py --steps 10 --warmup 3 --methods es --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
