#!/bin/bash
# Gibbs JIT v3 with the per-warp vote: layouts, tests, ncu
mkdir -p gpurun_out
{
python profiles/mcmc_time.py ped40 500000
FAMSEQ_JIT_TB=384 FAMSEQ_JIT_PREG=28 python profiles/mcmc_time.py ped40 500000
FAMSEQ_JIT_TB=512 FAMSEQ_JIT_PREG=10 python profiles/mcmc_time.py ped40 500000
FAMSEQ_JIT_TB=512 FAMSEQ_JIT_PREG=0 python profiles/mcmc_time.py ped40 500000
FAMSEQ_JIT_TB=768 FAMSEQ_JIT_PREG=0 python profiles/mcmc_time.py ped40 500000
FAMSEQ_JIT_TB=1024 FAMSEQ_JIT_PREG=0 python profiles/mcmc_time.py ped40 500000
FAMSEQ_JIT_TB=256 FAMSEQ_JIT_BLOCKS=2 FAMSEQ_JIT_PREG=0 python profiles/mcmc_time.py ped40 500000
python profiles/mcmc_time.py ped40 100000 1000 10000 flat
python profiles/mcmc_time.py ped40 100000 1000 10000 random
python profiles/mcmc_time.py ped14 1000000
python profiles/mcmc_time.py trio 4000000 100 1000
python profiles/mcmc_time.py ped100 100000
} > gpurun_out/r2g_mcmc_sweep.log 2>&1
cut -c 1-175 gpurun_out/r2g_mcmc_sweep.log
python -m pytest tests -m gpu -q --timeout 900 -k "mcmc or gibbs or random_pedigree or smoke" > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2g_tests.log
tail -4 gpurun_out/r2g_tests.log
bash profiles/ncu_capture_r2.sh r2g mcmc > gpurun_out/r2g_ncu.log 2>&1; tail -2 gpurun_out/r2g_ncu.log
