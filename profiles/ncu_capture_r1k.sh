export FAMSEQ_MCMC_JIT=1 FAMSEQ_JIT_TB=384 FAMSEQ_JIT_RACC=4 FAMSEQ_JIT_SACC=0 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=24 FAMSEQ_JIT_PF=3
CMD="python bench.py --steps 1 --warmup 1 --variants 1000000 --methods mcmc --mcmc-variants 56832 --no-cpu-baseline"
$CMD > gpurun_out/plain_k.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:famseq_gibbs -s 1 -c 1 -o gpurun_out/mcmc_r1k -f $CMD > gpurun_out/ncu_mcmc_k.log 2>&1
tail -1 gpurun_out/plain_k.log | cut -c1-100; tail -3 gpurun_out/ncu_mcmc_k.log
