export FAMSEQ_MCMC_JIT=1 FAMSEQ_JIT_TB=256 FAMSEQ_JIT_RACC=8 FAMSEQ_JIT_SACC=0 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=36 FAMSEQ_JIT_PF=3
CMD="python bench.py --steps 1 --warmup 1 --variants 1000000 --methods mcmc --mcmc-variants 37888 --no-cpu-baseline"
$CMD > gpurun_out/plain_l.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:famseq_gibbs -s 1 -c 1 -o gpurun_out/mcmc_r1l -f $CMD > gpurun_out/ncu_mcmc_l.log 2>&1
tail -1 gpurun_out/plain_l.log | cut -c1-100; tail -2 gpurun_out/ncu_mcmc_l.log
