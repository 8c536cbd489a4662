# BN (ped14, U = 5 register block) and the generated Gibbs kernel with the final default layout
CMD="python bench.py --steps 1 --warmup 1 --variants 1000000 --methods bn,mcmc --bn-variants 40000 --mcmc-variants 37888 --no-cpu-baseline"
$CMD > gpurun_out/plain_q.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bn_kernel -s 1 -c 1 -o gpurun_out/bn_r1q -f $CMD > gpurun_out/ncu_bn_q.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:famseq_gibbs -s 1 -c 1 -o gpurun_out/mcmc_r1q -f $CMD > gpurun_out/ncu_mcmc_q.log 2>&1
tail -1 gpurun_out/plain_q.log | cut -c1-200; tail -2 gpurun_out/ncu_bn_q.log; tail -2 gpurun_out/ncu_mcmc_q.log
