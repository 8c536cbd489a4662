CMD="python bench.py --steps 2 --warmup 1 --variants 1000000 --methods mcmc --mcmc-variants 113664 --no-cpu-baseline"
$CMD > gpurun_out/plain_i.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mcmc_kernel -s 1 -c 1 -o gpurun_out/mcmc_r1i -f $CMD > gpurun_out/ncu_mcmc_i.log 2>&1
tail -1 gpurun_out/plain_i.log | cut -c1-100
