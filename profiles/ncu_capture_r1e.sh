CMD="python bench.py --steps 2 --warmup 1 --variants 2000000 --bn-variants 30000 --mcmc-variants 45000 --no-cpu-baseline"
$CMD > gpurun_out/plain_e.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bn_kernel -s 1 -c 1 -o gpurun_out/bn_r1e -f $CMD > gpurun_out/ncu_bn_e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mcmc_kernel -s 1 -c 1 -o gpurun_out/mcmc_r1e -f $CMD > gpurun_out/ncu_mcmc_e.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1e.csv $CMD > gpurun_out/ncu_list_e.log 2>&1
tail -1 gpurun_out/plain_e.log | cut -c1-200
