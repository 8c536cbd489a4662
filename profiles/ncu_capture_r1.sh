set -x
CMD="python bench.py --steps 2 --warmup 1 --variants 2000000 --bn-variants 20000 --mcmc-variants 20000 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:es_kernel -s 1 -c 1 -o gpurun_out/es_r1 -f $CMD > gpurun_out/ncu_es.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bn_kernel -s 1 -c 1 -o gpurun_out/bn_r1 -f $CMD > gpurun_out/ncu_bn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mcmc_kernel -s 1 -c 1 -o gpurun_out/mcmc_r1 -f $CMD > gpurun_out/ncu_mcmc.log 2>&1
tail -3 gpurun_out/plain.log | cut -c1-400
ls -la gpurun_out
