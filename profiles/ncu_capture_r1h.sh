CMD="python bench.py --steps 2 --warmup 1 --variants 1000000 --methods es14 --es14-variants 500000 --no-cpu-baseline"
$CMD > gpurun_out/plain_h.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:es_kernel -s 1 -c 1 -o gpurun_out/es14_r1h -f $CMD > gpurun_out/ncu_es14.log 2>&1
tail -1 gpurun_out/plain_h.log | cut -c1-100
