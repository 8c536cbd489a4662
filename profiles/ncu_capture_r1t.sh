# headline kernel (ES, trio, nuclear-family kernel), final version: full ncu set on one 10 M-variant launch
CMD="python bench.py --steps 2 --warmup 1 --methods es --no-cpu-baseline"
$CMD > gpurun_out/plain_t.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:es_nuclear -s 3 -c 1 -o gpurun_out/es_nuclear_r1t -f $CMD > gpurun_out/ncu_esn_t.log 2>&1
tail -1 gpurun_out/plain_t.log | cut -c1-200; tail -2 gpurun_out/ncu_esn_t.log
