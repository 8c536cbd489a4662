#!/bin/bash
# ncu of the streaming nuclear kernel, compact input (trio, 10 M variants): where does a tile's time go?
mkdir -p gpurun_out
tag=r2s
FAMSEQ_ES_STAGGER_NS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:es_nuclear -s 3 -c 1 -f -o gpurun_out/${tag}_es_l3 \
  python profiles/ncu_targets.py es > gpurun_out/${tag}_es_l3.log 2>&1
python profiles/summarize.py $tag gpurun_out/${tag}_es_l3.ncu-rep > /dev/null 2>&1
cp profiles/${tag}_${tag}_es_l3.txt gpurun_out/ 2>/dev/null
ncu -i gpurun_out/${tag}_es_l3.ncu-rep --page source --csv --print-source sass 2>/dev/null | cut -d, -f1-10 | gzip > gpurun_out/${tag}_es_l3_sass.csv.gz
rm -f gpurun_out/${tag}_es_l3.ncu-rep
ls -la gpurun_out | tail
