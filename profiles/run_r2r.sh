#!/bin/bash
# streaming nuclear kernel: start delay between the resident blocks of an SM
mkdir -p gpurun_out
{
for ns in 0 40 130 400 1500; do
  echo "stagger $ns ns"
  for layout in compact compact_no_single; do
    FAMSEQ_ES_STAGGER_NS=$ns python profiles/es_time.py nuclear 1 10000000 $layout
  done
done
FAMSEQ_ES_STAGGER_NS=130 python profiles/es_time.py nuclear 2 10000000 compact
FAMSEQ_ES_STAGGER_NS=400 python profiles/es_time.py nuclear 2 10000000 compact
} > gpurun_out/r2r_stagger.log 2>&1
cat gpurun_out/r2r_stagger.log | cut -c 1-160
