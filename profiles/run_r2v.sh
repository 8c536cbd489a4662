#!/bin/bash
# streaming nuclear kernel: lists of consecutive tiles, tiles per block
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -x -k "compact or nuclear or phred" 2>&1 | tail -3
{
for t in 2 4 8 16 37; do
  echo "tiles per block $t"
  FAMSEQ_ES_STREAM=$t python profiles/es_time.py nuclear 1 10000000 compact
  FAMSEQ_ES_STREAM=$t python profiles/es_time.py nuclear 1 10000000 compact_no_single
done
for t in 4 8 16; do
  FAMSEQ_ES_STREAM=$t python profiles/es_time.py nuclear 2 10000000 compact
  FAMSEQ_ES_STREAM=$t python profiles/es_time.py nuclear 3 10000000 compact
done
} > gpurun_out/r2v_stream.log 2>&1
cat gpurun_out/r2v_stream.log | cut -c 1-160
