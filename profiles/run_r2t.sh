#!/bin/bash
# streaming nuclear kernel with the flags on the tile's TMA transaction
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -x -k "compact or nuclear or phred or cli or golden or smoke" 2>&1 | tail -5
{
for layout in compact compact_no_single; do
  for c in 1 2 3 4 5; do
    python profiles/es_time.py nuclear $c 10000000 $layout
    FAMSEQ_ES_STREAM=0 python profiles/es_time.py nuclear $c 10000000 $layout
  done
done
} > gpurun_out/r2t_stream.log 2>&1
cat gpurun_out/r2t_stream.log | cut -c 1-160
