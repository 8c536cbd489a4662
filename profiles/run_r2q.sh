#!/bin/bash
# nuclear kernel with compact input as a tile pipeline (FAMSEQ_ES_STREAM=0: one tile per block, as before) + ncu of the ES JIT pipeline
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -x -k "compact or nuclear or phred or cli or golden or smoke" 2>&1 | tail -5
{
for layout in compact compact_no_single; do
  for c in 1 2 3; do
    python profiles/es_time.py nuclear $c 10000000 $layout
    FAMSEQ_ES_STREAM=0 python profiles/es_time.py nuclear $c 10000000 $layout
  done
done
python profiles/es_time.py nuclear 1 10000000
} > gpurun_out/r2q_stream.log 2>&1
cat gpurun_out/r2q_stream.log | cut -c 1-160
bash profiles/ncu_capture_r2.sh r2q es14 > gpurun_out/r2q_ncu.log 2>&1
ls gpurun_out | head -30
