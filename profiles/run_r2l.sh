#!/bin/bash
# generated ES kernel (ped14): warps per block
mkdir -p gpurun_out
{
for tb in 32 64 128 256; do FAMSEQ_ES_JIT=1 FAMSEQ_ES_JIT_TB=$tb python profiles/es_time.py ped14 1000000; done
FAMSEQ_ES_JIT=1 FAMSEQ_ES_JIT_TB=128 FAMSEQ_ES_JIT_BLOCKS=2 python profiles/es_time.py ped14 1000000
FAMSEQ_ES_JIT=1 FAMSEQ_ES_JIT_TB=64 FAMSEQ_ES_JIT_BLOCKS=6 python profiles/es_time.py ped14 1000000
FAMSEQ_ES_JIT=1 FAMSEQ_ES_JIT_TB=32 FAMSEQ_ES_JIT_BLOCKS=12 python profiles/es_time.py ped14 1000000
} > gpurun_out/r2l_es14.log 2>&1
cat gpurun_out/r2l_es14.log | cut -c 1-160
FAMSEQ_ES_JIT_TB=128 python -m pytest tests -m gpu -q --timeout 900 -k "es_generated or random_pedigree or smoke" 2>&1 | tail -2
