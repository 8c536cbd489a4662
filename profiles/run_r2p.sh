#!/bin/bash
# generated ES kernel as a tile pipeline (persistent blocks, prefetch under the arithmetic)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -x -k "es_generated or random_pedigree or smoke or partial_sequencing" 2>&1 | tail -5
{
FAMSEQ_ES_JIT=1 python profiles/es_time.py ped14 1000000
FAMSEQ_ES_JIT=1 python profiles/es_time.py ped14 4000000
FAMSEQ_ES_JIT=1 FAMSEQ_ES_JIT_GRID=16 python profiles/es_time.py ped14 4000000
FAMSEQ_ES_JIT=1 FAMSEQ_ES_JIT_GRID=4 python profiles/es_time.py ped14 4000000
FAMSEQ_ES_JIT=1 FAMSEQ_ES_JIT_BLOCKS=9 python profiles/es_time.py ped14 4000000
FAMSEQ_ES_JIT=1 python profiles/es_time.py half_sibs 4000000
FAMSEQ_ES_JIT=1 python profiles/es_time.py three_wives 4000000
} > gpurun_out/r2p_es14.log 2>&1
cat gpurun_out/r2p_es14.log | cut -c 1-160
