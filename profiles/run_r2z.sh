#!/bin/bash
# generated ES kernel: warp-uniform range checks (vote) instead of per-lane branches around the divisions
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -x -k "es_generated or random_pedigree or smoke or partial_sequencing" 2>&1 | tail -3
{
FAMSEQ_ES_JIT=1 python profiles/es_time.py ped14 4000000
FAMSEQ_ES_JIT=1 python profiles/es_time.py ped14 1000000
FAMSEQ_ES_JIT=1 python profiles/es_time.py half_sibs 4000000
FAMSEQ_ES_JIT=1 python profiles/es_time.py three_wives 4000000
} > gpurun_out/r2z_es14.log 2>&1
cat gpurun_out/r2z_es14.log | cut -c 1-160
