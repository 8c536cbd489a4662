#!/bin/bash
# register budgets of the quad / three-child nuclear kernels
mkdir -p gpurun_out
{
for c in 2 3; do python profiles/es_time.py nuclear $c 4000000; FAMSEQ_ES_MINB_ALT=1 python profiles/es_time.py nuclear $c 4000000; done
python profiles/es_time.py nuclear 4 4000000; python profiles/es_time.py nuclear 5 4000000
} > gpurun_out/r2n_nuclear.log 2>&1
cat gpurun_out/r2n_nuclear.log | cut -c 1-160
python -m pytest tests -m gpu -q --timeout 900 -k "nuclear" 2>&1 | tail -2
FAMSEQ_ES_MINB_ALT=1 python -m pytest tests -m gpu -q --timeout 900 -k "nuclear" 2>&1 | tail -2
