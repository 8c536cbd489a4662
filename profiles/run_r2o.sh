#!/bin/bash
# register budgets of the four- / five-child nuclear kernels; parity of all nuclear kernels; trio bench unchanged?
mkdir -p gpurun_out
{
for c in 4 5; do python profiles/es_time.py nuclear $c 4000000; FAMSEQ_ES_MINB_ALT=1 python profiles/es_time.py nuclear $c 4000000; done
for c in 1 2 3; do python profiles/es_time.py nuclear $c 4000000; done
} > gpurun_out/r2o_nuclear.log 2>&1
cat gpurun_out/r2o_nuclear.log | cut -c 1-160
python -m pytest tests -m gpu -q --timeout 900 -k "nuclear or trio or compact or cli or golden" 2>&1 | tail -2
