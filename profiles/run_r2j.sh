#!/bin/bash
# identity-column-map specialisation of the nuclear ES kernel: on / off, parity, ncu
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -k "nuclear or trio or compact or golden or ragged or lrc or phred" > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2j_tests.log
tail -3 gpurun_out/r2j_tests.log
for ident in 1 0; do
  FAMSEQ_ES_IDENT=$ident python bench.py --steps 20 --warmup 5 --methods es --no-cpu-baseline 2>/dev/null > gpurun_out/r2j_bench_ident$ident.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2j_bench_ident$ident.json").read().strip().splitlines()[-1])
print("ident=$ident canonical %.4f ms frac %.3f | compact_in %.4f ms %.3f | no_single %.4f ms %.3f | e2e %.3g phred %.3g" % (d["ms_per_step"], d["roofline"]["frac"], d["layouts"]["compact_in"]["ms_per_step"], d["layouts"]["compact_in"]["roofline"]["frac"], d["layouts"]["compact_in_no_single"]["ms_per_step"], d["layouts"]["compact_in_no_single"]["roofline"]["frac"], d["e2e"]["value"], d["e2e_phred"]["value"]))
PY
done
for c in 2 3; do python profiles/es_time.py nuclear $c 4000000; FAMSEQ_ES_IDENT=0 python profiles/es_time.py nuclear $c 4000000; done
bash profiles/ncu_capture_r2.sh r2j es > gpurun_out/r2j_ncu.log 2>&1
grep -E "time_duration|inst_executed.sum " gpurun_out/r2j_r2j_es_l*.txt | cut -c 1-140
