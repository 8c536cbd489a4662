# trio ES (nuclear-family kernel): variants per block
timeout 300 python -m pytest tests -m gpu -q --timeout 200 -k "nuclear or trio or golden or large or empty or device_path" 2>&1 | tail -2
for tb in 32 64 128; do
  FAMSEQ_ES_TB=$tb timeout 300 python bench.py --methods es --steps 20 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tb=$tb', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'])"
done
FAMSEQ_ES_TB=32 timeout 300 python -m pytest tests -m gpu -q --timeout 200 -k "nuclear or trio" 2>&1 | tail -1
