export FAMSEQ_ES_PERSISTENT=1
timeout 300 python -m pytest tests -m gpu -q --timeout 200 -k "nuclear or trio or golden or large or empty or device_path or lrc" 2>&1 | tail -3
for p in 0 1; do
  FAMSEQ_ES_PERSISTENT=$p timeout 300 python bench.py --methods es --steps 20 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('persistent=$p', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'])"
done
