# BN ped14: exhaustive enumeration (default) against the opt-in closed-form sum over the childless innermost block
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "analytically" 2>&1 | tail -2
for f in 0 1; do
  FAMSEQ_BN_FACTOR=$f timeout 300 python bench.py --methods bn --variants 1000000 --steps 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); m=d['methods']['BN']; print('factor=$f', m['value'], m['ms_per_step'])"
done
