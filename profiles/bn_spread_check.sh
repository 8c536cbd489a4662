# BN ped14: threads per variant (3^h) -- 243 (one variant per block) against 81 (three per block) and 27 (nine per block)
for h in 5 4 3; do
  FAMSEQ_BN_SPREAD=$h timeout 300 python bench.py --methods bn --variants 1000000 --steps 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); m=d['methods']['BN']; print('h=$h', m['value'], m['ms_per_step'])"
done
FAMSEQ_BN_SPREAD=4 timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "bn or BN or golden" 2>&1 | tail -2
