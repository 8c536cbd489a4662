# generated ES kernel on ped14: resident blocks per SM forced through __launch_bounds__ (registers vs spills)
for b in 0 10; do
  FAMSEQ_ES_JIT_BLOCKS=$b timeout 200 python bench.py --methods es14 --variants 1000000 --steps 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); m=d['methods']['ES_ped14']; print('blocks=$b', m['value'], m['ms_per_step'], m['roofline']['frac'])"
done
