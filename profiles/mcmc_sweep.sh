python -m pytest tests -m gpu -q --timeout 900 -k "mcmc or smoke or loops" 2>&1 | tail -2
for B in 1 2 3; do
  FAMSEQ_MCMC_BLOCKS=$B python bench.py --methods mcmc --variants 1000000 --mcmc-variants 500000 --steps 3 --no-cpu-baseline > gpurun_out/mcmc_b$B.json 2> gpurun_out/mcmc_b$B.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/mcmc_b$B.json").read().strip().splitlines()[-1]); m=d["methods"]["MCMC"]
print("MCMC wglobal blocks=$B:", m["value"], "variants/s", m["ms_per_step"], "ms", m["roofline"]["frac"])
PY
done
FAMSEQ_MCMC_WGLOBAL=0 python bench.py --methods mcmc --variants 1000000 --mcmc-variants 500000 --steps 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); m=d['methods']['MCMC']; print('MCMC smem:', m['value'], m['ms_per_step'])"
