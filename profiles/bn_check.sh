python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 900 -k "bn or golden or random or ped14 or large" 2>&1 | tail -4
for G in 0 1; do
  FAMSEQ_BN_GENERIC=$G python bench.py --methods bn --variants 1000000 --bn-variants 400000 --steps 3 --no-cpu-baseline > gpurun_out/bn_g$G.json 2> gpurun_out/bn_g$G.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bn_g$G.json").read().strip().splitlines()[-1]); m=d["methods"]["BN"]
print("BN generic=$G:", m["value"], "variants/s", m["ms_per_step"], "ms", m["roofline"]["frac"])
PY
done
