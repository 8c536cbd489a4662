for U in 3 4; do
  FAMSEQ_BN_UNROLL=$U python bench.py --methods bn --variants 1000000 --bn-variants 300000 --steps 3 --no-cpu-baseline > gpurun_out/bn_u$U.json 2> gpurun_out/bn_u$U.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bn_u$U.json").read().strip().splitlines()[-1]); m=d["methods"]["BN"]
print("BN unroll $U:", m["value"], "variants/s", m["ms_per_step"], "ms", m["roofline"]["frac"])
PY
done
python tools/bench_cli.py --variants 1000000 > gpurun_out/cli_bench.json 2> gpurun_out/cli_bench.err; cat gpurun_out/cli_bench.json; tail -3 gpurun_out/cli_bench.err
