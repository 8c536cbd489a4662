python -m pytest tests -m gpu -q --timeout 900 -k "mcmc and not cli" 2>&1 | tail -3
run() { # label, env...
  label=$1; shift
  env FAMSEQ_MCMC_JIT=1 FAMSEQ_JIT_VERBOSE=1 "$@" python bench.py --methods mcmc --variants 1000000 --mcmc-variants ${MV:-300000} --steps 2 --warmup 1 --no-cpu-baseline 2> gpurun_out/jit_$label.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); m=d['methods']['MCMC']; print('$label', m['value'], 'variants/s', m['ms_per_step'], 'ms')"
  grep -E "Used|spill|Gibbs JIT" gpurun_out/jit_$label.err | head -3 | cut -c1-150
}
run default_with_x
