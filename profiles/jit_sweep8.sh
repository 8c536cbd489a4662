# MCMC ped40, final layout sweep around the default (kernel with autosomal + chrX sweeps).
python -m pytest tests -m gpu -q --timeout 900 -k "cuda_build or gibbs_kernel" 2>&1 | tail -2
run() { # label, env...
  label=$1; shift
  env FAMSEQ_MCMC_JIT=1 FAMSEQ_JIT_VERBOSE=1 "$@" python bench.py --methods mcmc --variants 1000000 --mcmc-variants ${MV:-300000} --steps 2 --warmup 1 --no-cpu-baseline 2> gpurun_out/jit_$label.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); m=d['methods']['MCMC']; print('$label', m['value'], 'variants/s', m['ms_per_step'], 'ms')"
  grep -E "Used|spill|Gibbs JIT" gpurun_out/jit_$label.err | head -3 | cut -c1-150
}
run d1
run d2 FAMSEQ_JIT_PF=2
run d3 FAMSEQ_JIT_RACC=16 FAMSEQ_JIT_SACC=8 FAMSEQ_JIT_SLK=28 FAMSEQ_JIT_PF=2
run d4 FAMSEQ_JIT_RACC=15 FAMSEQ_JIT_SACC=6 FAMSEQ_JIT_SLK=30 FAMSEQ_JIT_PF=2
run d5 FAMSEQ_JIT_RACC=16 FAMSEQ_JIT_SACC=4 FAMSEQ_JIT_SLK=32 FAMSEQ_JIT_PF=2
run d6 FAMSEQ_JIT_RACC=14 FAMSEQ_JIT_SACC=8 FAMSEQ_JIT_SLK=28 FAMSEQ_JIT_PF=2
