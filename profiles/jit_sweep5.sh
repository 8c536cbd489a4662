# MCMC ped40: accumulators in registers/shared memory, own factors from L2 -- how many reductions (RED) can the L2 take?
run() { # label, env...
  label=$1; shift
  env FAMSEQ_MCMC_JIT=1 FAMSEQ_JIT_VERBOSE=1 "$@" python bench.py --methods mcmc --variants 1000000 --mcmc-variants ${MV:-300000} --steps 2 --warmup 1 --no-cpu-baseline 2> gpurun_out/jit_$label.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); m=d['methods']['MCMC']; print('$label', m['value'], 'variants/s', m['ms_per_step'], 'ms')"
  grep -E "Used|spill|Gibbs JIT" gpurun_out/jit_$label.err | head -3 | cut -c1-150
}
run b1 FAMSEQ_JIT_RACC=16 FAMSEQ_JIT_SACC=24 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=12 FAMSEQ_JIT_PF=2
run b2 FAMSEQ_JIT_RACC=16 FAMSEQ_JIT_SACC=24 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=12 FAMSEQ_JIT_PF=3
run b3 FAMSEQ_JIT_RACC=19 FAMSEQ_JIT_SACC=21 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=15 FAMSEQ_JIT_PF=2
run b4 FAMSEQ_JIT_RACC=16 FAMSEQ_JIT_SACC=12 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=24 FAMSEQ_JIT_PF=2
run b5 FAMSEQ_JIT_RACC=16 FAMSEQ_JIT_SACC=16 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=20 FAMSEQ_JIT_PF=2
run b6 FAMSEQ_JIT_RACC=20 FAMSEQ_JIT_SACC=16 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=20 FAMSEQ_JIT_PF=1
run b7 FAMSEQ_JIT_RACC=18 FAMSEQ_JIT_SACC=0 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=36 FAMSEQ_JIT_PF=1
run b8 FAMSEQ_JIT_RACC=20 FAMSEQ_JIT_SACC=0 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=36 FAMSEQ_JIT_PF=1
