# MCMC ped40: layout variants of the specialised Gibbs kernel after spreading the L2-resident members over the sweep.
python -m pytest tests -m gpu -q --timeout 900 -k "mcmc" 2>&1 | tail -3
run() { # label, env...
  label=$1; shift
  env FAMSEQ_MCMC_JIT=1 FAMSEQ_JIT_VERBOSE=1 "$@" python bench.py --methods mcmc --variants 1000000 --mcmc-variants ${MV:-300000} --steps 2 --warmup 1 --no-cpu-baseline 2> gpurun_out/jit_$label.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); m=d['methods']['MCMC']; print('$label', m['value'], 'variants/s', m['ms_per_step'], 'ms')"
  grep -E "Used|spill|Gibbs JIT" gpurun_out/jit_$label.err | head -3 | cut -c1-150
}
run a256_acc8r_lk36s_pf1 FAMSEQ_JIT_RACC=8 FAMSEQ_JIT_SACC=0 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=36 FAMSEQ_JIT_PF=1
run a256_acc8r_lk36s_pf2 FAMSEQ_JIT_RACC=8 FAMSEQ_JIT_SACC=0 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=36 FAMSEQ_JIT_PF=2
run a256_acc12r_lk4r36s FAMSEQ_JIT_RACC=12 FAMSEQ_JIT_SACC=0 FAMSEQ_JIT_RLK=4 FAMSEQ_JIT_SLK=36
run a256_acc16r_lk36s_pf1 FAMSEQ_JIT_RACC=16 FAMSEQ_JIT_SACC=0 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=36 FAMSEQ_JIT_PF=1
run a256_acc0_lk4r36s FAMSEQ_JIT_RACC=0 FAMSEQ_JIT_SACC=0 FAMSEQ_JIT_RLK=4 FAMSEQ_JIT_SLK=36
run a256_acc8r8s_lk28s_pf2 FAMSEQ_JIT_RACC=8 FAMSEQ_JIT_SACC=8 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=28 FAMSEQ_JIT_PF=2
run a384_acc4r_lk24s_pf2 FAMSEQ_JIT_TB=384 FAMSEQ_JIT_RACC=4 FAMSEQ_JIT_SACC=0 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=24 FAMSEQ_JIT_PF=2
run a384_acc0_lk4r24s_pf2 FAMSEQ_JIT_TB=384 FAMSEQ_JIT_RACC=0 FAMSEQ_JIT_SACC=0 FAMSEQ_JIT_RLK=4 FAMSEQ_JIT_SLK=24 FAMSEQ_JIT_PF=2
run a384_acc8r_lk24s_pf1 FAMSEQ_JIT_TB=384 FAMSEQ_JIT_RACC=8 FAMSEQ_JIT_SACC=0 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=24 FAMSEQ_JIT_PF=1
