# MCMC ped40: layout variants of the branch-free specialised Gibbs kernel.
python -m pytest tests -m gpu -q --timeout 900 -k "mcmc" 2>&1 | tail -3
run() { # label, env...
  label=$1; shift
  env FAMSEQ_MCMC_JIT=1 FAMSEQ_JIT_VERBOSE=1 "$@" python bench.py --methods mcmc --variants 1000000 --mcmc-variants ${MV:-300000} --steps 2 --warmup 1 --no-cpu-baseline 2> gpurun_out/jit_$label.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); m=d['methods']['MCMC']; print('$label', m['value'], 'variants/s', m['ms_per_step'], 'ms')"
  grep -E "Used|spill|Gibbs JIT" gpurun_out/jit_$label.err | head -3 | cut -c1-150
}
run default
run tb256_accRED FAMSEQ_JIT_RACC=8 FAMSEQ_JIT_SACC=0 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=36 FAMSEQ_JIT_PF=3
run tb384_b FAMSEQ_JIT_TB=384 FAMSEQ_JIT_RACC=4 FAMSEQ_JIT_SACC=0 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=24 FAMSEQ_JIT_PF=3
run tb384_c FAMSEQ_JIT_TB=384 FAMSEQ_JIT_RACC=0 FAMSEQ_JIT_SACC=0 FAMSEQ_JIT_RLK=4 FAMSEQ_JIT_SLK=24 FAMSEQ_JIT_PF=3
run tb384_d FAMSEQ_JIT_TB=384 FAMSEQ_JIT_RACC=0 FAMSEQ_JIT_SACC=0 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=24 FAMSEQ_JIT_PF=4
run tb512_a FAMSEQ_JIT_TB=512 FAMSEQ_JIT_RACC=0 FAMSEQ_JIT_SACC=0 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=18 FAMSEQ_JIT_PF=2
run tb512_b FAMSEQ_JIT_TB=512 FAMSEQ_JIT_RACC=0 FAMSEQ_JIT_SACC=0 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=18 FAMSEQ_JIT_PF=1
run tb640 FAMSEQ_JIT_TB=640 FAMSEQ_JIT_RACC=0 FAMSEQ_JIT_SACC=0 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=14 FAMSEQ_JIT_PF=1
