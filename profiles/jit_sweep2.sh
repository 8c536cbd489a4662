# MCMC ped40: layout variants of the run-time specialised Gibbs kernel.
python -m pytest tests -m gpu -q --timeout 900 -k "mcmc" 2>&1 | tail -3
run() { # label, env...
  label=$1; shift
  env FAMSEQ_MCMC_JIT=1 FAMSEQ_JIT_VERBOSE=1 "$@" python bench.py --methods mcmc --variants 1000000 --mcmc-variants ${MV:-300000} --steps 2 --warmup 1 --no-cpu-baseline 2> gpurun_out/jit_$label.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); m=d['methods']['MCMC']; print('$label', m['value'], 'variants/s', m['ms_per_step'], 'ms')"
  grep -E "Used|spill|Gibbs JIT" gpurun_out/jit_$label.err | head -3
}
run default
run pf1 FAMSEQ_JIT_PF=1
run pf2 FAMSEQ_JIT_PF=2
run pf5 FAMSEQ_JIT_PF=5
run acc6_34_lkL2_pf3 FAMSEQ_JIT_RACC=6 FAMSEQ_JIT_SACC=34 FAMSEQ_JIT_SLK=0 FAMSEQ_JIT_PF=3
run acc6_34_lkL2_pf6 FAMSEQ_JIT_RACC=6 FAMSEQ_JIT_SACC=34 FAMSEQ_JIT_SLK=0 FAMSEQ_JIT_PF=6
run accRED_lksmem FAMSEQ_JIT_RACC=0 FAMSEQ_JIT_SACC=0 FAMSEQ_JIT_RLK=4 FAMSEQ_JIT_SLK=36
run tb384_a FAMSEQ_JIT_TB=384 FAMSEQ_JIT_RACC=6 FAMSEQ_JIT_SACC=24 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=0 FAMSEQ_JIT_PF=3
run tb384_b FAMSEQ_JIT_TB=384 FAMSEQ_JIT_RACC=4 FAMSEQ_JIT_SACC=0 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=24 FAMSEQ_JIT_PF=3
run tb512_a FAMSEQ_JIT_TB=512 FAMSEQ_JIT_RACC=0 FAMSEQ_JIT_SACC=0 FAMSEQ_JIT_RLK=0 FAMSEQ_JIT_SLK=18 FAMSEQ_JIT_PF=2
run tb128x2 FAMSEQ_JIT_TB=128 FAMSEQ_JIT_BLOCKS=2
