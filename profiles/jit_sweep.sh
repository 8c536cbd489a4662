# MCMC ped40: table-driven kernel vs the run-time specialised one, and layout variants of the latter.
python -m pytest tests -m gpu -q --timeout 900 -k "mcmc" 2>&1 | tail -3
run() { # label, env...
  label=$1; shift
  env "$@" python bench.py --methods mcmc --variants 1000000 --mcmc-variants ${MV:-300000} --steps 2 --warmup 1 --no-cpu-baseline 2> gpurun_out/jit_$label.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); m=d['methods']['MCMC']; print('$label', m['value'], 'variants/s', m['ms_per_step'], 'ms')"
  grep -E "Used|spill|Gibbs JIT" gpurun_out/jit_$label.err | head -4
}
run generic FAMSEQ_MCMC_JIT=0
run jit_default FAMSEQ_MCMC_JIT=1 FAMSEQ_JIT_VERBOSE=1
run jit_tb224_r0 FAMSEQ_MCMC_JIT=1 FAMSEQ_JIT_VERBOSE=1 FAMSEQ_JIT_TB=224 FAMSEQ_JIT_RACC=0
run jit_tb256_r6 FAMSEQ_MCMC_JIT=1 FAMSEQ_JIT_VERBOSE=1 FAMSEQ_JIT_TB=256 FAMSEQ_JIT_RACC=6
run jit_tb320_r12 FAMSEQ_MCMC_JIT=1 FAMSEQ_JIT_VERBOSE=1 FAMSEQ_JIT_TB=320 FAMSEQ_JIT_RACC=12
run jit_tb352_r14 FAMSEQ_MCMC_JIT=1 FAMSEQ_JIT_VERBOSE=1 FAMSEQ_JIT_TB=352 FAMSEQ_JIT_RACC=14
run jit_tb288_r10_lk8 FAMSEQ_MCMC_JIT=1 FAMSEQ_JIT_VERBOSE=1 FAMSEQ_JIT_TB=288 FAMSEQ_JIT_RACC=10 FAMSEQ_JIT_RLK=8
