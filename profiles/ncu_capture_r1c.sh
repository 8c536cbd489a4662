set -x
CMD="python bench.py --steps 2 --warmup 1 --variants 2000000 --bn-variants 20000 --mcmc-variants 20000 --no-cpu-baseline"
$CMD > gpurun_out/plain_c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:es_nuclear -s 1 -c 1 -o gpurun_out/es_nuclear_r1c -f $CMD > gpurun_out/ncu_esn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mcmc_kernel -s 1 -c 1 -o gpurun_out/mcmc_r1c -f $CMD > gpurun_out/ncu_mcmc_c.log 2>&1
tail -2 gpurun_out/plain_c.log | cut -c1-300
for U in 3 4; do echo "== BN unroll $U"; FAMSEQ_BN_UNROLL=$U python bench.py --methods bn --variants 1000000 --bn-variants 200000 --steps 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); m=d['methods']['BN']; print(m['value'], m['ms_per_step'], m['roofline']['frac'])"; done
