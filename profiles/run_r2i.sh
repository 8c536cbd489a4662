#!/bin/bash
# round 2 record set (outputs kept small): MCMC tests incl. the pilot, both bench arms, launch list, ncu summaries
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -k "mcmc or gibbs or pilot or random_pedigree" > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2i_tests.log
tail -4 gpurun_out/r2i_tests.log
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2i_bench_ref.json 2> gpurun_out/r2i_bench.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench.json 2>> gpurun_out/r2i_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2i_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2i_launches.csv \
  python bench.py --steps 3 --warmup 3 --methods es,es14,bn,mcmc --bn-variants 20000 --mcmc-variants 37888 --no-cpu-baseline > gpurun_out/r2i_ncu_bench.log 2>&1
bash profiles/ncu_capture_r2.sh r2i es bn mcmc es14 > gpurun_out/r2i_ncu.log 2>&1
tail -5 gpurun_out/r2i_ncu.log
du -sh gpurun_out
