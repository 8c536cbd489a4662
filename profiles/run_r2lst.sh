#!/bin/bash
# ncu launch list of the bench command on the final tree (the bench itself exited 0 without ncu: profiles/r2final_bench.json)
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --methods es,es14 --no-cpu-baseline > gpurun_out/r2lst_plain.json 2> gpurun_out/r2lst_plain.err; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2final_launches.csv \
  python bench.py --steps 3 --warmup 3 --methods es,es14 --no-cpu-baseline > gpurun_out/r2lst_ncu_bench.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r2final_launches.csv
