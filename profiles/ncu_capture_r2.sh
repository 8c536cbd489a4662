#!/bin/bash
# Round-2 ncu captures (one GPU).  Each program has already exited 0 without ncu in the same gpurun call.
# usage: bash profiles/ncu_capture_r2.sh <tag> [es] [bn] [mcmc] [es14]
set -u
tag=$1; shift
mkdir -p gpurun_out
for what in "$@"; do
  case $what in
    es)   k=es_nuclear_kernel; skip=1; n=1;;   # every layout is launched twice: skip the first of each
    bn)   k=bn_kernel; skip=1; n=1;;
    mcmc) k=famseq_gibbs; skip=1; n=1;;
    es14) k=famseq_es; skip=1; n=1;;
  esac
  timeout 900 python profiles/ncu_targets.py $what > gpurun_out/${tag}_${what}_plain.log 2>&1 || { echo "$what failed without ncu"; continue; }
  if [ $what = es ]; then
    # three layouts x two launches: capture launches 2, 4, 6 (ids 1, 3, 5)
    for id in 1 3 5; do
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $id -c 1 -f -o gpurun_out/${tag}_es_l$id \
        python profiles/ncu_targets.py es > gpurun_out/${tag}_es_l$id.log 2>&1
    done
  else
    timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c $n -f -o gpurun_out/${tag}_$what \
      python profiles/ncu_targets.py $what > gpurun_out/${tag}_$what.log 2>&1
  fi
done
ls -la gpurun_out/*.ncu-rep
