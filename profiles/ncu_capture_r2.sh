#!/bin/bash
# Round-2 ncu captures (one GPU).  Each program has already exited 0 without ncu in the same gpurun call.
# usage: bash profiles/ncu_capture_r2.sh <tag> [es] [bn] [mcmc] [es14]
# The .ncu-rep files are summarised on the box (profiles/summarize.py -> profiles/<tag>_*.txt, copied to gpurun_out/) and
# removed: six of them exceed what gpurun brings back.
set -u
tag=$1; shift
mkdir -p gpurun_out
summarise() { # report -> gpurun_out/<name>.txt (+ the SASS page with per-instruction execution counts, gzipped)
  python profiles/summarize.py $tag "$1" > /dev/null 2>&1
  name=$(basename "$1" .ncu-rep)
  cp profiles/${tag}_${name}.txt gpurun_out/ 2>/dev/null
  ncu -i "$1" --page source --csv --print-source sass 2>/dev/null | cut -d, -f1-10 | gzip > gpurun_out/${name}_sass.csv.gz
  rm -f "$1"
}
for what in "$@"; do
  case $what in
    es)   k=es_nuclear;;   # es_nuclear_kernel (FP64 input) and es_nuclear_stream_kernel (compact input)
    bn)   k=bn_kernel;;
    mcmc) k=famseq_gibbs;;
    es14) k=famseq_es;;
  esac
  timeout 900 python profiles/ncu_targets.py $what > gpurun_out/${tag}_${what}_plain.log 2>&1 || { echo "$what failed without ncu"; continue; }
  if [ $what = es ]; then
    # three layouts x two launches: capture launches 2, 4, 6 (ids 1, 3, 5)
    for id in 1 3 5; do
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $id -c 1 -f -o gpurun_out/${tag}_es_l$id \
        python profiles/ncu_targets.py es > gpurun_out/${tag}_es_l$id.log 2>&1
      summarise gpurun_out/${tag}_es_l$id.ncu-rep
    done
  else
    timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/${tag}_$what \
      python profiles/ncu_targets.py $what > gpurun_out/${tag}_$what.log 2>&1
    summarise gpurun_out/${tag}_$what.ncu-rep
  fi
done
ls -la gpurun_out/ | tail -30
