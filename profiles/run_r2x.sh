#!/bin/bash
# ncu of the compact-input trio kernels of the final tree (the r2w capture filtered on the old kernel name)
mkdir -p gpurun_out
bash profiles/ncu_capture_r2.sh r2x es > gpurun_out/r2x_ncu.log 2>&1
ls gpurun_out | grep r2x
