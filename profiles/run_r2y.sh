#!/bin/bash
# where the command line's wall clock goes (finer timers)
mkdir -p gpurun_out
python profiles/cli_stats.py 2000000 4 > gpurun_out/r2y_cli.log 2>&1
cat gpurun_out/r2y_cli.log | cut -c 1-600
