#!/bin/bash
# usage: scripts/ncu_one.sh <kernel-regex> <out-name> [workload]   -- one `ncu --set full` capture of matching kernels (3 launches)
set -e
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:$1" -c 3 -o gpurun_out/$2 -f python scripts/kbench.py ${3:-c2-dense128} --iters=1 > gpurun_out/$2.log 2>&1
tail -2 gpurun_out/$2.log
