#!/bin/bash
# tools/profile_kernel.sh <regex> <outname> -- one ncu --set full capture of one kernel of bench.py (run under gpurun)
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras --parity-fraction 0 --pairs ${PAIRS:-2} ${EXTRA}"
python bench.py $ARGS > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$1 -s 1 -c 1 -f -o gpurun_out/$2 python bench.py $ARGS > gpurun_out/ncu_$2.log 2>&1
tail -c 300 gpurun_out/plain.log
