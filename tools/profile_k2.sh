#!/bin/bash
# tools/profile_k2.sh -- one ncu --set full capture of the full-search kernel (run under gpurun)
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e --pairs ${PAIRS:-2} ${EXTRA}"
python bench.py $ARGS > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:full_search -s 1 -c 1 -f -o gpurun_out/prof_k2 python bench.py $ARGS > gpurun_out/ncu_k2.log 2>&1
tail -c 600 gpurun_out/plain.log
