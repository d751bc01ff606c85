#!/bin/bash
# tools/profile.sh -- ncu launch list + full captures of K1/K2/K3 (run under gpurun; results in gpurun_out/)
set -x
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e --pairs ${PAIRS:-4}"
python bench.py $ARGS > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:full_search -s 1 -c 1 -f -o gpurun_out/prof_k2 python bench.py $ARGS > gpurun_out/ncu_k2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:subpel -s 1 -c 1 -f -o gpurun_out/prof_k3 python bench.py $ARGS > gpurun_out/ncu_k3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:halfpel -s 1 -c 1 -f -o gpurun_out/prof_k1 python bench.py $ARGS > gpurun_out/ncu_k1.log 2>&1
tail -3 gpurun_out/plain.log; ls -la gpurun_out
