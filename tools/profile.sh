#!/bin/bash
# tools/profile.sh -- ncu launch list + `--set full` captures of K1 / K2 / K3 for the bench workload (run under gpurun).
# Writes gpurun_out/{launches.csv, prof_k*.ncu-rep, ncu_k*.json, traffic.json}; copy the summaries to profiles/ afterwards.
set -x
PAIRS=${PAIRS:-16}
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras --parity-fraction 0 --pairs $PAIRS"
python bench.py $ARGS > gpurun_out/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:full_search -s 1 -c 1 -f -o gpurun_out/prof_k2 python bench.py $ARGS > gpurun_out/ncu_k2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:subpel -s 1 -c 1 -f -o gpurun_out/prof_k3 python bench.py $ARGS > gpurun_out/ncu_k3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:halfpel -s 1 -c 1 -f -o gpurun_out/prof_k1 python bench.py $ARGS > gpurun_out/ncu_k1.log 2>&1
for k in k1 k2 k3; do python tools/summarize_profile.py gpurun_out/prof_$k.ncu-rep > gpurun_out/ncu_$k.json; done
python tools/make_traffic.py $PAIRS gpurun_out/ncu_k1.json gpurun_out/ncu_k2.json gpurun_out/ncu_k3.json > gpurun_out/traffic.json
ncu -i gpurun_out/prof_k2.ncu-rep --page source --csv > gpurun_out/k2_source.csv 2>/dev/null
tail -c 400 gpurun_out/plain.log; cat gpurun_out/traffic.json
