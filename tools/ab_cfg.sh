#!/bin/bash
# tools/ab_cfg.sh <config> lib1.so lib2.so ... -- A/B the bench kernel times of several builds of the library on one configuration
CFG=$1; shift
for L in "$@"; do
  for f in ${FIELDS:-random}; do
  JSVM_ME_LIB=$PWD/$L timeout 300 python bench.py --config $CFG --steps 4 --warmup 2 --no-cpu-baseline --no-e2e --no-extras --parity-fraction ${PARITY:-0.01} --pairs ${PAIRS:-8} --field $f 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$L', '$CFG', '$f', round(d['value'],1), round(d['roofline']['frac'],4), d['roofline']['kernel'], {k: round(v,3) for k,v in d['kernel_ms_per_step'].items()}, d['parity']['mismatches'])"
  done
done
