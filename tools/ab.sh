#!/bin/bash
# tools/ab.sh lib1.so lib2.so ... -- A/B the bench kernel times of several builds of the library (run under gpurun)
for L in "$@"; do
  for f in ${FIELDS:-random}; do
  JSVM_ME_LIB=$PWD/$L timeout 120 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-e2e --parity-fraction ${PARITY:-0.01} --pairs ${PAIRS:-8} --field $f 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$L', '$f', round(d['value'],1), round(d['roofline']['frac'],4), {k: round(v,3) for k,v in d['kernel_ms_per_step'].items()})"
  done
done
