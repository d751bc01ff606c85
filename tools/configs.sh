#!/bin/bash
# tools/configs.sh -- bench lines of C1 .. C5 x {random, zero} start fields (resident legs only), one JSON line each (run under gpurun)
for c in C1 C2 C3 C4 C5; do for f in random zero; do
  timeout 300 python bench.py --config $c --field $f --steps 6 --warmup 3 --no-cpu-baseline --no-extras --parity-fraction 0.01 --pairs 16 2>/dev/null | tail -1
done; done
