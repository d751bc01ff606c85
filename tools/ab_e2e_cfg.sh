#!/bin/bash
# tools/ab_e2e_cfg.sh CONFIG "ENV=1" ... -- A/B the end-to-end leg of bench.py --config CONFIG under environment knobs (run under gpurun)
C=$1; shift
for rep in 1 2; do
for V in "$@"; do
  env $V timeout 300 python bench.py --config $C --steps 5 --warmup 3 --no-cpu-baseline --no-extras --parity-fraction 0.01 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$C', '$V', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), 'ms', d['parity']['mismatches'])"
done
done
