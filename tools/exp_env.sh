#!/bin/bash
# usage: tools/exp_env.sh "ENV=1 ENV2=2" "..." : run the device-resident leg of bench.py under each environment
for e in "$@"; do
  echo -n "[$e] "
  env $e python bench.py --steps 200 --warmup 10 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('step_ms=%.4f kernel_ms=%.4f frac=%.3f loss=%.9g' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['loss']))"
done
