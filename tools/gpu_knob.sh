#!/bin/bash
# device-only c1 bench lines under different environment knobs.  usage: tools/gpu_knob.sh "<VAR=val ...>" "<VAR=val ...>" ...
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.0f' % d['value'], {k: round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()})"
done
