#!/bin/bash
# GPU visit: bench lines of the given workloads (native arm, optionally the reference arm).  usage: tools/gpu_bench.sh <tag> "<workloads>" [ref]
TAG=$1; WL=${2:-c1}; REF=$3
O=gpurun_out; mkdir -p $O
for w in $WL; do
  python bench.py --workload $w --steps 10 --warmup 3 > $O/bench_${w}_$TAG.json 2> $O/bench_${w}_$TAG.err; echo "bench $w rc=$?"; tail -n 3 $O/bench_${w}_$TAG.err
  python - <<PY
import json
try:
    d=json.loads(open('$O/bench_${w}_$TAG.json').read().strip().splitlines()[-1])
    print('$w value %.0f %s e2e %.0f cpu %s' % (d['value'], d['unit'], d['e2e']['value'], d.get('cpu_baseline', {}).get('value')))
    print({k: round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()}, 'dom', d['roofline']['kernel'], 'frac %.3f' % d['roofline']['frac'])
except Exception as e: print('no bench line', e)
PY
  if [ -n "$REF" ]; then python bench.py --workload $w --impl reference --steps 3 --warmup 1 > $O/benchref_${w}_$TAG.json 2> $O/benchref_${w}_$TAG.err; echo "ref $w rc=$?"; cut -c1-300 $O/benchref_${w}_$TAG.json; fi
done
