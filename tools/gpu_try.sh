#!/bin/bash
# short GPU visit while iterating on a kernel: extractor parity tests, then a device-only c1 bench line (no e2e / cpu legs).  usage: tools/gpu_try.sh <tag> [full]
TAG=${1:-t}; FULL=$2; O=gpurun_out; mkdir -p $O
if [ -n "$FULL" ]; then timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; else timeout 600 python -m pytest tests/test_gpu_extractor.py -m gpu -x -q > $O/pytest_$TAG.log 2>&1; fi
echo "pytest rc=$?"; tail -n 12 $O/pytest_$TAG.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_$TAG.log 2> $O/bench_$TAG.err; echo "bench rc=$?"; tail -n 3 $O/bench_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open('$O/bench_$TAG.log').read().strip().splitlines()[-1])
    print('value %.0f' % d['value']); print({k: round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()})
except Exception as e: print('no bench line', e)
PY
