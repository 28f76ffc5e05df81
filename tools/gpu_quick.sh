#!/bin/bash
# quick GPU visit: parity tests + one bench line (no ncu).  usage: tools/gpu_quick.sh <tag> [pytest-args]
TAG=${1:-q}; shift
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q "$@" > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -n 15 $O/pytest_$TAG.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$TAG.log 2> $O/bench_$TAG.err; echo "bench rc=$?"; tail -n 5 $O/bench_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open('$O/bench_$TAG.log').read().strip().splitlines()[-1])
    print('value %.0f e2e %.0f' % (d['value'], d['e2e']['value'])); print({k: round(v,3) for k,v in d['roofline']['stage_ms_per_step'].items()})
except Exception as e: print('no bench line', e)
PY
python - <<PY
import torch, time
a = torch.empty(256*1024*1024, dtype=torch.uint8).pin_memory(); d = torch.empty_like(a, device='cuda')
for _ in range(2): d.copy_(a, non_blocking=True)
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(5): d.copy_(a, non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print('pinned H2D GB/s %.1f' % (5*a.numel()/dt/1e9))
t=time.perf_counter()
for _ in range(5): a.copy_(d, non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print('pinned D2H GB/s %.1f' % (5*a.numel()/dt/1e9))
PY
