#!/bin/bash
# 8-GPU visit, second pass: C1 at N = 2, 4, 8 with the rank -> device spread, C5 at N = 2, 4, 8 with host-packed masks.  usage: tools/gpu_scale8b.sh <tag>
TAG=${1:-s8b}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2 4 8; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29610 + n)) bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c1_n${n}_$TAG.json 2> $O/bench_c1_n${n}_$TAG.err; echo "c1 n$n rc=$?"
  timeout 300 $TR --nproc-per-node $n --master-port $((29620 + n)) bench.py --workload c5 --gpus $n --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_c5_n${n}_$TAG.json 2> $O/bench_c5_n${n}_$TAG.err; echo "c5 n$n rc=$?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob('$O/bench_c*_n*_$TAG.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'value %.0f e2e %.0f' % (d['value'], d['e2e']['value']), d['e2e'].get('h2d_GBps_per_gpu'), d['config'].get('device_map'))
    except Exception as e: print(f, 'no line', e)
PY
