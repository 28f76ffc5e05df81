#!/bin/bash
# 8-GPU visit: H2D contention matrix (with and without result traffic), C1 at N = 8, C5 at N = 1, 2, 4, 8.  usage: tools/gpu_scale8.sh <tag>
TAG=${1:-s8}; O=gpurun_out; mkdir -p $O
nvidia-smi topo -m > $O/topo_$TAG.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 8 --master-port 29511 tools/h2d_matrix.py > $O/h2d_matrix_$TAG.jsonl 2> $O/h2d_matrix_$TAG.err; echo "matrix rc=$?"
timeout 200 $TR --nproc-per-node 8 --master-port 29512 tools/h2d_matrix.py --d2h > $O/h2d_matrix_d2h_$TAG.jsonl 2>> $O/h2d_matrix_$TAG.err; echo "matrix d2h rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c1_n8_$TAG.json 2> $O/bench_c1_n8_$TAG.err; echo "c1 n8 rc=$?"
for n in 1 2 4 8; do
  if [ $n = 1 ]; then timeout 300 python bench.py --workload c5 --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_c5_n${n}_$TAG.json 2> $O/bench_c5_n${n}_$TAG.err
  else timeout 300 $TR --nproc-per-node $n --master-port $((29520 + n)) bench.py --workload c5 --gpus $n --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_c5_n${n}_$TAG.json 2> $O/bench_c5_n${n}_$TAG.err; fi
  echo "c5 n$n rc=$?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob('$O/bench_c*_n*_$TAG.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'value %.0f e2e %.0f' % (d['value'], d['e2e']['value']), d['e2e'].get('h2d_GBps_per_gpu'))
    except Exception as e: print(f, 'no line', e)
PY
tail -n 16 $O/h2d_matrix_$TAG.jsonl
