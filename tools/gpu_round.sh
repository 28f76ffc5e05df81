#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), launch list, one full ncu capture of the hot kernels.
# usage: tools/gpu_round.sh <tag> [ncu-kernel-regex]
TAG=${1:-rXX}; KRE=${2:-'k_fast|k_gauss7|k_orient|k_pyr'}
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_$TAG.log
python bench.py --steps 10 --warmup 3 > $O/bench_$TAG.log 2> $O/bench_$TAG.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/benchref_$TAG.log 2> $O/benchref_$TAG.err; echo "benchref rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --batch 256 --streams 1 --no-cpu-baseline --no-e2e"
$CMD > $O/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu_l_$TAG.log 2>&1
$CMD > $O/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -s 36 -c 12 -f -o $O/prof_$TAG $CMD > $O/ncu_f_$TAG.log 2>&1
echo done
