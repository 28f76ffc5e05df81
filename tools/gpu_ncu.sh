#!/bin/bash
# one full ncu capture of selected kernels of a short bench run.  usage: tools/gpu_ncu.sh <tag> <kernel-regex> [count]
TAG=$1; KRE=$2; CNT=${3:-2}
O=gpurun_out; mkdir -p $O
CMD="python bench.py --steps 1 --warmup 1 --batch 256 --streams 1 --no-cpu-baseline --no-e2e"
$CMD > $O/plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:$KRE" -s $CNT -c $CNT -f -o $O/prof_$TAG $CMD > $O/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; tail -n 3 $O/ncu_$TAG.log | cut -c1-300
