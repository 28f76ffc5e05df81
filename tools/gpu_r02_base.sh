#!/bin/bash
# GPU visit: parity tests, one bench line per BASELINE config (c1 with the reference arm), launch list of the c1 step.  usage: tools/gpu_r02_base.sh <tag>
TAG=${1:-r02a}; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi_$TAG.txt; nproc >> $O/smi_$TAG.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_$TAG.log; tail -n 4 $O/pytest_$TAG.log
timeout 600 bash tools/gpu_bench.sh $TAG "c1" ref
timeout 900 bash tools/gpu_bench.sh $TAG "c2 c3 c4 c5"
CMD="python bench.py --steps 1 --warmup 1 --batch 256 --streams 1 --no-cpu-baseline --no-e2e"
$CMD > $O/plain_$TAG.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu_l_$TAG.log 2>&1
echo done
