#!/bin/bash
# GPU visit: full parity suite, one bench line per BASELINE config, matcher ncu capture (C2), launch list + full capture of the C1 step.
# usage: tools/gpu_r02_final.sh <tag>
TAG=${1:-r02e}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_$TAG.log; tail -n 3 $O/pytest_$TAG.log
timeout 600 bash tools/gpu_bench.sh $TAG "c1" ref
timeout 900 bash tools/gpu_bench.sh $TAG "c2 c3 c4 c5"
CMD2="python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD2 > $O/plainc2_$TAG.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_window_search|k_resolve_init" -s 6 -c 6 -f -o $O/prof_c2_$TAG $CMD2 > $O/ncu_c2_$TAG.log 2>&1
echo "ncu c2 rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --batch 256 --streams 1 --no-cpu-baseline --no-e2e"
$CMD > $O/plain_$TAG.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu_l_$TAG.log 2>&1
$CMD > $O/plain2_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -s 36 -c 12 -f -o $O/prof_$TAG $CMD > $O/ncu_f_$TAG.log 2>&1
echo done
