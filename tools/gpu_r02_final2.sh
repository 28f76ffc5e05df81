#!/bin/bash
# final GPU visit of the round (r02j): full parity suite, smoke, all five configs (C1 with the reference arm), launch list, single-frame timings.  usage: tools/gpu_r02_final2.sh <tag>
TAG=${1:-r02j}; O=gpurun_out; mkdir -p $O
timeout 1300 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_$TAG.log; tail -n 3 $O/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 700 bash tools/gpu_bench.sh $TAG "c1" ref
timeout 900 bash tools/gpu_bench.sh $TAG "c2 c3 c4 c5"
python tools/lat1_probe.py 2>&1 | grep -v Warn | tee $O/lat1_$TAG.txt
python tools/qt_stamps_probe.py > $O/qt_stamps_$TAG.txt 2>&1; tail -n 16 $O/qt_stamps_$TAG.txt
cat $O/host_dropin_timings.txt
timeout 600 python tools/parity_report.py --frames 60 > $O/parity_$TAG.md 2> $O/parity_$TAG.err; echo "parity rc=$?"; grep -c "| 0 |" $O/parity_$TAG.md
CMD="python bench.py --steps 1 --warmup 1 --batch 256 --streams 1 --no-cpu-baseline --no-e2e"
$CMD > $O/plain_$TAG.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu_l_$TAG.log 2>&1
echo done
