#!/bin/bash
# GPU visit: quadtree forms -- parity of the stage, extractor tests, phase stamps, single-frame stage times.  usage: tools/gpu_qt.sh <tag>
TAG=${1:-qt}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_extractor.py -m gpu -x -q -k "quadtree or oracle_configs or every_kernel_form or golden or getters or batch" 2>&1 | tail -15
timeout 120 python tools/qt_stamps_probe.py > $O/qt_stamps_$TAG.txt 2>&1; cat $O/qt_stamps_$TAG.txt
timeout 300 python tools/lat1_probe.py > $O/lat1_$TAG.txt 2>&1; cat $O/lat1_$TAG.txt
timeout 600 python -m pytest tests/test_host_dropin.py -m gpu -x -q 2>&1 | tail -3; cat $O/host_dropin_timings.txt
