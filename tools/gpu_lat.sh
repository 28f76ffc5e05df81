#!/bin/bash
# GPU visit: extractor parity tests + per-call latencies.  usage: tools/gpu_lat.sh <tag>
TAG=${1:-lat}; O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_extractor.py tests/test_host_dropin.py tests/test_gpu_frame.py -m gpu -x -q 2>&1 | tail -5
python tools/latency_probe.py > $O/latency_$TAG.txt 2>&1; cat $O/latency_$TAG.txt
ORBX_GRAPH=0 python tools/latency_probe.py 2>&1 | head -2
cat $O/host_dropin_timings.txt
