#!/bin/bash
# GPU visit: parity tests only (+ optional parity report).  usage: tools/gpu_tests.sh <tag> [report-frames]
TAG=${1:-t}; FR=${2:-0}
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -n 15 $O/pytest_$TAG.log
if [ "$FR" != "0" ]; then python tools/parity_report.py --frames $FR > $O/parity_$TAG.md 2> $O/parity_$TAG.err; echo "parity rc=$?"; head -n 12 $O/parity_$TAG.md | cut -c1-400; fi
