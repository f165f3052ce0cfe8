#!/bin/bash
cd "$(dirname "$0")/.."
T=${1:-r2_05}
mkdir -p gpurun_out
timeout 600 python tools/gpu_quick.py ${T} > gpurun_out/${T}_quick.log 2>&1
tail -22 gpurun_out/${T}_quick.log
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_tests.log 2>&1
tail -6 gpurun_out/${T}_tests.log
