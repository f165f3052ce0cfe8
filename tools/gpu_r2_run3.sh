#!/bin/bash
# GPU call 3: quick timings first, then the GPU tests, then profiles
cd "$(dirname "$0")/.."
T=${1:-r2_03}
mkdir -p gpurun_out
timeout 600 python tools/gpu_quick.py ${T} > gpurun_out/${T}_quick.log 2>&1
tail -22 gpurun_out/${T}_quick.log
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_tests.log 2>&1
tail -6 gpurun_out/${T}_tests.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_pyramid_launches.csv python tools/gpu_pyramid_once.py rk > gpurun_out/${T}_pyr.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/${T}_pyr_l3 python tools/gpu_pyramid_once.py rk >> gpurun_out/${T}_pyr.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/${T}_c3 python tools/gpu_run_cam.py 3840 2160 -19 2 >> gpurun_out/${T}_pyr.log 2>&1
ls -la gpurun_out | grep ${T}
