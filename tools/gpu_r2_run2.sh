#!/bin/bash
# round 2, GPU call 2: GPU tests, quick timings, bench, pyramid launch list + ncu capture of the last pyramid level
cd "$(dirname "$0")/.."
T=${1:-r2_02}
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_tests.log 2>&1
tail -4 gpurun_out/${T}_tests.log
timeout 600 python tools/gpu_quick.py ${T} > gpurun_out/${T}_quick.log 2>&1
tail -12 gpurun_out/${T}_quick.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 400 gpurun_out/${T}_bench.err
head -c 400 gpurun_out/${T}_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_pyramid_launches.csv python tools/gpu_pyramid_once.py rk > gpurun_out/${T}_pyr.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/${T}_pyr_l3 python tools/gpu_pyramid_once.py rk >> gpurun_out/${T}_pyr.log 2>&1
ls -la gpurun_out | grep ${T}
