#!/bin/bash
# round 2, GPU call 1: full GPU test suite, bench line, pyramid launch list + ncu capture of the last pyramid level
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2_01_gpu.txt 2>&1
nproc >> gpurun_out/r2_01_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_01_tests.log 2>&1
tail -5 gpurun_out/r2_01_tests.log
timeout 900 python bench.py > gpurun_out/r2_01_bench.json 2> gpurun_out/r2_01_bench.err
tail -c 600 gpurun_out/r2_01_bench.err
head -c 1500 gpurun_out/r2_01_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_01_bench_ref.json 2>> gpurun_out/r2_01_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_01_pyramid_launches.csv python tools/gpu_pyramid_once.py rk > gpurun_out/r2_01_pyr.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2_01_pyr_l3 python tools/gpu_pyramid_once.py rk >> gpurun_out/r2_01_pyr.log 2>&1
ls -la gpurun_out | tail -12
