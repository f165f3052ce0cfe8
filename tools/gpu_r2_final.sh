#!/bin/bash
# single-GPU call: tuning variants (optional), full GPU test suite, bench line + reference arm, ncu constants of the final kernel,
# launch lists and the pyramid's last-level capture
cd "$(dirname "$0")/.."
T=${1:-r2_final}
mkdir -p gpurun_out
for V in sb24 sb32 sb32sp12; do
  if [ -f bhusie_b200/lib/libbhray_${V}.so ]; then
    BHRAY_LIB=$PWD/bhusie_b200/lib/libbhray_${V}.so timeout 300 python tools/gpu_quick.py ${T}_${V} > gpurun_out/${T}_${V}_quick.log 2>&1
    echo "== ${V}"; grep -E "^c3_rk_fused|^pyramid" gpurun_out/${T}_${V}_quick.log
  fi
done
timeout 300 python tools/gpu_quick.py ${T} > gpurun_out/${T}_quick.log 2>&1
echo "== default"; grep -E "^c3_|^pyramid" gpurun_out/${T}_quick.log
( time timeout 2400 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_tests.log 2>&1
tail -4 gpurun_out/${T}_tests.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 300 gpurun_out/${T}_bench.err
head -c 300 gpurun_out/${T}_bench.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err
bash tools/capture_constants.sh ${T}_const > gpurun_out/${T}_const.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_bench_launches.csv python bench.py --steps 2 --warmup 1 --allow-short-warmup --no-extras --no-cpu-baseline > gpurun_out/${T}_bench_under_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_pyramid_launches.csv python tools/gpu_pyramid_once.py rk > gpurun_out/${T}_pyr.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/${T}_pyr_l3 python tools/gpu_pyramid_once.py rk >> gpurun_out/${T}_pyr.log 2>&1
ls -la gpurun_out | grep ${T} | head -40
