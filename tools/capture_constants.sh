#!/bin/bash
# On the GPU box: one `ncu --set full` capture of the headline kernel (trace_kernel of bench.py's workload) per numeric mode, with
# the hash of the kernel source it was built from.  Back in the container:
#     python tools/update_constants.py gpurun_out/<tag>_const
# turns them into profiles/trace_kernel_dram.json (bench.py refuses constants whose source hash is not the build's).
cd "$(dirname "$0")/.."
T=${1:-r2_const}
mkdir -p gpurun_out
python -m bhusie_b200.build --hash > gpurun_out/${T}.srchash
for MODE in fused literal; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel --launch-skip 1 --launch-count 1 -f \
      -o gpurun_out/${T}_${MODE} python bench.py --steps 1 --warmup 1 --allow-short-warmup --no-cpu-baseline --no-extras --numeric-mode ${MODE} \
      > gpurun_out/${T}_${MODE}.log 2>&1
done
ls -la gpurun_out | grep ${T}
