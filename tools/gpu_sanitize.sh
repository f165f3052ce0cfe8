#!/bin/bash
# compute-sanitizer (memcheck + racecheck + initcheck) over a small pass of every kernel: gpurun_out/sanitizer_*.log
mkdir -p gpurun_out
cat > /tmp/san_run.py <<'PY'
import sys
sys.path.insert(0, ".")
from bhusie_b200 import assets, pipelines as P, uniforms as U
tex = assets.small_textures()
blob, _ = P.model_from_arrays(*assets.uv_sphere(12, 16, radius=4.0))
for mode in (P.NUMERIC_LITERAL, P.NUMERIC_FUSED):
    ctx = P.Context(0, numeric_mode=mode)
    ctx.set_textures(tex); ctx.upload_models(blob)
    for method in (0, 1):
        det = U.RayDetails(integration_method=method, model_count=1, angle_division_threshold=0.08)
        pyr = P.RayPyramid(ctx, base=(16, 9), iters=3, aux=7)
        pyr.pass_(U.Camera(), U.BlackHole(), det)
        pyr.levels[-1].read(); pyr.sky.read()
        t = P.RayPipeline(ctx, 40, 22, aux=7); t.set_tiling(4, 1, 3); t.pass_(U.Camera(position=(0, 0, -45)), U.BlackHole(), det); t.read()
        pyr.close(); t.close()
    ctx.close()
print("sanitizer workload done")
PY
for tool in memcheck racecheck initcheck; do
  compute-sanitizer --tool $tool --error-exitcode 7 python /tmp/san_run.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$tool.log | tail -1)"
done
