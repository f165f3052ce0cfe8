#!/bin/bash
# 1/2/4/8-GPU strong scaling of the 4K headline workload + the 8K C4 config; writes gpurun_out/scale_*.json
# usage (on a box with >= NMAX GPUs): tools/run_scaling.sh [NMAX]
NMAX=${1:-8}
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/scale_4k_n1.json
for n in 2 4 8; do
  [ $n -le $NMAX ] || continue
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29510+n)) \
      bench.py --gpus $n --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/scale_4k_n$n.json
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NMAX --master-addr 127.0.0.1 --master-port 29529 \
    bench.py --gpus $NMAX --steps 10 --warmup 3 --width 7680 --height 4320 2>&1 | tail -1 > gpurun_out/c4_8k_n$NMAX.json
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/scale_4k_n*.json")) + sorted(glob.glob("gpurun_out/c4_8k_n*.json")):
    try:
        d = json.loads(open(f).read())
        print(f, "N=%d" % d["n_gpus"], "value %.0f" % d["value"], "ms %.3f" % d["ms_per_step"], "fps %.1f" % d["fps"],
              "e2e %.0f" % d["e2e"]["value"], d["clocks"])
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-400:])
PY
