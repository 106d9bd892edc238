#!/bin/bash
# Timing experiments of the tile-exchange path on one box: per-phase device times (profile_iteration, rank 0) under
# different settings.   gpurun --gpus G -- tools/mgpu_probe.sh G [tag]
G=${1:-2}; TAG=${2:-probe}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
i=0
for cfg in "QF_XCHG_PUSH=sm" "QF_XCHG_PUSH=inline" "QF_XCHG_PUSH=inline QF_XCHG_UPPER=0"; do
    i=$((i+1))
    out=gpurun_out/${TAG}_g${G}_$i.json
    env $cfg timeout 600 $RUN --master-port 2952$i bench.py --gpus $G --steps 40 --warmup 5 --no-cpu-baseline 2>gpurun_out/${TAG}_g${G}_$i.err | tail -1 > $out
    python - "$out" "$cfg" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    ph = {k.replace("_ms", ""): round(v * 1e3) for k, v in (d.get("phase_ms_sharded") or {}).items()}
    print(f"{sys.argv[2]:45s} value {d['value']:7.1f}  phases(us) {ph}")
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
