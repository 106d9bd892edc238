#!/bin/bash
# A/B of the multi-GPU data paths on ONE box (run through gpurun --gpus G):
#     tools/mgpu_ab.sh G [N] [steps]
# parity first (tests/mgpu_check.py: all modes must agree with the single-GPU run bit for bit across ranks), then one
# bench line per data path.  Results land in gpurun_out/mgpu_ab_g$G_$mode.json.
set -u
G=${1:-2}; N=${2:-2048}; STEPS=${3:-30}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
mkdir -p gpurun_out
$RUN --master-port 29511 tests/mgpu_check.py 2>&1 | grep "MGPU_\|rank 0" | tail -9
for mode in push pull pushcopy; do
    out=gpurun_out/mgpu_ab_g${G}_${mode}.json
    QF_COMM=$mode $RUN --master-port 29512 bench.py --gpus $G --n $N --steps $STEPS --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $out
    python - "$out" "$mode" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
ph = {k: round(v * 1e3) for k, v in (d.get("phase_ms_sharded") or {}).items()}
print(f"{sys.argv[2]:9s} value {d['value']:8.1f} steps/s   e2e {d['e2e']['value']:7.1f}   phases(us) {ph}")
PY
done
