#!/bin/bash
# Parity + A/B of the multi-GPU data paths on ONE box (run through gpurun --gpus G):
#     tools/mgpu_ab.sh G [N] [steps] [tag]
# parity first (tests/mgpu_check.py --big: every data path against the single-GPU run, the oracle, bit-identical ranks),
# then one bench line per data path.  Results land in gpurun_out/<tag>_g$G_$mode.json, the parity log in
# gpurun_out/<tag>_g$G_parity.log.
set -u
G=${1:-2}; N=${2:-2048}; STEPS=${3:-30}; TAG=${4:-mgpu_ab}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 600 $RUN --master-port 29511 tests/mgpu_check.py --big > gpurun_out/${TAG}_g${G}_parity.log 2>&1
grep "MGPU_\|rank 0" gpurun_out/${TAG}_g${G}_parity.log | tail -12
grep -q MGPU_OK gpurun_out/${TAG}_g${G}_parity.log || tail -30 gpurun_out/${TAG}_g${G}_parity.log
for mode in tile pull; do
    out=gpurun_out/${TAG}_g${G}_${mode}.json
    QF_COMM=$mode timeout 600 $RUN --master-port 29512 bench.py --gpus $G --size $N --steps $STEPS --warmup 3 --no-cpu-baseline 2>gpurun_out/${TAG}_g${G}_${mode}.err | tail -1 > $out
    python - "$out" "$mode" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    ph = {k: round(v * 1e3) for k, v in (d.get("phase_ms_sharded") or {}).items()}
    print(f"{sys.argv[2]:6s} value {d['value']:8.1f} steps/s   e2e {d['e2e']['value']:7.1f}   parity {d.get('parity')}   phases(us) {ph}")
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
