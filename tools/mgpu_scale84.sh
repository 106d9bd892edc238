#!/bin/bash
# 8-GPU box: parity of the multi-GPU data paths + the headline workload at 8 and 4 GPUs (tile exchange).
TAG=${1:-r02}
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $RUN --nproc-per-node 8 --master-port 29511 tests/mgpu_check.py --big > gpurun_out/${TAG}_g8_parity.log 2>&1
grep "MGPU_" gpurun_out/${TAG}_g8_parity.log || tail -30 gpurun_out/${TAG}_g8_parity.log
for G in 8 4; do
    out=gpurun_out/${TAG}_scale_g${G}.json
    timeout 600 $RUN --nproc-per-node $G --master-port 2951$G bench.py --gpus $G --steps 40 --warmup 5 --no-cpu-baseline 2>gpurun_out/${TAG}_scale_g${G}.err | tail -1 > $out
    python - "$out" "$G" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    ph = {k.replace("_ms", ""): round(v * 1e3) for k, v in (d.get("phase_ms_sharded") or {}).items()}
    p = d.get("parity") or {}
    print(f"G={sys.argv[2]} value {d['value']:8.1f} steps/s  e2e {d['e2e']['value']:7.1f}  parity rel_err {p.get('rel_err')} its_equal {p.get('iterations_equal')} ranks_identical {p.get('ranks_bit_identical')}  phases(us) {ph}")
except Exception as e:
    print("G=" + sys.argv[2], "FAILED", e)
PY
done
