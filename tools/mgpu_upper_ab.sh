#!/bin/bash
# A/B of the W~ exchange on one box: upper-only + local mirror (QF_XCHG_UPPER=1) against both triangles (=0).
#     gpurun --gpus G -- tools/mgpu_upper_ab.sh G [tag]
G=${1:-2}; TAG=${2:-r02f}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
QF_XCHG_UPPER=1 timeout 600 $RUN --master-port 29511 tests/mgpu_check.py > gpurun_out/${TAG}_g${G}_parity_upper.log 2>&1
grep "MGPU_" gpurun_out/${TAG}_g${G}_parity_upper.log || tail -20 gpurun_out/${TAG}_g${G}_parity_upper.log
QF_XCHG_PUSH=ce QF_XCHG_UPPER=1 timeout 600 $RUN --master-port 29513 tests/mgpu_check.py > gpurun_out/${TAG}_g${G}_parity_ce.log 2>&1
grep "MGPU_" gpurun_out/${TAG}_g${G}_parity_ce.log || tail -20 gpurun_out/${TAG}_g${G}_parity_ce.log
for up in 0 1 ce0 ce1; do
    out=gpurun_out/${TAG}_g${G}_upper$up.json
    case $up in ce*) export QF_XCHG_PUSH=ce;; *) export QF_XCHG_PUSH=sm;; esac
    QF_XCHG_UPPER=${up#ce} timeout 600 $RUN --master-port 29515 bench.py --gpus $G --steps 40 --warmup 5 --no-cpu-baseline 2>gpurun_out/${TAG}_g${G}_upper$up.err | tail -1 > $out
    python - "$out" "$up" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    ph = {k: round(v * 1e3) for k, v in (d.get("phase_ms_sharded") or {}).items()}
    p = d.get("parity") or {}
    print(f"upper={sys.argv[2]} value {d['value']:8.1f} steps/s  e2e {d['e2e']['value']:7.1f}  parity rel_err {p.get('rel_err')} its_equal {p.get('iterations_equal')} ranks_identical {p.get('ranks_bit_identical')}  phases(us) {ph}")
except Exception as e:
    print("upper=" + sys.argv[2], "FAILED", e)
PY
done
