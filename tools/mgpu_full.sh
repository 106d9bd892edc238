#!/bin/bash
# One 8-GPU box: parity of every multi-GPU data path, the strong-scaling curve of the headline workload on the same box
# (8, 4 and 2 GPUs, tile exchange), and BASELINE config 5 (ensemble, 8 members per GPU).
#     gpurun --gpus 8 -- tools/mgpu_full.sh [tag]
set -u
TAG=${1:-r02}
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 900 $RUN --nproc-per-node 8 --master-port 29511 tests/mgpu_check.py --big > gpurun_out/${TAG}_g8_parity.log 2>&1
grep "MGPU_\|rank 0" gpurun_out/${TAG}_g8_parity.log | cut -c1-260 | tail -12
grep -q MGPU_OK gpurun_out/${TAG}_g8_parity.log || tail -30 gpurun_out/${TAG}_g8_parity.log
for G in 8 4 2; do
    out=gpurun_out/${TAG}_scale_g${G}.json
    timeout 600 $RUN --nproc-per-node $G --master-port 2951$G bench.py --gpus $G --steps 40 --warmup 5 --no-cpu-baseline 2>gpurun_out/${TAG}_scale_g${G}.err | tail -1 > $out
    python - "$out" "$G" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    ph = {k: round(v * 1e3) for k, v in (d.get("phase_ms_sharded") or {}).items()}
    p = d.get("parity") or {}
    print(f"G={sys.argv[2]} value {d['value']:8.1f} steps/s  e2e {d['e2e']['value']:7.1f}  parity rel_err {p.get('rel_err')} its_equal {p.get('iterations_equal')} ranks_identical {p.get('ranks_bit_identical')}  phases(us) {ph}")
except Exception as e:
    print("G=" + sys.argv[2], "FAILED", e)
PY
done
out=gpurun_out/${TAG}_ensemble_g8.json
timeout 600 $RUN --nproc-per-node 8 --master-port 29519 bench.py --workload ensemble --gpus 8 --members 8 --steps 200 --warmup 5 2>gpurun_out/${TAG}_ensemble_g8.err | tail -1 > $out
python - "$out" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(f"ensemble 8x8 N=256: {d['value']:.0f} member-steps/s  e2e {d['e2e']['value']:.0f}  parity {d['parity']}  phases(us) { {k: round(v*1e3,1) for k,v in d['phase_ms'].items()} }  gemm1 frac {d['roofline']['frac']:.3f}")
except Exception as e:
    print("ensemble FAILED", e)
PY
