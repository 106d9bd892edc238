#!/bin/bash
# Small-N experiments on one GPU: Poisson with 8 positions per thread, plain instead of cooperative GEMM launch.
TAG=${1:-r02l}
for n in 512 1024; do
  i=0
  for cfg in "QF_NONE=1" "QF_POISSON_L=8" "QF_GEMM_COOP=0"; do
    i=$((i+1))
    out=gpurun_out/${TAG}_n${n}_$i.json
    env $cfg python bench.py --n $n --steps 200 --warmup 10 --no-cpu-baseline > $out 2>gpurun_out/${TAG}.err
    python - "$out" "$n" "$cfg" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("N=%s %-18s value %.1f e2e %.1f gemm1 frac %.3f poisson frac %.3f phases(us) %s" % (sys.argv[2], sys.argv[3], d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline_poisson"]["frac"], {k.replace("_ms",""): round(v * 1e3, 1) for k, v in d["phase_ms"].items() if not k.startswith("x_")}))
except Exception as e:
    print(sys.argv[2], sys.argv[3], "FAILED", e)
PY
  done
done
python tools/run_c3.py --steps 3000 --steps-out 100 2>gpurun_out/${TAG}_c3.err | tail -1 | tee gpurun_out/${TAG}_c3_g1.json
