#!/bin/bash
# Same-box A/B of the fused GEMM-2 tail (QF_FUSE_POST=1, default) against the separate k_post launch (=0).
#     gpurun -- tools/gpu_ab_fuse.sh [tag]
TAG=${1:-r02b}
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for n in 512 1024 2048; do for f in 1 0; do
  QF_FUSE_POST=$f python bench.py --n $n --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_n${n}_f$f.json 2>gpurun_out/${TAG}.err
  python - "$n" "$f" "gpurun_out/${TAG}_n${n}_f$f.json" <<'PY'
import json, sys
n, f, path = sys.argv[1:]
try:
    d = json.load(open(path))
    print("N=%s fuse=%s value %.1f e2e %.1f phases(us) %s" % (n, f, d["value"], d["e2e"]["value"], {k: round(v * 1e3, 1) for k, v in d["phase_ms"].items()}))
except Exception as e:
    print("N=%s fuse=%s FAILED %s" % (n, f, e)); print(open("gpurun_out/%s.err" % path.split("/")[1].split("_n")[0]).read()[-1500:])
PY
done; done
