#!/bin/bash
# Final validation of a round on one GPU: the GPU test-suite, smoke(), the bench lines kept under profiles/.
TAG=${1:-r02}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/${TAG}_bench_reference_n2048.json 2>gpurun_out/${TAG}_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n2048.json 2>gpurun_out/${TAG}_bench.err
for n in 1024 512; do python bench.py --n $n --steps 200 --warmup 10 > gpurun_out/${TAG}_bench_n$n.json 2>>gpurun_out/${TAG}_bench.err; done
python bench.py --workload ensemble --members 8 --steps 200 --warmup 5 > gpurun_out/${TAG}_ensemble_g1.json 2>>gpurun_out/${TAG}_bench.err
python tools/run_c3.py --steps 3000 --steps-out 100 2>>gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_c3_g1.json
python - "$TAG" <<'PY'
import json, sys
t = sys.argv[1]
for f in ("bench_reference_n2048", "bench_n2048", "bench_n1024", "bench_n512", "ensemble_g1", "c3_g1"):
    try:
        d = json.load(open(f"gpurun_out/{t}_{f}.json"))
        extra = ""
        if "roofline" in d: extra += f" gemm1 frac {d['roofline']['frac']:.3f}"
        if "roofline_poisson" in d: extra += f" poisson frac {d['roofline_poisson']['frac']:.3f}"
        if d.get("parity"): extra += f" parity {d['parity'].get('rel_err')}"
        if d.get("cpu_baseline"): extra += f" cpu {d['cpu_baseline']['value']:.2f} ({d['cpu_baseline']['kind']})"
        print(f"{f}: value {d['value']:.1f} e2e {(d.get('e2e') or {}).get('value')}{extra}")
    except Exception as e:
        print(f, "FAILED", e)
PY
