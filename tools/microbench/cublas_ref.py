# Library comparators on the box: cuBLAS DGEMM / ZGEMM via torch.matmul.
import torch, json, sys
dev = torch.device("cuda:0")
res = {}
def bench(fn, n=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
for N in (256, 512, 1024, 2048, 4096, 8192):
    A = torch.randn(N, N, dtype=torch.float64, device=dev); B = torch.randn(N, N, dtype=torch.float64, device=dev)
    C = torch.empty_like(A)
    ms = bench(lambda: torch.matmul(A, B, out=C))
    res[f"dgemm_{N}"] = {"ms": ms, "tflops": 2*N**3/ms/1e9}
    if N <= 4096:
        Az = torch.randn(N, N, dtype=torch.complex128, device=dev); Bz = torch.randn(N, N, dtype=torch.complex128, device=dev)
        Cz = torch.empty_like(Az)
        ms = bench(lambda: torch.matmul(Az, Bz, out=Cz))
        res[f"zgemm_{N}"] = {"ms": ms, "tflops": 8*N**3/ms/1e9}
    print(N, res.get(f"dgemm_{N}"), res.get(f"zgemm_{N}"), flush=True)
# sustained DGEMM 8192 for 3 s
N=8192
A = torch.randn(N, N, dtype=torch.float64, device=dev); B = torch.randn(N, N, dtype=torch.float64, device=dev); C = torch.empty_like(A)
import time
torch.cuda.synchronize(); t0=time.time(); n=0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
while time.time()-t0 < 3.0:
    torch.matmul(A, B, out=C); n+=1
    if n % 4 == 0: torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
res["dgemm_8192_sustained"] = {"tflops": n*2*N**3/e0.elapsed_time(e1)/1e9, "n": n}
print(res["dgemm_8192_sustained"])
json.dump(res, open("gpurun_out/cublas_ref.json", "w"), indent=1)
