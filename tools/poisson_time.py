import sys, numpy as np, torch
sys.path.insert(0, ".")
import oracle
from quflow_b200._cuda import Handle
N = int(sys.argv[1])
h = Handle(N)
W = torch.from_numpy(oracle.random_skewherm(N, 1)).cuda()
P = torch.empty_like(W)
for _ in range(5): h.solve_poisson(W, out=P)
err = np.linalg.norm(P.cpu().numpy() - oracle.solve_poisson(W.cpu().numpy())) / np.linalg.norm(P.cpu().numpy())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for r in range(5):
    e0.record()
    for _ in range(20): h.solve_poisson(W, out=P)
    e1.record(); e1.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20)
print("N=%d %s  %.1f us  (%.1f%% of 6553.9 GB/s)  rel.err vs oracle %.1e" % (N, sys.argv[2], best * 1e3, 100 * 32.0 * N * N / (best * 1e-3) / 1e9 / 6553.9, err))
