import sys; sys.path.insert(0, '.')
import numpy as np, torch, oracle, quflow_b200 as qf
from quflow_b200._cuda import get_handle
for N in (64, 256, 300, 1024, 2048):
    W = oracle.random_skewherm(N, 1)
    P = qf.solve_poisson(torch.from_numpy(W).cuda()).cpu().numpy()
    Pr = oracle.solve_poisson(W)
    print(N, "rel err", np.linalg.norm(P - Pr) / np.linalg.norm(Pr), flush=True)
