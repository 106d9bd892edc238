import sys; sys.path.insert(0,'.')
import numpy as np, oracle, quflow_b200 as qf
from quflow_b200._cuda import Handle
for N,G in ((256,2),(256,4),(512,8),(384,3)):
    W0 = oracle.random_skewherm(N, 77); dt = 0.25*qf.hbar(N)
    h1, hG = Handle(N), Handle(N); hG.set_emulated_ranks(G)
    Wa, Wb = W0.copy(), W0.copy()
    ra, ia = h1.isomp(Wa, dt, 15, want_iters=True); rb, ib = hG.isomp(Wb, dt, 15, want_iters=True)
    Wr = oracle.isomp(W0.copy(), dt, 15)
    print(N, G, list(ia[0])==list(ib[0]), np.linalg.norm(Wa-Wb)/np.linalg.norm(Wa), np.linalg.norm(Wa-Wr)/np.linalg.norm(Wr), np.linalg.norm(Wb-Wr)/np.linalg.norm(Wr))
    # one-step check
    Wa, Wb = W0.copy(), W0.copy()
    h1.isomp(Wa, dt, 1); hG.isomp(Wb, dt, 1)
    print("   1 step:", np.linalg.norm(Wa-Wb)/np.linalg.norm(Wa), np.abs(Wa-Wb).max())
