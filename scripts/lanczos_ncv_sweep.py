"""Final-factor eigen-solve (top-r |lambda| of X_avgd) at several Krylov basis sizes:
mat-vecs, restarts and device time.  cfg4 by default."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import bench
import torch
from sig_sdp_mmw_b200.mmw import mmw
from sig_sdp_mmw_b200.lanczos import thick_restart_lanczos

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg4_100k"
state, Z, rr, dtype = bench.make_state(wl, 0)
alg = mmw(nit=150, eta=bench.ETA, rank_radio=rr, dtype=dtype, omega="device", device=0, order=1, seed=1)
alg.run_with_state(0, Z, state)
sol = alg.last_solver
dev = torch.device("cuda", 0)
K = sol.plan.n
rank = int(min(K - 1, (Z - 1) * rr))
stream = torch.cuda.current_stream().cuda_stream
sol.xavg_matrix(1.0 / 150, stream)
g = torch.Generator(device="cpu").manual_seed(2)
v0 = torch.randn(K, dtype=torch.float64, generator=g).to(dev)
ref = None
for ncv in [None, 80, 120, 140, 170, 200]:
    for rep in range(2):
        torch.cuda.synchronize(); t = time.perf_counter()
        lam, V, info = thick_restart_lanczos(alg._matmat(sol, torch, dev), K, rank, "LM", v0, ncv=ncv, tol=alg.eig_tol,
                                             native_steps=alg._native_steps(sol, torch))
        torch.cuda.synchronize(); dt = time.perf_counter() - t
    lam = lam.cpu().numpy()
    if ref is None:
        ref = lam
    print(f"ncv {info['ncv']:4d} k {rank} matvecs {info['matvecs']:5d} restarts {info['restarts']:3d} "
          f"time {dt * 1e3:7.1f} ms  max|dlam| {np.abs(lam - ref).max():.2e}", flush=True)
