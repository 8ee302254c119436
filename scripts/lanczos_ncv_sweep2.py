"""Final factor at cfg4: time of thick_restart_lanczos for (ncv, keep_extra) pairs on the 20- and 150-iteration matrices."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from sig_sdp_mmw_b200 import _lib, mmw
from bench import make_state, ETA
from sig_sdp_mmw_b200.lanczos import thick_restart_lanczos
state, Z, rr, dtype = make_state("cfg4_100k", 0)
plan = _lib.Plan(state, device=0, order=1)
sol = _lib.Solver(plan, Z, Z * rr, ETA)
dev = torch.device("cuda", 0)
done = 0
for nit in (20, 150):
    sol.iterate(nit - done, None, 1, None); done = nit
    torch.cuda.synchronize()
    sol.xavg_matrix(1.0 / nit, None)
    alg = mmw()
    mm = alg._matmat(sol, torch, dev)
    v0 = torch.randn(plan.n, dtype=torch.float64, generator=torch.Generator().manual_seed(1)).to(dev)
    for ncv, ke in ((100, None), (100, 8), (90, 8), (80, 8), (80, 12), (70, 8), (64, 6)):
        best = 1e9
        for rep in range(2):
            torch.cuda.synchronize(); t = time.perf_counter()
            lam, V, info = thick_restart_lanczos(mm, plan.n, 30, "LM", v0, ncv=ncv, tol=1e-10, native_steps=mmw._native_steps(sol, torch), keep_extra=ke)
            torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
        print("nit=%d ncv=%d keep_extra=%s: %.1f ms matvecs %d restarts %d converged %s lam30 %.10f" % (nit, ncv, ke, best * 1e3, info["matvecs"], info["restarts"], info["converged"], float(lam.abs().min())))
