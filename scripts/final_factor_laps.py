"""Where the time of mmw._final_factor goes (cfg4 by default): the statements of the method, timed one by one with a
device synchronisation after each, for nit = 20 and nit = 150."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import bench
import torch
from sig_sdp_mmw_b200 import _lib
from sig_sdp_mmw_b200.mmw import mmw
from sig_sdp_mmw_b200.lanczos import chebyshev_filtered_lanczos, thick_restart_lanczos

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg4_100k"
state, Z, rr, dtype = bench.make_state(wl, 0)
dev = torch.device("cuda", 0)
mmw(nit=3, eta=bench.ETA, rank_radio=rr, dtype=dtype, omega="device", device=0, order=1, seed=1).run_with_state(0, Z, state)
for nit in (20, 150, 150):
    alg = mmw(nit=nit, eta=bench.ETA, rank_radio=rr, dtype=dtype, omega="device", device=0, order=1, seed=1)
    plan = alg._plan_for(state)
    sol = _lib.Solver(plan, Z, Z * rr, bench.ETA, _lib.F64 if dtype == "float64" else _lib.F32)
    sol.iterate(nit, None, 1, None)
    torch.cuda.synchronize()
    laps = []
    def lap(name, t0):
        torch.cuda.synchronize()
        laps.append((name, 1e3 * (time.perf_counter() - t0)))
        return time.perf_counter()
    t_all = t = time.perf_counter()
    K = plan.n
    rank = int(min(K - 1, (Z - 1) * rr))
    sol.xavg_matrix(1.0 / nit, None); t = lap("xavg_matrix", t)
    g = torch.Generator(device=dev).manual_seed(2)
    v0 = torch.randn(K, dtype=torch.float64, generator=g, device=dev); t = lap("v0", t)
    perm = torch.from_numpy(plan.perm()).to(dev).long(); t = lap("perm", t)
    out = chebyshev_filtered_lanczos(alg._matmat(sol, torch, dev), K, rank, v0[perm], alg._native_steps(sol, torch),
                                     sol.lanczos_filter, tol=alg.eig_tol, degree=alg.eig_filter_degree); t = lap("filtered eig", t)
    lam, V, info = out
    X_half_int = V * torch.sqrt(lam.abs())[None, :]
    X_half = torch.empty_like(X_half_int)
    X_half[perm] = X_half_int; t = lap("scale + unpermute", t)
    sv = lam.abs().cpu().numpy(); t = lap("singular values", t)
    host = torch.empty(X_half.shape, dtype=X_half.dtype, pin_memory=True); t = lap("pinned alloc", t)
    host.copy_(X_half, non_blocking=True); t = lap("d2h", t)
    print("nit", nit, "total %.1f ms" % (1e3 * (time.perf_counter() - t_all)), " ".join("%s %.2f" % l for l in laps), file=sys.stderr)
    print("   eig info", {k: v for k, v in info.items() if k in ("cycle_s", "restart_s", "total_s", "lanczos_steps", "matvecs")}, file=sys.stderr)
    # the whole method through the object, for comparison
    t0 = time.perf_counter()
    alg._final_factor(sol, Z, nit, torch, dev, None)
    torch.cuda.synchronize()
    print("   _final_factor() %.1f ms" % (1e3 * (time.perf_counter() - t0)), file=sys.stderr)
    del sol
