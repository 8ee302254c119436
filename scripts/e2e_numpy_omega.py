"""End-to-end time of the PARITY mode an unchanged driver gets: mmw(nit=K) with omega="numpy" (one np.random.randn(K, D)
per iteration on the host, as the reference draws them, mmw.py:226) next to omega="device", cfg4, K = 20."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import bench
import torch
from sig_sdp_mmw_b200.mmw import mmw

state, Z, rr, dtype = bench.make_state("cfg4_100k", 0)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
mmw(nit=3, eta=bench.ETA, rank_radio=rr, dtype=dtype, omega="device", device=0, order=1, seed=1).run_with_state(0, Z, state)
for omega in ("numpy", "device", "numpy"):
    alg = mmw(nit=K, eta=bench.ETA, rank_radio=rr, dtype=dtype, omega=omega, device=0, order=1, seed=1)
    np.random.seed(0)
    t0 = time.perf_counter()
    ok, X_half = alg.run_with_state(0, Z, state)
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    t1 = time.perf_counter(); np.random.randn(state[0].shape[0], Z * rr); draw = time.perf_counter() - t1
    print("omega=%s K=%d: %.1f ms end to end (%.1f iterations/s); one randn(n, D) draw on this host: %.1f ms"
          % (omega, K, 1e3 * t, K / t, 1e3 * draw))
