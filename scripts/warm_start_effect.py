"""What warm-starting the duals across binary-search probes buys (cfg3-size graph): for a descending sequence of Z,
e_max of the running mean after nit iterations and the rounding remainder, cold against warm."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from sig_sdp_mmw_b200 import _lib, mmw
from sig_sdp_mmw_b200.topology import sparse_env
cs, rho = (63, 125e-4) if len(sys.argv) < 2 else (int(sys.argv[1]), float(sys.argv[2]))
state = sparse_env(cell_size=cs, sta_density_per_1m2=rho, seed=0).generate_S_Q_hmax()
K = state[0].shape[0]
plan = _lib.Plan(state, device=0, order=1)
for nit in (50, 150):
    prev = None
    for Z in (40, 32, 26, 22, 19, 17, 16):
        row = []
        for mode in ("cold", "warm"):
            sol = _lib.Solver(plan, Z, 2 * Z, 0.04)
            if mode == "warm" and prev is not None:
                sol.warm_start(prev)
            sol.iterate(nit, None, 1, None)
            torch.cuda.synchronize()
            row.append(sol.gap_prepare(None))
            if mode == "warm":
                prev = sol
        print("n=%d nit=%d Z=%d: e_max(X_avgd/i) cold %.4f warm %.4f" % (K, nit, Z, row[0], row[1]))
# through the drop-in object and the rounding: remainder per probe
for ws in (False, True):
    alg = mmw(nit=150, eta=0.04, omega="device", warm_start=ws, seed=1)
    out = []
    for Z in (22, 19, 17, 16, 15):
        _, gX = alg.run_with_state(0, Z, state)
        np.random.seed(0)
        z, _, rem = alg.rounding(Z, gX, state, nattempt=3)
        out.append((Z, rem))
    print("warm_start=%s (Z, remainder):" % ws, out)
