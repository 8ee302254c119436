"""Where the end-to-end time of one solve goes (cfg4 by default): run with SIGSDP_PLAN_TIMING=1
to get the library's own stage times on stderr next to the Python-side laps printed here."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import bench
from sig_sdp_mmw_b200.mmw import mmw
import torch

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg4_100k"
state, Z, rr, dtype = bench.make_state(wl, 0)
mmw(nit=3, eta=bench.ETA, rank_radio=rr, dtype=dtype, omega="device", device=0, order=1, seed=1).run_with_state(0, Z, state)
torch.cuda.synchronize()
for rep in range(2):
    alg = mmw(nit=150, eta=bench.ETA, rank_radio=rr, dtype=dtype, omega="device", device=0, order=1, seed=1)
    print(f"---- solve {rep}", file=sys.stderr)
    t0 = time.perf_counter()
    ok, X_half = alg.run_with_state(0, Z, state)
    t1 = time.perf_counter()
    print(f"e2e {1e3 * (t1 - t0):.1f} ms; eig info {alg.last_eig_info}", file=sys.stderr)

# Python-side laps of the set-up (what LOGGED_NP_DATA["mmw_state_process"] covers)
from sig_sdp_mmw_b200 import _lib
from sig_sdp_mmw_b200 import sdp_solver as _sd
import importlib
_sdm = importlib.import_module("sig_sdp_mmw_b200.sdp_solver")
for rep in range(2):
    t0 = time.perf_counter(); _sdm._digest(state[0]); _sdm._digest(state[1]); _sdm._digest_vec(state[2]); t1 = time.perf_counter()
    plan = _lib.Plan(state, device=0, order=1); t2 = time.perf_counter()
    sol = _lib.Solver(plan, Z, Z * rr, bench.ETA, _lib.F64 if dtype == "float64" else _lib.F32); t3 = time.perf_counter()
    print(f"python laps: digest {1e3*(t1-t0):.1f} ms, Plan() {1e3*(t2-t1):.1f} ms, Solver() {1e3*(t3-t2):.1f} ms", file=sys.stderr)
    del sol, plan
