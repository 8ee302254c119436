"""Dumps X_avgd / nit of the bench workload after 20 and 150 iterations (pattern + values, internal
numbering) to gpurun_out/xavg_cfg4.npz, for eigen-solver experiments off the GPU."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sig_sdp_mmw_b200 import _lib
from bench import make_state, ETA
state, Z, rr, dtype = make_state("cfg4_100k", 0)
plan = _lib.Plan(state, device=0, order=1)
sol = _lib.Solver(plan, Z, Z * rr, ETA)
rp, col = plan.pattern()
out = {"rowptr": rp, "col": col}
done = 0
for nit in (20, 150):
    sol.iterate(nit - done, None, 1, None)
    done = nit
    torch.cuda.synchronize()
    sol.xavg_matrix(1.0 / nit, None)
    torch.cuda.synchronize()
    out["val%d" % nit] = sol.matrix_values()
os.makedirs("gpurun_out", exist_ok=True)
np.savez("gpurun_out/xavg_cfg4.npz", **out)
print("saved", {k: v.shape for k, v in out.items()})
