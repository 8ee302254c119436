import time, torch, numpy as np, sys
sys.path.insert(0, '.')
from sig_sdp_mmw_b200 import _lib
from sig_sdp_mmw_b200.topology import sparse_env
state = sparse_env(cell_size=200, sta_density_per_1m2=6.25e-3, seed=0).generate_S_Q_hmax()
plan = _lib.Plan(state, device=0, order=1)
sol = _lib.Solver(plan, 16, 32, 0.04)
sol.iterate(5, None, 1, None); torch.cuda.synchronize()
sol.xavg_matrix(0.2)
n, m = plan.n, 100
dev = torch.device("cuda", 0)
Q = torch.randn(m + 1, n, dtype=torch.float64, device=dev); Qt = Q.T
w = torch.randn(n, dtype=torch.float64, device=dev)
y = torch.empty_like(w)
def timeit(name, f, reps=200):
    for _ in range(5): f()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); print("%-28s %8.1f us" % (name, (time.perf_counter() - t) / reps * 1e6))
st = torch.cuda.current_stream().cuda_stream
timeit("symv (1 vec)", lambda: sol.symv(w.data_ptr(), y.data_ptr(), 1, st))
timeit("torch.mv(Q, w)", lambda: torch.mv(Q, w))
timeit("torch.addmv(w, Qt, h)", lambda: torch.addmv(w, Qt, torch.ones(m + 1, dtype=torch.float64, device=dev)))
h = torch.mv(Q, w)
timeit("addmv only", lambda: torch.addmv(w, Qt, h, alpha=-1.0))
timeit("norm", lambda: torch.linalg.norm(w))
timeit("Q[j+1] = w / beta", lambda: Q.__setitem__(5, w / 2.0))
al = torch.zeros(m, dtype=torch.float64, device=dev)
timeit("al[j] = h[j] + h2[j]", lambda: al.__setitem__(3, h[3] + h[4]))
timeit("Q @ Q.T (gram 101x101)", lambda: Q @ Qt, 20)
