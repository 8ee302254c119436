"""Time of the device greedy pass alone (CUDA events around sigsdp_round_greedy_device) against the host pass."""
import ctypes as C, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from sig_sdp_mmw_b200 import _lib
from sig_sdp_mmw_b200.topology import sparse_env
lib = _lib.load()
for kw, Z in ((dict(cell_size=200, sta_density_per_1m2=6.25e-3, seed=0), 16), (dict(cell_size=63, sta_density_per_1m2=125e-4, seed=0), 16),
              (dict(cell_size=20, sta_density_per_1m2=6.25e-3, seed=0), 8)):
    state = sparse_env(**kw).generate_S_Q_hmax(); K = state[0].shape[0]
    plan = _lib.Plan(state, device=0, order=1)
    rs = np.random.RandomState(0)
    rank = torch.from_numpy(rs.permutation(K).astype(np.int32)).cuda()
    pref = torch.from_numpy(np.stack([rs.permutation(Z) for _ in range(K)]).astype(np.int32)).cuda()
    z = torch.empty(K, dtype=torch.int32, device="cuda")
    rem, rounds = C.c_int64(), C.c_int64()
    st = torch.cuda.current_stream().cuda_stream
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.sigsdp_round_greedy_device(plan.handle, Z, rank.data_ptr(), pref.data_ptr(), z.data_ptr(), C.byref(rem), C.byref(rounds), st))
        e1.record(); torch.cuda.synchronize()
    S, Q, h = plan._S, plan._Q, plan._h
    zi = np.empty(K, np.int32); r2 = C.c_int64(); rk = rank.cpu().numpy(); pf = np.ascontiguousarray(pref.cpu().numpy())
    t = time.perf_counter()
    _lib.check(lib.sigsdp_round_greedy(K, Z, _lib._p(S[0], C.c_int32), _lib._p(S[1], C.c_int32), _lib._p(S[2], C.c_double), _lib._p(Q[0], C.c_int32),
                                       _lib._p(Q[1], C.c_int32), _lib._p(Q[2], C.c_double), _lib._p(h, C.c_double), _lib._p(rk, C.c_int32), _lib._p(pf, C.c_int32),
                                       _lib._p(zi, C.c_int32), C.byref(r2)))
    th = time.perf_counter() - t
    print("n=%d Z=%d: device %.2f ms (%d rounds, %.1f us/round), host %.2f ms, identical=%s remainder=%d" % (
        K, Z, e0.elapsed_time(e1), rounds.value, 1e3 * e0.elapsed_time(e1) / max(rounds.value, 1), th * 1e3, bool((z.cpu().numpy() == zi).all()), rem.value))
