"""Worker of the row-shard GPU tests (run in its own process: a barrier that times out traps
the CUDA context, which must not take the test session with it).

    python tests/rowshard_worker.py group <case> <nranks> <dtype> <nit> [tiling]   shards of one process (one GPU)
    torchrun ... tests/rowshard_worker.py dist <case> <dtype> <nit>                one process per GPU

<case>: a golden fixture name, or cfg3 / cfg2 (the bench topologies).  Compares the sharded
solver with the oracle on the same injected Omega and prints "ROWSHARD OK ..."."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import mmw_oracle as orc  # noqa: E402
from tests.golden_util import load_case  # noqa: E402


def make_case(name):
    from sig_sdp_mmw_b200.topology import sparse_env
    if name == "cfg3":
        return sparse_env(cell_size=63, sta_density_per_1m2=125e-4, seed=0).generate_S_Q_hmax(), 16, 2, 0.04
    if name == "cfg2":
        return sparse_env(cell_size=50, sta_density_per_1m2=5e-3, seed=0).generate_S_Q_hmax(), 8, 8, 0.04
    g = load_case(name)
    return g["state"], g["Z"], g["rank_radio"], g["eta"]


def check(state, Z, rr, eta, nit, om, dual, X, Xbar, L, Yh, dtype):
    p = orc.build_problem(Z, state)
    st = orc.MMWState(p, eta)
    for i in range(nit):
        st.step(om[i])
    if dtype == "f64":
        rt, ax = 1e-9, 1e-13
        np.testing.assert_allclose(dual[0], st.Y, rtol=rt, atol=1e-300)
        np.testing.assert_allclose(dual[1], st.e_acc, rtol=rt, atol=rt)
        np.testing.assert_allclose(dual[2], st.Ybar, rtol=rt, atol=1e-300)
        for a, b in zip(X, (st.Xd, st.Xg, st.Xa)):
            np.testing.assert_allclose(a, b, rtol=rt, atol=ax)
        for a, b in zip(Xbar, (st.Xbar_d, st.Xbar_g, st.Xbar_a)):
            np.testing.assert_allclose(a, b, rtol=rt, atol=ax * 100)
        for a, b in zip(L, (st.Ld, st.Lg, st.La)):
            np.testing.assert_allclose(a, b, rtol=rt, atol=1e-14)
        np.testing.assert_allclose(Yh, st.Yh, rtol=rt, atol=ax)
    else:
        np.testing.assert_allclose(dual[0], st.Y, rtol=2e-3)
        np.testing.assert_allclose(Xbar[0], st.Xbar_d, rtol=2e-3)
        np.testing.assert_allclose(Xbar[1], st.Xbar_g, rtol=2e-3, atol=2e-3)
        np.testing.assert_allclose(Yh, st.Yh, rtol=2e-3, atol=1e-4)
    return st


def main():
    import torch
    from sig_sdp_mmw_b200 import _lib
    from sig_sdp_mmw_b200.rowshard import RowShardGroup, RowShardRank
    mode, case = sys.argv[1], sys.argv[2]
    state, Z, rr, eta = make_case(case)
    K, D = state[0].shape[0], Z * rr
    if mode == "group":
        nranks, dtype, nit = int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
        tiling = int(sys.argv[6]) if len(sys.argv) > 6 else -1
        code = _lib.F64 if dtype == "f64" else _lib.F32
        om = np.random.RandomState(3).randn(nit, K, D)
        om_d = torch.from_numpy(om).cuda()
        plan = _lib.Plan(state, device=0, order=1)
        grp = RowShardGroup(plan, Z, D, eta, nranks, dtype=code, tiling=tiling)
        infos = [s.shard_info() for s in grp.shards]
        assert infos[0]["row_lo"] == 0 and infos[-1]["row_hi"] == K
        assert all(a["row_hi"] == b["row_lo"] for a, b in zip(infos, infos[1:]))
        half = nit // 2                       # two launches: the barrier epochs carry over
        grp.iterate(half, om_d.data_ptr(), 0)
        grp.synchronize()
        grp.iterate(nit - half, om_d[half:].data_ptr(), 0)
        grp.synchronize()
        st = check(state, Z, rr, eta, nit, om, grp.gather_dual(), grp.gather_X(False), grp.gather_X(True), grp.gather_L(),
                   grp.gather_sketch(), dtype)
        # every rank took the same decisions from the same bits
        h0 = grp.shards[0].history(nit)
        for s in grp.shards[1:]:
            h = s.history(nit)
            for k in h0:
                np.testing.assert_array_equal(h[k], h0[k])
        if dtype == "f64":
            np.testing.assert_array_equal(h0["nterms"], st.nterms)
        # reset and run again: identical bits (deterministic reductions, epochs restart)
        Y1 = grp.gather_dual()[0]
        grp.reset()
        grp.iterate(nit, om_d.data_ptr(), 0)
        grp.synchronize()
        np.testing.assert_array_equal(grp.gather_dual()[0], Y1)
        print("ROWSHARD OK group case=%s nranks=%d dtype=%s grid=%s rows=%s halo_send=%s" % (
            case, nranks, dtype, [s.grid for s in grp.shards], [(i["row_lo"], i["row_hi"]) for i in infos],
            [i["halo_send_rows"] for i in infos]))
    else:
        import torch.distributed as dist
        dtype, nit = sys.argv[3], int(sys.argv[4])
        code = _lib.F64 if dtype == "f64" else _lib.F32
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        rank, world = dist.get_rank(), dist.get_world_size()
        om = np.random.RandomState(3).randn(nit, K, D)
        om_d = torch.from_numpy(om).cuda()
        plan = _lib.Plan.collective(state, local, 1, min_world=2)   # built by rank 0, broadcast over NCCL, imported here
        sh = RowShardRank(plan, Z, D, eta, dtype=code)
        sh.iterate(nit, om_d.data_ptr(), 0, None)
        torch.cuda.synchronize()
        out = (sh.gather_dual(), sh.gather_X(False), sh.gather_X(True), sh.gather_L(), sh.gather_sketch())
        h = sh.solver.history(nit)
        hs = [None] * world
        dist.all_gather_object(hs, {k: v.tolist() for k, v in h.items()})
        assert all(x == hs[0] for x in hs), "ranks disagree on the Taylor schedule / norms"
        if rank == 0:
            check(state, Z, rr, eta, nit, om, *out, dtype)
            print("ROWSHARD OK dist case=%s world=%d dtype=%s info=%s" % (case, world, dtype, sh.solver.shard_info()))
        sh.barrier()
        del sh
        # the drop-in object end to end: every rank gets the factor a single GPU computes for the same device Omega
        from sig_sdp_mmw_b200 import mmw
        kw = dict(nit=nit, eta=eta, rank_radio=rr, dtype="float64" if dtype == "f64" else "float32", omega="device", device=local, seed=7)
        ok, Xh = mmw(row_shard=True, **kw).run_with_state(0, Z, state)
        assert ok and Xh.shape == (K, min(K - 1, (Z - 1) * rr))
        G = Xh @ Xh.T if K <= 2000 else Xh[:500] @ Xh[:500].T
        Gs = [None] * world
        dist.all_gather_object(Gs, G)
        assert all(np.allclose(g, Gs[0], atol=1e-9) for g in Gs), "ranks returned different factors"
        if rank == 0:
            _, Xh1 = mmw(**kw).run_with_state(0, Z, state)
            G1 = Xh1 @ Xh1.T if K <= 2000 else Xh1[:500] @ Xh1[:500].T
            np.testing.assert_allclose(G, G1, atol=1e-8 if dtype == "f64" else 5e-3)
            print("ROWSHARD OK mmw(row_shard=True) case=%s world=%d" % (case, world))
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
