"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the oracle
on the same seeded inputs and against the committed reference fixtures.

Stated tolerances:
  fp64 : relative 1e-9 on Y, X, L_accu, Y_h, gap rows, Gram of X_half after the full run
  fp32 sketch : relative 2e-3 on the same (the dual / loss state stays fp64)
"""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import mmw_oracle as orc
from tests.golden_util import CASES, load_case, omega_stream

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from sig_sdp_mmw_b200 import _lib, mmw  # noqa: E402
from sig_sdp_mmw_b200.topology import sparse_env  # noqa: E402

RTOL64 = 1e-9


def _require_gpu():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    assert _lib.load().sigsdp_device_count() >= 1


def _run_device(g, nit, dtype=_lib.F64, mode=_lib.MODE_FUSED, order=0, chunk=None, tiling=-1):
    _require_gpu()
    S = g["state"][0]
    K = S.shape[0]
    D = g["Z"] * g["rank_radio"]
    om = np.stack(omega_stream(g["seed"], K, D, nit))
    plan = _lib.Plan(g["state"], device=0, order=order)
    sol = _lib.Solver(plan, g["Z"], D, g["eta"], dtype, mode, tiling)
    om_d = torch.from_numpy(om).cuda()
    chunk = chunk or nit
    done = 0
    while done < nit:
        c = min(chunk, nit - done)
        sol.iterate(c, om_d[done:done + c].contiguous().data_ptr(), 0, None)
        torch.cuda.synchronize()
        done += c
    return plan, sol, om


def _run_oracle(g, nit, om):
    p = orc.build_problem(g["Z"], g["state"])
    st = orc.MMWState(p, g["eta"])
    for i in range(nit):
        st.step(om[i])
    return p, st


def _assert_state_close(sol, st, rtol, atol_x=1e-13):
    Y, e_acc, Ybar = sol.dual()
    np.testing.assert_allclose(Y, st.Y, rtol=rtol, atol=1e-300)
    np.testing.assert_allclose(e_acc, st.e_acc, rtol=rtol, atol=rtol)
    # the oracle has already added nothing for the last Y: both hold sum_{j<nit} Y_j
    np.testing.assert_allclose(Ybar, st.Ybar, rtol=rtol, atol=1e-300)
    for a, b in zip(sol.X(False), (st.Xd, st.Xg, st.Xa)):
        np.testing.assert_allclose(a, b, rtol=rtol, atol=atol_x)
    for a, b in zip(sol.X(True), (st.Xbar_d, st.Xbar_g, st.Xbar_a)):
        np.testing.assert_allclose(a, b, rtol=rtol, atol=atol_x * 100)
    for a, b in zip(sol.L(), (st.Ld, st.Lg, st.La)):
        np.testing.assert_allclose(a, b, rtol=rtol, atol=1e-14)
    np.testing.assert_allclose(sol.sketch(), st.Yh, rtol=rtol, atol=atol_x)


@pytest.mark.parametrize("name", CASES)
def test_iterations_match_oracle_and_reference_fp64(name):
    g = load_case(name)
    nit = g["nit"]
    plan, sol, om = _run_device(g, nit)
    p, st = _run_oracle(g, nit, om)
    _assert_state_close(sol, st, RTOL64)
    # against the unmodified reference's own outputs (fixture)
    Y, _, _ = sol.dual()
    np.testing.assert_allclose(Y, g["Y_last"], rtol=RTOL64)
    np.testing.assert_allclose(sol.sketch(), g["Yh_last"], rtol=RTOL64, atol=1e-13)
    Ld, Lg, La = sol.L()
    A = orc.sym_matrix(p, Ld / 2, Lg / 2, La / 2)
    ref = sp.csr_matrix((g["L_last_data"], g["L_last_indices"], g["L_last_indptr"]), shape=A.shape)
    assert abs(A - ref).max() <= 1e-11 * max(1.0, abs(ref).max())
    # Taylor controller: same (m*, s) as scipy; executed terms equal up to rounding at the
    # early-exit threshold on a few iterations
    h = sol.history(nit)
    np.testing.assert_array_equal(h["m_star"], g["m_star"])
    np.testing.assert_array_equal(h["s"], g["s_scale"])
    np.testing.assert_allclose(h["a1norm"], g["a1norm"], rtol=1e-10)
    diff = np.abs(h["nterms"] - g["nterms"])
    assert diff.max() <= 1 and (diff > 0).mean() <= 0.1
    assert sol.total_terms() == int(h["nterms"].sum())


def test_every_iteration_matches_reference_trace():
    g = load_case("n75_z8")
    _require_gpu()
    K = g["state"][0].shape[0]
    D = g["Z"] * g["rank_radio"]
    om = omega_stream(g["seed"], K, D, g["nit"])
    plan = _lib.Plan(g["state"], device=0)
    sol = _lib.Solver(plan, g["Z"], D, g["eta"])
    for i in range(g["nit"]):
        om_d = torch.from_numpy(om[i]).cuda()
        sol.iterate(1, om_d.data_ptr(), 0, None)
        torch.cuda.synchronize()
        Y, _, _ = sol.dual()
        np.testing.assert_allclose(Y, g["Y_all"][i], rtol=RTOL64)
        np.testing.assert_allclose(sol.sketch(), g["Yh_all"][i], rtol=RTOL64, atol=1e-13)


@pytest.mark.parametrize("name", ["n75_z6_rr3", "n500_z13"])
def test_stepwise_mode_equals_fused(name):
    g = load_case(name)
    nit = min(g["nit"], 20)
    _, a, _ = _run_device(g, nit, mode=_lib.MODE_FUSED)
    _, b, _ = _run_device(g, nit, mode=_lib.MODE_STEPWISE)
    # same device functions, but compiled into different kernels (different FMA contraction)
    for x, y in zip(a.dual() + a.X(True) + a.L(), b.dual() + b.X(True) + b.L()):
        np.testing.assert_allclose(x, y, rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(a.sketch(), b.sketch(), rtol=1e-12, atol=1e-15)
    np.testing.assert_array_equal(a.history(nit)["nterms"], b.history(nit)["nterms"])


@pytest.mark.parametrize("name,order,tiling,dtype", [
    ("n300_z10", 0, 0, _lib.F64),      # direct-gather kernels
    ("n300_z10", 1, 32, _lib.F64),     # staged, small tiles
    ("n500_z13", 1, 128, _lib.F64),    # staged, large tiles, D = 26 (ragged lanes)
    ("n500_z4_cfg1", 0, -1, _lib.F64), # automatic
    ("n75_z6_rr3", 1, 64, _lib.F64),
])
def test_tiling_variants_match_oracle(name, order, tiling, dtype):
    g = load_case(name)
    nit = min(g["nit"], 40)
    _, sol, om = _run_device(g, nit, dtype=dtype, order=order, tiling=tiling)
    assert sol.tile_rows == (tiling if tiling >= 0 else sol.tile_rows)
    _, st = _run_oracle(g, nit, om)
    _assert_state_close(sol, st, RTOL64)
    _, step, _ = _run_device(g, nit, dtype=dtype, order=order, tiling=tiling, mode=_lib.MODE_STEPWISE)
    np.testing.assert_allclose(step.sketch(), sol.sketch(), rtol=1e-12, atol=1e-15)


def test_chunked_calls_equal_one_call():
    g = load_case("n300_z10")
    _, a, _ = _run_device(g, 30)
    _, b, _ = _run_device(g, 30, chunk=7)
    for x, y in zip(a.dual() + a.X(True), b.dual() + b.X(True)):
        np.testing.assert_array_equal(x, y)


@pytest.mark.parametrize("order", [1, 16])
def test_locality_renumbering_is_transparent(order):
    g = load_case("n300_z10")
    nit = 40
    _, sol, om = _run_device(g, nit, order=order)
    _, st = _run_oracle(g, nit, om)
    _assert_state_close(sol, st, RTOL64)


@pytest.mark.parametrize("name", ["n75_z8", "n500_z13"])
def test_fp32_sketch_within_stated_tolerance(name):
    g = load_case(name)
    nit = g["nit"]
    _, sol, om = _run_device(g, nit, dtype=_lib.F32)
    _, st = _run_oracle(g, nit, om)
    Y, e_acc, _ = sol.dual()
    np.testing.assert_allclose(Y, st.Y, rtol=2e-3)
    Xd, Xg, Xa = sol.X(True)
    np.testing.assert_allclose(Xd, st.Xbar_d, rtol=2e-3)
    np.testing.assert_allclose(Xg, st.Xbar_g, rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(sol.sketch(), st.Yh, rtol=2e-3, atol=1e-4)


def test_wide_sketch_multi_chunk_rows():
    """D larger than one group pass (D = 2 * 70 = 140 columns, fp64)."""
    g = load_case("n75_z8")
    g = dict(g, Z=70, rank_radio=2)
    nit = 6
    _, sol, om = _run_device(g, nit)
    _, st = _run_oracle(g, nit, om)
    _assert_state_close(sol, st, RTOL64)


def test_device_normals_moments_and_determinism():
    _require_gpu()
    for dt in (_lib.F64, _lib.F32):
        a = _lib.debug_normals(7, 3, 4096, 32, dt)
        b = _lib.debug_normals(7, 3, 4096, 32, dt)
        c = _lib.debug_normals(8, 3, 4096, 32, dt)
        np.testing.assert_array_equal(a, b)
        assert not np.array_equal(a, c)
        assert abs(a.mean()) < 0.01 and abs(a.std() - 1) < 0.01
        assert abs((a ** 3).mean()) < 0.03 and abs((a ** 4).mean() - 3) < 0.1
        assert abs(np.corrcoef(a[:, 0], a[:, 1])[0, 1]) < 0.05


def test_device_omega_mode_converges_like_oracle():
    """Throughput mode (Philox Omega): different stream, same statistics: the running-mean
    constraint violation e_max follows the oracle's trajectory within sampling noise."""
    g = load_case("n500_z13")
    _require_gpu()
    K = g["state"][0].shape[0]
    D = g["Z"] * g["rank_radio"]
    plan = _lib.Plan(g["state"], device=0)
    sol = _lib.Solver(plan, g["Z"], D, g["eta"])
    sol.iterate(60, None, 123, None)
    e_dev = sol.gap_prepare()
    p = orc.build_problem(g["Z"], g["state"])
    e_orc = []
    for seed in range(3):          # e_max is a max over ~2k constraints: noisy across streams
        st = orc.MMWState(p, g["eta"])
        rs = np.random.RandomState(seed)
        for i in range(61):
            st.step(rs.randn(K, D))
        e_orc.append(np.max(orc.dual_errors(p, st.Xbar_d / 61, st.Xbar_g / 61, st.Xbar_a / 61)))
    assert 0.6 * min(e_orc) <= e_dev <= 1.6 * max(e_orc)
    Xd, _, _ = sol.X(False)
    np.testing.assert_allclose(Xd.mean(), 1.0, rtol=1e-12)      # trace normalisation
    Y, _, _ = sol.dual()
    np.testing.assert_allclose(Y.sum(), 1.0, rtol=1e-12)


# --------------------------------------------------------------------- class mmw
@pytest.mark.parametrize("name", CASES)
def test_solver_object_matches_reference_outputs(name):
    """run_with_state on a seeded numpy stream reproduces the reference's returned factor
    (Gram and singular values), gap log and, from that factor, its rounding."""
    g = load_case(name)
    _require_gpu()
    alg = mmw(nit=g["nit"], rank_radio=g["rank_radio"], eta=g["eta"], log_gap=g["log_gap"])
    np.random.seed(g["seed"])
    ok, X_half = alg.run_with_state(0, g["Z"], g["state"])
    assert ok is True
    ref = g["X_half"]
    assert X_half.shape == ref.shape and X_half.dtype == np.float64
    np.testing.assert_allclose(X_half @ X_half.T, ref @ ref.T, rtol=0, atol=1e-9)
    np.testing.assert_allclose(np.sort((X_half ** 2).sum(axis=0)), np.sort((ref ** 2).sum(axis=0)), rtol=1e-9)
    K = g["state"][0].shape[0]
    for key in ("mmw_all_it", "mmw_state_process", "mmw_dual", "mmw_loss", "mmw_expm", "mmw_per_it", "mmw_xavg"):
        rows = alg.LOGGED_NP_DATA[key]
        assert rows.shape[1] == 6 and (rows[:, 3] == g["Z"]).all() and (rows[:, 4] == K).all()
    assert alg.LOGGED_NP_DATA["mmw_dual"].shape[0] == g["nit"]
    assert (alg.LOGGED_NP_DATA["mmw_per_it"][:, 5] > 0).all()
    if g["log_gap"]:
        gap = alg.LOGGED_NP_DATA["gap"][:, 3:]
        np.testing.assert_allclose(gap, g["gap"], rtol=1e-8, atol=1e-9)
    # rounding from the reference's factor on the reference's seeds: the sequential host pass and the device
    # rounds (sigsdp_round_greedy_device) must both reproduce the reference's colouring exactly
    for mode in ("host", "device"):
        alg.greedy = mode
        for seed, z_ref, rem_ref in zip(g["round_seeds"], g["round_z"], g["round_rem"]):
            np.random.seed(int(seed))
            z, Z, rem = alg.rounding(g["Z"], g["X_half"], g["state"])
            assert Z == g["Z"] and rem == int(rem_ref)
            assert z.dtype == np.float64
            np.testing.assert_array_equal(z, z_ref)
    alg.greedy = "auto"
    np.random.seed(2000)
    z, _, rem = alg.rounding_one_attempt(g["Z"], g["X_half"], g["state"])
    assert rem == int(g["round1_rem"])
    np.testing.assert_array_equal(z, g["round1_z"])


def test_state_is_not_mutated():
    g = load_case("n75_z8")
    S, Q, h = g["state"]
    S0, Q0, h0 = S.copy(), Q.copy(), h.copy()
    alg = mmw(nit=5, eta=0.04)
    _, gX = alg.run_with_state(0, g["Z"], g["state"])
    alg.rounding(g["Z"], gX, g["state"])
    assert abs(S - S0).nnz == 0 and abs(Q - Q0).nnz == 0 and (h == h0).all()


def test_argmax_colours_and_conflict_counts():
    g = load_case("n300_z10")
    _require_gpu()
    alg = mmw(nit=3, eta=0.04)
    alg.rounding_one_attempt(g["Z"], g["X_half"], g["state"])       # builds the plan
    rs = np.random.RandomState(5)
    randv = rs.randn(g["Z"], g["X_half"].shape[1])
    z = alg.argmax_colours(g["Z"], g["X_half"], randv)
    z_ref = orc.argmax_colours(g["X_half"], randv)
    np.testing.assert_array_equal(z, z_ref)
    n_vio, n_asso, I = alg.conflict_counts(z, g["state"], return_interference=True)
    I_ref, v_ref, a_ref = orc.conflict_counts(z_ref, g["state"])
    assert (n_vio, n_asso) == (v_ref, a_ref)
    np.testing.assert_allclose(I, I_ref, rtol=1e-12, atol=1e-12)
    # a proper colouring from the greedy pass has no association conflict
    np.random.seed(1000)
    zz, _, rem = alg.rounding(g["Z"], g["X_half"], g["state"])
    if rem == 0:
        assert alg.conflict_counts(zz, g["state"])[1] == 0


def Gd_ref(V, lam):
    return (V * lam.abs()) @ V.T


def test_lanczos_path_equals_dense_path():
    """Final factor through thick-restart Lanczos (forced) vs the dense route."""
    from sig_sdp_mmw_b200.lanczos import eig_dense, thick_restart_lanczos
    g = load_case("n500_z13")
    _, sol, _ = _run_device(g, 20)
    sol.xavg_matrix(1.0 / 20)
    n = sol.plan.n

    def mm(X):
        X = X.contiguous()
        Y = torch.empty_like(X)
        sol.symv(X.data_ptr(), Y.data_ptr(), X.shape[0], None)
        return Y
    dev = torch.device("cuda", 0)
    k = 24
    lam_d, V_d = eig_dense(mm, n, k, "LM", dev)
    v0 = torch.randn(n, dtype=torch.float64, device=dev)
    lam_l, V_l, info = thick_restart_lanczos(mm, n, k, "LM", v0, ncv=70)
    assert not info["dense"]
    # the same recurrence with the steps run natively in the library
    lam_n, V_n, info_n = thick_restart_lanczos(mm, n, k, "LM", v0, ncv=70, native_steps=mmw._native_steps(sol, torch))
    assert not info_n["dense"]
    np.testing.assert_allclose(lam_n.cpu().numpy(), lam_d.cpu().numpy(), rtol=1e-10)
    Gn = (V_n * lam_n.abs()) @ V_n.T
    assert float((Gd_ref(V_d, lam_d) - Gn).abs().max()) < 1e-9
    np.testing.assert_allclose(lam_l.cpu().numpy(), lam_d.cpu().numpy(), rtol=1e-10)
    Gd = (V_d * lam_d.abs()) @ V_d.T
    Gl = (V_l * lam_l.abs()) @ V_l.T
    assert float((Gd - Gl).abs().max()) < 1e-9
    # the materialised matrix is X_avgd / nit on the pattern
    vals = sol.matrix_values()
    rp, col = sol.plan.pattern()
    M = sp.csr_matrix((vals, col, rp), shape=(n, n))
    Xd, Xg, Xa = sol.X(True)
    p = orc.build_problem(g["Z"], g["state"])
    ref = orc.sym_matrix(p, Xd / 20, Xg / 20, Xa / 20)
    assert abs(M - ref).max() < 1e-14


def _mm_of(sol):
    def mm(X):
        X = X.contiguous()
        Y = torch.empty_like(X)
        sol.symv(X.data_ptr(), Y.data_ptr(), X.shape[0], None)
        return Y
    return mm


def test_native_filtered_step_applies_the_chebyshev_polynomial():
    """sigsdp_solver_lanczos_filter: one native step on p(M) = T_d((M - c)/e) satisfies alpha_0 q_0 + beta_0 q_1 = p(M) q_0
    with p(M) q_0 formed here from the library's plain mat-vec by the three-term recurrence."""
    g = load_case("n500_z13")
    _, sol, _ = _run_device(g, 20)
    sol.xavg_matrix(1.0 / 20)
    n = sol.plan.n
    mm = _mm_of(sol)
    dev = torch.device("cuda", 0)
    steps = mmw._native_steps(sol, torch)
    q0 = torch.randn(n, dtype=torch.float64, device=dev)
    q0 /= torch.linalg.norm(q0)
    for deg, lo, cut in ((2, 0.1, 0.9), (5, 0.3, 1.1), (8, 0.2, 1.0)):
        c, e = 0.5 * (lo + cut), 0.5 * (cut - lo)
        t0, t1 = q0[None], (mm(q0[None]) - c * q0[None]) / e
        for _ in range(deg - 1):
            t0, t1 = t1, 2.0 * (mm(t1) - c * t1) / e - t0
        Q = torch.zeros((3, n), dtype=torch.float64, device=dev)
        Q[0] = q0
        al = torch.zeros(2, dtype=torch.float64, device=dev)
        be = torch.zeros(2, dtype=torch.float64, device=dev)
        sol.lanczos_filter(deg, lo, cut)
        steps(Q, 2, 0, 1, al, be)
        torch.cuda.synchronize()
        got = al[0] * Q[0] + be[0] * Q[1]
        scale = float(t1.abs().max())
        assert float((got - t1[0]).abs().max()) <= 1e-12 * scale
    sol.lanczos_filter(0)
    Q = torch.zeros((3, n), dtype=torch.float64, device=dev)
    Q[0] = q0
    steps(Q, 2, 0, 1, al, be)
    torch.cuda.synchronize()
    assert float((al[0] * Q[0] + be[0] * Q[1] - mm(q0[None])[0]).abs().max()) <= 1e-13
    with pytest.raises(_lib.SigSdpError):
        sol.lanczos_filter(4, 1.0, 1.0)


def test_filtered_final_factor_equals_plain_final_factor():
    """The Chebyshev-filtered eigen-solver (default from EIG_FILTER_MIN_NODES nodes) and the plain thick-restart
    Lanczos return the same factor: same singular values, same X_half X_half^T (applied to random vectors; the
    factor itself is unique only up to rotations inside eigenvalue clusters)."""
    _require_gpu()
    from sig_sdp_mmw_b200.topology import sparse_env as _env
    state = _env(cell_size=80, sta_density_per_1m2=75e-4, seed=1).generate_S_Q_hmax()    # 19,200 stations
    Z, rr, nit = 8, 2, 30
    res = {}
    for filt in (True, False):
        alg = mmw(nit=nit, rank_radio=rr, eta=0.04, omega="device", seed=5)
        alg.eig_filter = filt
        ok, X_half = alg.run_with_state(0, Z, state)
        assert ok
        res[filt] = (X_half, alg.last_singular_values.copy(), dict(alg.last_eig_info))
    (Xf, sf, inf_f), (Xp, sp_, inf_p) = res[True], res[False]
    assert "filter" in inf_f and "filter" not in inf_p and inf_f["converged"] and inf_p["converged"]
    assert inf_f["lanczos_steps"] * 2 < inf_p["matvecs"]
    np.testing.assert_allclose(sf, sp_, rtol=0, atol=1e-9)
    R = np.random.RandomState(0).randn(Xf.shape[0], 4)
    Gf, Gp = Xf @ (Xf.T @ R), Xp @ (Xp.T @ R)
    assert np.abs(Gf - Gp).max() <= 1e-6 * np.abs(Gp).max()
    # the default picks the filter at this size, and declines below the threshold
    alg = mmw(nit=nit, rank_radio=rr, eta=0.04, omega="device", seed=5)
    assert alg.eig_filter == "auto" and state[0].shape[0] >= alg.EIG_FILTER_MIN_NODES
    alg.run_with_state(0, Z, state)
    assert "filter" in alg.last_eig_info


def test_lanczos_breakdown_identity_matrix():
    """nit = 1: X_avgd = X_0 = I (mmw.py:67,77), every Krylov space is invariant after one step
    (beta = 0).  The solver must return finite orthonormal eigenvectors with eigenvalue 1 instead
    of NaN -- natively and through the torch steps."""
    from sig_sdp_mmw_b200.lanczos import thick_restart_lanczos
    g = load_case("n300_z10")
    _, sol, _ = _run_device(g, 1)
    sol.xavg_matrix(1.0)
    n, dev = sol.plan.n, torch.device("cuda", 0)
    v0 = torch.randn(n, dtype=torch.float64, device=dev)
    for native in (mmw._native_steps(sol, torch), None):
        lam, V, info = thick_restart_lanczos(_mm_of(sol), n, 18, "LM", v0, ncv=40, native_steps=native)
        assert not info["dense"] and info["converged"] and info["breakdowns"] > 0
        assert torch.isfinite(V).all() and torch.isfinite(lam).all()
        np.testing.assert_allclose(lam.cpu().numpy(), 1.0, rtol=1e-12)
        assert float((V.T @ V - torch.eye(18, dtype=torch.float64, device=dev)).abs().max()) < 1e-10
    # and the drop-in object end to end
    alg = mmw(nit=1, eta=g["eta"], rank_radio=g["rank_radio"])
    np.random.seed(0)
    ok, X_half = alg.run_with_state(0, g["Z"], g["state"])
    assert ok and np.isfinite(X_half).all()
    np.testing.assert_allclose(X_half.T @ X_half, np.eye(X_half.shape[1]), atol=1e-10)


def test_lanczos_on_disconnected_graph_matches_dense():
    """Two graphs side by side (block-diagonal state): a Krylov space started inside one component
    never leaves it; the factor must still equal the dense eigen-decomposition of X_avgd / nit."""
    from sig_sdp_mmw_b200.lanczos import eig_dense, thick_restart_lanczos
    _require_gpu()
    a, b = load_case("n75_z8"), load_case("n300_z10")
    S = sp.block_diag([a["state"][0], b["state"][0]], format="csr")
    Q = sp.block_diag([a["state"][1], b["state"][1]], format="csr")
    h = np.concatenate([a["state"][2], b["state"][2]])
    state = (S, Q, h)
    Z, rr, nit = 8, 2, 12
    n = S.shape[0]
    om = np.random.RandomState(5).randn(nit, n, Z * rr)
    plan = _lib.Plan(state, device=0, order=1)
    sol = _lib.Solver(plan, Z, Z * rr, 0.04)
    om_d = torch.from_numpy(om).cuda()
    sol.iterate(nit, om_d.data_ptr(), 0, None)
    torch.cuda.synchronize()
    st = orc.MMWState(orc.build_problem(Z, state), 0.04)
    for i in range(nit):
        st.step(om[i])
    np.testing.assert_allclose(sol.dual()[0], st.Y, rtol=RTOL64)
    sol.xavg_matrix(1.0 / nit)
    dev = torch.device("cuda", 0)
    k = 14
    lam_d, V_d = eig_dense(_mm_of(sol), n, k, "LM", dev)
    # a start vector supported on the small component only: its Krylov space breaks down after 75 steps
    v0 = torch.zeros(n, dtype=torch.float64, device=dev)
    iperm = np.argsort(plan.perm())
    v0[torch.from_numpy(iperm[:75]).to(dev)] = torch.randn(75, dtype=torch.float64, device=dev)
    lam_l, V_l, info = thick_restart_lanczos(_mm_of(sol), n, k, "LM", v0, ncv=90, tol=1e-11,
                                            native_steps=mmw._native_steps(sol, torch))
    assert info["converged"] and info["breakdowns"] >= 1 and torch.isfinite(V_l).all()
    np.testing.assert_allclose(lam_l.cpu().numpy(), lam_d.cpu().numpy(), rtol=1e-10)
    assert float((Gd_ref(V_d, lam_d) - Gd_ref(V_l, lam_l)).abs().max()) < 1e-8


def test_warm_start_across_probes():
    """sigsdp_solver_warm_start: the dual state of one probe seeds the next (same plan, another Z)."""
    g = load_case("n300_z10")
    K = g["state"][0].shape[0]
    plan = _lib.Plan(g["state"], device=0, order=1)
    om = np.random.RandomState(1).randn(10, K, 2 * 9)
    om_d = torch.from_numpy(om).cuda()
    # from a solver that has not iterated: identical to a cold start, bit for bit
    cold, fresh, warm0 = (_lib.Solver(plan, 9, 18, g["eta"]) for _ in range(3))
    warm0.warm_start(fresh)
    for s_ in (cold, warm0):
        s_.iterate(10, om_d.data_ptr(), 0, None)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(warm0.dual()[0], cold.dual()[0])
    # from the previous probe (Z = 10): the weights start where that probe ended
    prev = _lib.Solver(plan, 10, 20, g["eta"])
    prev.iterate(30, None, 3, None)
    torch.cuda.synchronize()
    Yp, ep, _ = prev.dual()
    warm = _lib.Solver(plan, 9, 18, g["eta"])
    warm.warm_start(prev)
    Y0, e0, _ = warm.dual()
    np.testing.assert_array_equal(e0, ep)
    np.testing.assert_array_equal(Y0, Yp)
    warm.iterate(10, om_d.data_ptr(), 0, None)
    torch.cuda.synchronize()
    Y, e_acc, Ybar = warm.dual()
    assert np.isfinite(Y).all() and abs(Y.sum() - 1.0) < 1e-12 and np.all(Y > 0)
    assert np.abs(e_acc - ep).max() > 0 and abs(Ybar.sum() - 10.0) < 1e-9        # Y_avgd = Y_prev(warm) + 9 more weights
    assert np.abs(Y - cold.dual()[0]).max() > 1e-6                               # and it is a different trajectory
    with pytest.raises(_lib.SigSdpError):
        warm.warm_start(prev)                                                    # only before the first iteration
    # the drop-in object: second call on the same state starts warm
    alg = mmw(nit=20, eta=g["eta"], omega="device", warm_start=True)
    _, X1 = alg.run_with_state(0, 10, g["state"])
    first = alg.last_solver
    _, X2 = alg.run_with_state(1, 9, g["state"])
    assert np.isfinite(X2).all() and alg.last_solver is not first
    assert np.abs(alg.last_solver.dual()[1]).max() > np.abs(first.dual()[1]).max() * 0.5


# --------------------------------------------------- size-independent properties
def test_properties_at_5k_nodes():
    """cfg2-sized graph (5,000 nodes): invariants that need no oracle run."""
    _require_gpu()
    state = sparse_env(cell_size=50, sta_density_per_1m2=5e-3, seed=0).generate_S_Q_hmax()
    K = state[0].shape[0]
    Z, rr = 8, 2
    plan = _lib.Plan(state, device=0)
    for dt, tol in ((_lib.F64, 1e-11), (_lib.F32, 1e-5)):
        sol = _lib.Solver(plan, Z, Z * rr, 0.04, dt)
        sol.iterate(25, None, 1, None)
        Y, e_acc, Ybar = sol.dual()
        np.testing.assert_allclose(Y.sum(), 1.0, rtol=1e-12)
        np.testing.assert_allclose(Ybar.sum(), 25.0, rtol=1e-12)
        assert (Y > 0).all()
        Xd, Xg, Xa = sol.X(False)
        np.testing.assert_allclose(Xd.mean(), 1.0, rtol=tol)
        gi, gj, _, _, ai, aj = plan.edges()
        # Gram entries obey Cauchy-Schwarz against the diagonal
        assert (np.abs(Xg) <= np.sqrt(Xd[gi] * Xd[gj]) * (1 + 1e-6)).all()
        assert (np.abs(Xa) <= np.sqrt(Xd[ai] * Xd[aj]) * (1 + 1e-6)).all()
        Yh = sol.sketch()
        d = (Yh ** 2).sum(axis=1)
        np.testing.assert_allclose(d / d.mean(), Xd, rtol=max(tol, 1e-10) * 100)
        np.testing.assert_allclose((Yh[gi] * Yh[gj]).sum(axis=1) / d.mean(), Xg, rtol=1e-4, atol=max(tol, 1e-10) * 100)
        # determinism: a second solver with the same seed gives the same bits
        sol2 = _lib.Solver(plan, Z, Z * rr, 0.04, dt)
        sol2.iterate(25, None, 1, None)
        np.testing.assert_array_equal(sol2.dual()[0], Y)
    # oracle on the same injected Omega for a few iterations at this size
    rs = np.random.RandomState(3)
    om = rs.randn(4, K, Z * rr)
    sol = _lib.Solver(plan, Z, Z * rr, 0.04)
    om_d = torch.from_numpy(om).cuda()
    sol.iterate(4, om_d.data_ptr(), 0, None)
    torch.cuda.synchronize()
    p = orc.build_problem(Z, state)
    st = orc.MMWState(p, 0.04)
    for i in range(4):
        st.step(om[i])
    _assert_state_close(sol, st, RTOL64)


# ------------------------------------------------------------- ragged / degenerate inputs
def _synthetic_state(n, seed, p_gain=0.3, groups=3, isolated=0, no_gain=False, no_asso=False):
    """A state tuple with the reference's structure (positive gains, S diagonal = own link,
    Q = cliques of stations sharing an AP) but arbitrary sparsity, to exercise ragged cases."""
    rs = np.random.RandomState(seed)
    S = sp.random(n, n, density=0.0 if no_gain else p_gain, random_state=rs, format="lil")
    S = (S * 2.0).tolil()
    for i in range(n):
        S[i, i] = 3.0 + rs.rand()
    grp = rs.randint(groups, size=n)
    if isolated:
        S = S.tolil()
        for i in range(isolated):            # stations with no link to anybody
            S[i, :] = 0
            S[:, i] = 0
            S[i, i] = 3.5
            grp[i] = groups + i
    Q = sp.lil_matrix((n, n))
    if not no_asso:
        for a in range(groups):
            idx = np.nonzero(grp == a)[0]
            for i in idx:
                for j in idx:
                    if i != j:
                        Q[i, j] = 1.0
    S = S.tocsr()
    S.eliminate_zeros()
    S.sort_indices()
    Q = Q.tocsr()
    Q.sort_indices()
    h = S.diagonal() / 1.85 - 1.0
    return S, Q, h


@pytest.mark.parametrize("kw,Z,rr", [
    (dict(n=2, seed=0, p_gain=1.0, groups=1), 3, 2),            # the smallest graph the solver accepts
    (dict(n=9, seed=1, p_gain=0.4, groups=2), 5, 1),            # odd sketch width (D = 5, padded rows)
    (dict(n=40, seed=2, no_asso=True), 4, 2),                   # no association edges at all
    (dict(n=40, seed=3, no_gain=True, groups=5), 9, 2),         # no interference edges at all
    (dict(n=60, seed=4, p_gain=0.15, isolated=7), 6, 2),        # isolated stations (diagonal-only rows)
    (dict(n=33, seed=5, p_gain=0.9, groups=2), 17, 3),          # dense rows, D = 51 > one lane-group pass
])
def test_ragged_and_degenerate_graphs_match_oracle(kw, Z, rr):
    _require_gpu()
    state = _synthetic_state(**kw)
    K = state[0].shape[0]
    D = Z * rr
    nit = 8
    om = np.random.RandomState(9).randn(nit, K, D)
    om_d = torch.from_numpy(om).cuda()
    p = orc.build_problem(Z, state)
    st = orc.MMWState(p, 0.05)
    for i in range(nit):
        st.step(om[i])
    for order, tiling in ((0, 0), (1, -1)):
        plan = _lib.Plan(state, device=0, order=order)
        assert (plan.E_g, plan.E_a) == (p.E_g, p.E_a)
        sol = _lib.Solver(plan, Z, D, 0.05, _lib.F64, _lib.MODE_FUSED, tiling)
        sol.iterate(nit, om_d.data_ptr(), 0, None)
        torch.cuda.synchronize()
        _assert_state_close(sol, st, RTOL64)
    # and the drop-in object end to end (dense eigen path at these sizes)
    alg = mmw(nit=nit, rank_radio=rr, eta=0.05)
    np.random.seed(4)
    ok, X_half = alg.run_with_state(0, Z, state)
    assert ok and X_half.shape == (K, min(K - 1, (Z - 1) * rr)) and np.isfinite(X_half).all()
    z, _, rem = alg.rounding(Z, X_half, state)
    assert z.shape == (K,) and 0 <= rem <= K


def test_bad_arguments_raise():
    _require_gpu()
    g = load_case("n75_z8")
    plan = _lib.Plan(g["state"], device=0)
    with pytest.raises(_lib.SigSdpError):
        _lib.Solver(plan, 1, 2, 0.1)             # Z < 2
    with pytest.raises(_lib.SigSdpError):
        _lib.Solver(plan, 4, 8, -1.0)            # eta <= 0
    with pytest.raises(_lib.SigSdpError):
        _lib.Solver(plan, 4, 8, 0.1, 7)          # unknown dtype
    sol = _lib.Solver(plan, 4, 8, 0.1)
    with pytest.raises(_lib.SigSdpError):
        sol.sketch()                             # nothing computed yet
    with pytest.raises(_lib.SigSdpError):
        sol.history(5)
    with pytest.raises(_lib.SigSdpError):
        sol.split_step(True)                     # not a column shard
    with pytest.raises(_lib.SigSdpError):
        _lib.Solver(plan, 4, 4, 0.1, D_total=8, col0=1)   # misaligned shard
