"""GPU parity tests (-m gpu) for the kernels the benchmark actually runs: the two-chunk
staged Taylor / Gram kernels with the split-row slot table (`k_fused<double,16>` at D = 32,
`k_fused<float,16>` at D = 64, `<float,8>` at D = 32, `<double,32>` at D = 64), at fixture
sizes against the unmodified reference's outputs and at the BASELINE cfg2 / cfg3 / cfg4
sizes against the oracle on the same injected Omega (this also covers the wave-balanced
tile height, which only shrinks tiles for n > 18,944).  Plus: the device conflict counter
against the reference's own rounding.py, and the batch kernel against the oracle through
the exported Philox normals.

Tolerances as stated in tests/test_gpu_parity.py: fp64 1e-9 relative, fp32 sketch 2e-3."""
import numpy as np
import pytest

from oracle import mmw_oracle as orc
from tests.golden_util import load_case, load_r2_pins, omega_stream

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from sig_sdp_mmw_b200 import _lib, mmw  # noqa: E402
from sig_sdp_mmw_b200.batch import BatchSolver, instance_seed  # noqa: E402
from sig_sdp_mmw_b200.topology import sparse_env  # noqa: E402
from tests.test_gpu_parity import RTOL64, _assert_state_close, _require_gpu, _run_device, _run_oracle  # noqa: E402


def _assert_fp32_close(sol, st):
    Y, e_acc, _ = sol.dual()
    np.testing.assert_allclose(Y, st.Y, rtol=2e-3)
    Xd, Xg, Xa = sol.X(True)
    np.testing.assert_allclose(Xd, st.Xbar_d, rtol=2e-3)
    np.testing.assert_allclose(Xg, st.Xbar_g, rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(Xa, st.Xbar_a, rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(sol.sketch(), st.Yh, rtol=2e-3, atol=1e-4)


@pytest.mark.parametrize("name,lanes64,lanes32", [("n300_z16_d32", 16, 8), ("n500_z8_d64", 32, 16)])
def test_two_chunk_kernels_match_reference_fixture(name, lanes64, lanes32):
    """D = 32 and D = 64: every dtype takes the two-chunk staged kernels with a slot table."""
    g = load_case(name)
    nit = g["nit"]
    _, sol, om = _run_device(g, nit, order=1)
    assert sol.lanes == lanes64 and sol.tile_rows > 0 and sol.Dp == g["Z"] * g["rank_radio"]
    _, st = _run_oracle(g, nit, om)
    _assert_state_close(sol, st, RTOL64)
    np.testing.assert_allclose(sol.dual()[0], g["Y_last"], rtol=RTOL64)
    np.testing.assert_allclose(sol.sketch(), g["Yh_last"], rtol=RTOL64, atol=1e-13)
    h = sol.history(nit)
    np.testing.assert_array_equal(h["m_star"], g["m_star"])
    _, s32, _ = _run_device(g, nit, dtype=_lib.F32, order=1)
    assert s32.lanes == lanes32 and s32.tile_rows > 0
    _assert_fp32_close(s32, st)
    np.testing.assert_allclose(s32.sketch(), g["Yh_last"], rtol=2e-3, atol=1e-4)


BENCH_CASES = {
    # BASELINE configs 2-4 as bench.py builds them: env kwargs, Z, rank_radio, dtype, iterations
    "cfg2_5k": (dict(cell_size=50, sta_density_per_1m2=5e-3), 8, 8, "float32", 6),
    "cfg3_20k": (dict(cell_size=63, sta_density_per_1m2=125e-4), 16, 2, "float64", 5),
    "cfg4_100k": (dict(cell_size=200, sta_density_per_1m2=6.25e-3), 16, 2, "float64", 4),
}


@pytest.mark.parametrize("name", sorted(BENCH_CASES))
def test_benchmarked_configuration_matches_oracle(name):
    """The exact solver configuration `bench.py --workload <name>` times (same topology, Z, D,
    dtype, locality order, automatic tiling), a few iterations on injected Omega vs the oracle."""
    _require_gpu()
    kw, Z, rr, dtype, nit = BENCH_CASES[name]
    state = sparse_env(seed=0, **kw).generate_S_Q_hmax()
    K, D = state[0].shape[0], Z * rr
    om = np.random.RandomState(1).randn(nit, K, D)
    plan = _lib.Plan(state, device=0, order=1)
    code = _lib.F64 if dtype == "float64" else _lib.F32
    sol = _lib.Solver(plan, Z, D, 0.04, code)
    assert sol.tile_rows > 0 and sol.lanes == 16          # staged two-chunk kernels, 16 lanes per row
    if name == "cfg4_100k":
        assert sol.tile_rows < 64                         # wave-balanced tile height
    om_d = torch.from_numpy(om).cuda()
    sol.iterate(nit, om_d.data_ptr(), 0, None)
    torch.cuda.synchronize()
    del om_d
    p = orc.build_problem(Z, state)
    st = orc.MMWState(p, 0.04)
    for i in range(nit):
        st.step(om[i])
    if dtype == "float64":
        _assert_state_close(sol, st, RTOL64)
        np.testing.assert_array_equal(sol.history(nit)["nterms"], st.nterms)
    else:
        _assert_fp32_close(sol, st)
        # and the same configuration with an fp64 sketch is exact
        s64 = _lib.Solver(plan, Z, D, 0.04, _lib.F64)
        om_d = torch.from_numpy(om).cuda()
        s64.iterate(nit, om_d.data_ptr(), 0, None)
        torch.cuda.synchronize()
        _assert_state_close(s64, st, RTOL64)


def test_device_conflict_counter_matches_reference_rounding_py():
    """R2 pinned: sigsdp_round_conflicts vs rand_rounding.get_interference /
    get_violation_pct of the unmodified reference (rounding.py:56-66, fixture r2_pins)."""
    _require_gpu()
    alg = mmw(nit=1, eta=0.04)
    for kw, Z, z, I_ref, pct in load_r2_pins():
        state = sparse_env(**kw).generate_S_Q_hmax()
        n_vio, n_asso, I = alg.conflict_counts(z, state, return_interference=True)
        np.testing.assert_allclose(I, I_ref, rtol=1e-12, atol=1e-12)
        assert n_vio == int(round(pct * z.size))
        assert n_asso == orc.conflict_counts(z, state)[2]


@pytest.mark.parametrize("blocks", ["1", "auto"])
@pytest.mark.parametrize("dtype,code,tol", [("float64", _lib.F64, 1e-9), ("float32", _lib.F32, 2e-3)])
def test_batch_kernel_matches_oracle_on_exported_normals(dtype, code, tol, blocks, monkeypatch):
    """The batch kernel (cfg5) against the ORACLE: the Philox normals each instance draws are
    exported through sigsdp_debug_normals and fed to the oracle.  blocks = "1": one thread block
    per instance; "auto": a batch this small gets several co-resident blocks per instance (a team
    with its own barrier), the way 128 instances per GPU run on an 8-GPU split."""
    _require_gpu()
    if blocks == "1":
        monkeypatch.setenv("SIGSDP_BATCH_BLOCKS", "1")
    else:
        monkeypatch.delenv("SIGSDP_BATCH_BLOCKS", raising=False)
    states = [sparse_env(cell_size=5 + (i % 3), sta_density_per_1m2=75e-4, seed=20 + i).generate_S_Q_hmax() for i in range(5)]
    states.append(sparse_env(cell_size=20, sta_density_per_1m2=6.25e-3, seed=3).generate_S_Q_hmax())   # the cfg5 instance size
    Z, rr, eta, nit, seed = 8, 2, 0.04, 12, 42
    D = Z * rr
    bsol = BatchSolver(states, Z, eta, rank_radio=rr, dtype=dtype)
    bsol.iterate(nit, seed=seed)
    torch.cuda.synchronize()
    bpi = [b.blocks_per_instance() for b in bsol.batches]
    # (fp32 at D = 16 has 128-row tiles: the 75-node instances are one tile, i.e. one block, either way)
    assert all(x == 1 for x in bpi) if blocks == "1" else (max(bpi) > 1 or dtype == "float32")
    for i, state in enumerate(states):
        K = state[0].shape[0]
        p = orc.build_problem(Z, state)
        st = orc.MMWState(p, eta)
        for it in range(nit):
            st.step(_lib.debug_normals(instance_seed(seed, i), it, K, D, code))
        sol = bsol.solvers[i]
        if dtype == "float64":
            _assert_state_close(sol, st, tol)
        else:
            _assert_fp32_close(sol, st)


@pytest.mark.parametrize("kw,Z", [(dict(cell_size=63, sta_density_per_1m2=125e-4, seed=0), 16),     # cfg3 size, dense-ish
                                  (dict(cell_size=30, sta_density_per_1m2=75e-4, seed=1), 6)])      # few slots: many users stay unassigned
def test_device_greedy_pass_reproduces_the_sequential_pass(kw, Z):
    """sigsdp_round_greedy_device against the sequential host pass (itself pinned to the reference's rounding on
    its own seeds, tests/test_host_logic.py): bit-identical colours and remainder for the same factor and directions."""
    _require_gpu()
    from sig_sdp_mmw_b200 import rand_sdp_solver
    state = sparse_env(**kw).generate_S_Q_hmax()
    K = state[0].shape[0]
    gX = np.random.RandomState(5).randn(K, 2 * Z)
    out = {}
    for mode in ("device", "host"):
        alg = rand_sdp_solver()
        alg.greedy = mode
        np.random.seed(11)
        out[mode] = alg.rounding_one_attempt(Z, gX, state)
        if mode == "device":
            rounds = alg.last_greedy_rounds
    np.testing.assert_array_equal(out["device"][0], out["host"][0])
    assert out["device"][2] == out["host"][2]
    assert 1 <= rounds <= K
