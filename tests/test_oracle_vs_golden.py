"""Pins oracle/mmw_oracle.py (the CPU restatement) against outputs of the
unmodified reference stored in tests/golden/*.npz (see oracle/make_golden.py)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import mmw_oracle as orc
from tests.golden_util import CASES, load_case, load_r2_pins, omega_stream


@pytest.fixture(scope="module", params=CASES)
def solved(request):
    g = load_case(request.param)
    K = g["state"][0].shape[0]
    D = g["Z"] * g["rank_radio"]
    om = omega_stream(g["seed"], K, D, g["nit"])
    out = orc.run(g["Z"], g["state"], g["nit"], g["eta"], lambda i: om[i],
                  rank_radio=g["rank_radio"], log_gap=g["log_gap"])
    return g, out


def test_dual_weights_and_sketch(solved):
    g, out = solved
    st = out["state"]
    np.testing.assert_allclose(st.Y, g["Y_last"], rtol=1e-10, atol=1e-300)
    np.testing.assert_allclose(st.Yh, g["Yh_last"], rtol=1e-10, atol=1e-13)


def test_loss_matrix(solved):
    g, out = solved
    p, st = out["problem"], out["state"]
    A = orc.sym_matrix(p, st.Ld / 2.0, st.Lg / 2.0, st.La / 2.0)
    ref = sp.csr_matrix((g["L_last_data"], g["L_last_indices"], g["L_last_indptr"]), shape=A.shape)
    assert abs(A - ref).max() <= 1e-12 * max(1.0, abs(ref).max())


def test_taylor_schedule(solved):
    g, out = solved
    # same (m*, s) and the same number of executed terms as scipy on every iteration
    assert list(out["state"].nterms) == list(g["nterms"])


def test_gap_rows(solved):
    g, out = solved
    if not g["log_gap"]:
        pytest.skip("case logged no gap")
    np.testing.assert_allclose(out["gap"], g["gap"], rtol=1e-9, atol=1e-10)


def test_final_factor(solved):
    g, out = solved
    Xh, ref = out["X_half"], g["X_half"]
    assert Xh.shape == ref.shape
    np.testing.assert_allclose(Xh @ Xh.T, ref @ ref.T, rtol=0, atol=1e-10)


def test_trace_scalars(solved):
    g, out = solved
    if "Y_all" not in g:
        pytest.skip("no per-iteration trace")
    # re-run and compare every iteration's Y and Y_h
    K = g["state"][0].shape[0]
    D = g["Z"] * g["rank_radio"]
    om = omega_stream(g["seed"], K, D, g["nit"])
    p = orc.build_problem(g["Z"], g["state"])
    st = orc.MMWState(p, g["eta"])
    for i in range(g["nit"]):
        st.step(om[i])
        np.testing.assert_allclose(st.Y, g["Y_all"][i], rtol=1e-10)
        np.testing.assert_allclose(st.Yh, g["Yh_all"][i], rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("name", CASES)
def test_rounding(name):
    g = load_case(name)
    for seed, z_ref, rem_ref in zip(g["round_seeds"], g["round_z"], g["round_rem"]):
        rs = np.random.RandomState(int(seed))
        z, Z, rem = orc.rounding(g["Z"], g["X_half"], g["state"],
                                 lambda a, b: rs.randn(a, b), lambda Zz, n: rs.randint(Zz, size=n))
        assert rem == int(rem_ref)
        np.testing.assert_array_equal(z, z_ref)
    rs = np.random.RandomState(2000)
    z, Z, rem = orc.rounding_one_attempt(g["Z"], g["X_half"], g["state"], rs.randn(g["Z"], g["X_half"].shape[1]),
                                         lambda n: rs.randint(g["Z"], size=n))
    assert rem == int(g["round1_rem"])
    np.testing.assert_array_equal(z, g["round1_z"])


def test_conflict_counts_match_reference_rounding_py():
    """R2: oracle.conflict_counts against the unmodified reference's
    rand_rounding.get_interference / get_violation_pct (rounding.py:56-66)."""
    from sig_sdp_mmw_b200.topology import sparse_env
    for kw, Z, z, I_ref, pct in load_r2_pins():
        state = sparse_env(**kw).generate_S_Q_hmax()
        I, n_vio, n_asso = orc.conflict_counts(z, state)
        np.testing.assert_allclose(I, I_ref, rtol=1e-13, atol=1e-13)
        assert n_vio == int(round(pct * z.size))
        Q = sp.triu(sp.csr_matrix(state[1]), k=1).tocoo()
        assert n_asso == int(np.sum(z[Q.row] == z[Q.col]))
