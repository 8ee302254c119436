"""The device-side plan builder (csrc/plan_device.cu) must produce the SAME plan as the host builder
(csrc/plan_host.cpp): every array of the plan image, byte for byte, and the same errors.  The host
builder is the one the not-gpu suite pins against the oracle's build_problem and the golden fixtures."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sig_sdp_mmw_b200 import _lib, mmw  # noqa: E402
from sig_sdp_mmw_b200.topology import sparse_env  # noqa: E402
from tests.test_gpu_parity import _synthetic_state  # noqa: E402

pytestmark = pytest.mark.gpu


def _require_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    _lib.load()


class _builder:
    def __init__(self, which):
        self.which = which

    def __enter__(self):
        self.old = os.environ.get("SIGSDP_PLAN_BUILDER")
        os.environ["SIGSDP_PLAN_BUILDER"] = self.which

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop("SIGSDP_PLAN_BUILDER", None)
        else:
            os.environ["SIGSDP_PLAN_BUILDER"] = self.old


def _both(state, order):
    with _builder("host"):
        ph = _lib.Plan(state, device=0, order=order)
    with _builder("device"):
        pd = _lib.Plan(state, device=0, order=order)
    return ph, pd


def _assert_same_plan(ph, pd):
    for k in ("n", "E_g", "E_a", "nnz", "nnzT", "order", "max_row"):
        assert getattr(ph, k) == getattr(pd, k), k
    ih, idv = ph.image(), pd.image()
    assert ih.shape == idv.shape
    assert np.array_equal(ih, idv)
    for a, b in zip(ph.edges(), pd.edges()):
        assert np.array_equal(a, b)
    for a, b in zip(ph.vectors(), pd.vectors()):
        assert np.array_equal(a, b)


def _with_explicit_zeros(state, seed):
    """Stored zeros in S and Q (the reference's eliminate_zeros / != 0 tests must see them as absent)."""
    S, Q, h = state
    rs = np.random.RandomState(seed)
    S = S.copy().tocsr()
    Q = Q.copy().tocsr()
    if S.nnz:
        z = rs.choice(S.nnz, size=max(1, S.nnz // 7), replace=False)
        S.data[z] = 0.0
    if Q.nnz:
        # keep Q symmetric: zero both (i, j) and (j, i)
        Qc = Q.tocoo()
        pick = rs.rand(Qc.nnz) < 0.1
        kill = set()
        for i, j in zip(Qc.row[pick], Qc.col[pick]):
            kill.add((min(i, j), max(i, j)))
        data = np.array([0.0 if (min(i, j), max(i, j)) in kill else v for i, j, v in zip(Qc.row, Qc.col, Qc.data)])
        Q = sp.csr_matrix((data, Q.indices, Q.indptr), shape=Q.shape)   # same structure, zeros stored
    return S, Q, h


@pytest.mark.parametrize("order", [0, 1])
@pytest.mark.parametrize("kw", [
    dict(n=2, seed=0, p_gain=1.0, groups=1),
    dict(n=40, seed=2, no_asso=True),
    dict(n=40, seed=3, no_gain=True, groups=5),
    dict(n=60, seed=4, p_gain=0.15, isolated=7),
    dict(n=300, seed=6, p_gain=0.05, groups=20),
])
def test_device_builder_equals_host_builder_small(kw, order):
    _require_gpu()
    state = _synthetic_state(**kw)
    _assert_same_plan(*_both(state, order))
    _assert_same_plan(*_both(_with_explicit_zeros(state, 11), order))


@pytest.mark.parametrize("cell_size,order", [(40, 1), (40, 0), (100, 1)])   # 4,800 and 30,000 stations
def test_device_builder_equals_host_builder_reference_graphs(cell_size, order):
    _require_gpu()
    state = sparse_env(cell_size=cell_size, sta_density_per_1m2=75e-4, seed=3).generate_S_Q_hmax()
    ph, pd = _both(state, order)
    _assert_same_plan(ph, pd)
    # and the solve on top of it is the same solve
    out = []
    for plan in (ph, pd):
        sol = _lib.Solver(plan, 6, 12, 0.05, _lib.F64, _lib.MODE_FUSED, -1)
        sol.iterate(5, None, 77, None)
        torch.cuda.synchronize()
        out.append(sol.X(True))
    for a, b in zip(*out):
        assert np.array_equal(a, b)


def test_device_builder_rejects_what_the_host_builder_rejects():
    _require_gpu()
    S, Q, h = _synthetic_state(n=40, seed=8, p_gain=0.5, groups=1)   # one clique: every gain pair is also an asso pair
    Q2 = Q.tolil()
    Q2[0, 1] = 0.0
    Q2 = Q2.tocsr()                                                   # asymmetric Q
    for which in ("host", "device"):
        with _builder(which):
            with pytest.raises(_lib.SigSdpError):
                _lib.Plan((S, Q2, h), device=0, order=1)
    # a pair that is both gain and asso cannot come out of the filter (asso pairs are dropped from T): both accept
    _assert_same_plan(*_both((S, Q, h), 1))


def test_default_builder_is_device_for_large_graphs_and_solver_matches():
    _require_gpu()
    state = sparse_env(cell_size=92, sta_density_per_1m2=75e-4, seed=5).generate_S_Q_hmax()   # 25,392 stations
    os.environ.pop("SIGSDP_PLAN_BUILDER", None)
    res = []
    for which in (None, "host"):
        if which:
            os.environ["SIGSDP_PLAN_BUILDER"] = which
        try:
            alg = mmw(nit=20, rank_radio=2, eta=0.05)
            np.random.seed(3)
            ok, X_half = alg.run_with_state(0, 4, state)
            assert ok
            res.append(X_half)
        finally:
            os.environ.pop("SIGSDP_PLAN_BUILDER", None)
    assert np.array_equal(res[0], res[1])
