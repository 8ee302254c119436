"""lanczos.chebyshev_filtered_lanczos on the CPU: the native step routine (sigsdp_solver_lanczos_steps with
sigsdp_solver_lanczos_filter) is emulated in torch, so the host logic -- probe, cut from the Lanczos spectral
density, true-residual acceptance, the guards that hand the problem back to the plain solver -- is covered
without a GPU.  The GPU suite runs the same comparisons through the library (test_gpu_parity.py)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from sig_sdp_mmw_b200.lanczos import chebyshev_filtered_lanczos, thick_restart_lanczos


class _EmulatedNative:
    """What the library does per step: w = p(M) q_j, CGS2 against rows 0..j, alpha, beta, q_{j+1}."""

    def __init__(self, M):
        self.M = M
        self.deg, self.c, self.e = 0, 0.0, 1.0
        self.spmv = 0

    def matmat(self, X):
        self.spmv += X.shape[0]
        return torch.from_numpy(np.ascontiguousarray((self.M @ X.numpy().T).T))

    def set_filter(self, degree, lo, cut):
        self.deg = degree if degree >= 2 else 0
        self.c, self.e = 0.5 * (lo + cut), (0.5 * (cut - lo) if degree >= 2 else 1.0)

    def apply(self, x):
        if self.deg < 2:
            return self.matmat(x[None])[0]
        t0, t1 = x, (self.matmat(x[None])[0] - self.c * x) / self.e
        for _ in range(self.deg - 1):
            t0, t1 = t1, 2.0 * (self.matmat(t1[None])[0] - self.c * t1) / self.e - t0
        return t1

    def steps(self, Q, m, j0, j1, al, be):
        for j in range(j0, j1):
            w = self.apply(Q[j])
            h = Q[:j + 1] @ w
            w = w - Q[:j + 1].T @ h
            h2 = Q[:j + 1] @ w
            w = w - Q[:j + 1].T @ h2
            al[j] = h[j] + h2[j]
            beta = torch.linalg.norm(w)
            ok = bool(beta > 1e-12 * max(abs(float(al[j])), float(be[j - 1]) if j > 0 else 0.0))
            be[j] = beta if ok else 0.0
            Q[j + 1] = w / beta if ok else torch.zeros_like(w)


def _clustered_top_matrix(n, seed, lo=0.5, hi=1.5, negative=None):
    """Sparse symmetric matrix with a known spectrum shaped like X_avgd / nit: a dense bulk and 40 eigenvalues on
    top of it only 1e-3 apart, mixed by a few sparse Givens sweeps so that it is not diagonal."""
    rs = np.random.RandomState(seed)
    lam = lo + (hi - 0.045 - lo) * rs.rand(n)
    lam[:40] = hi - 1e-3 * np.arange(40) - 2e-4 * rs.rand(40)
    if negative is not None:
        lam[0] = negative
    M = sp.diags(lam).tocsr()
    for sweep in range(4):
        perm = rs.permutation(n)
        th = rs.rand(n // 2) * 2 * np.pi
        i, j = perm[0:2 * (n // 2):2], perm[1:2 * (n // 2):2]
        G = sp.lil_matrix((n, n))
        G.setdiag(1.0)
        G = G.tocsr().tolil()
        cs, sn = np.cos(th), np.sin(th)
        G[i, i] = cs; G[j, j] = cs; G[i, j] = sn; G[j, i] = -sn
        G = G.tocsr()
        M = (G @ M @ G.T).tocsr()
    return ((M + M.T) * 0.5).tocsr(), np.sort(lam)[::-1]


def _subspace_gap(V, W):
    """|| (I - V V^T) W ||_2 for orthonormal columns."""
    return float(np.linalg.norm(W - V @ (V.T @ W), 2))


@pytest.mark.parametrize("seed,k", [(0, 12), (1, 20)])
def test_filtered_solver_returns_the_top_eigenpairs_with_far_fewer_lanczos_steps(seed, k):
    n = 4000
    M, lam_true = _clustered_top_matrix(n, seed)
    v0 = torch.from_numpy(np.random.RandomState(7).randn(n))
    nat = _EmulatedNative(M)
    out = chebyshev_filtered_lanczos(nat.matmat, n, k, v0, nat.steps, nat.set_filter, tol=1e-10, probe_steps=60, degree=8)
    assert out is not None
    lam, V, info = out
    assert nat.deg == 0                                              # the filter is switched off again
    lam, V = lam.numpy(), V.numpy()
    assert np.allclose(np.sort(lam)[::-1], lam_true[:k], rtol=0, atol=1e-9)
    assert np.all(np.diff(np.abs(lam)) >= 0)                         # ascending |lambda| like svds
    R = M @ V - V * lam
    assert np.linalg.norm(R, axis=0).max() <= 1e-10 * np.abs(lam).max() * 1.01
    assert np.allclose(V.T @ V, np.eye(k), atol=1e-9)
    assert info["filter"]["cut"] < lam_true[k - 1] and info["filter"]["lo"] <= lam_true[-1]
    # against the plain solver: same subspace, several times fewer (re-orthogonalised) Lanczos steps
    nat2 = _EmulatedNative(M)
    lam2, V2, info2 = thick_restart_lanczos(nat2.matmat, n, k, "LM", v0, tol=1e-10, native_steps=nat2.steps)
    assert np.allclose(lam2.numpy(), lam, atol=1e-9)
    assert _subspace_gap(V2.numpy(), V) < 1e-6
    assert info["lanczos_steps"] * 2 < info2["matvecs"]


def test_filter_declines_when_a_large_negative_eigenvalue_makes_lm_two_sided():
    n = 3000
    M, lam_true = _clustered_top_matrix(n, 3, negative=-2.5)          # |-2.5| is the largest magnitude
    v0 = torch.from_numpy(np.random.RandomState(1).randn(n))
    nat = _EmulatedNative(M)
    assert chebyshev_filtered_lanczos(nat.matmat, n, 10, v0, nat.steps, nat.set_filter, tol=1e-10) is None
    assert nat.deg == 0


def test_filter_declines_on_probe_breakdown_and_small_problems():
    n = 2000
    M = sp.identity(n, format="csr") * 0.7                           # every vector is an eigenvector: beta_0 = 0
    v0 = torch.from_numpy(np.random.RandomState(2).randn(n))
    nat = _EmulatedNative(M)
    assert chebyshev_filtered_lanczos(nat.matmat, n, 5, v0, nat.steps, nat.set_filter) is None
    M3, _ = _clustered_top_matrix(3000, 6)
    nat3 = _EmulatedNative(M3)
    v3 = torch.from_numpy(np.random.RandomState(2).randn(3000))
    assert chebyshev_filtered_lanczos(nat3.matmat, 3000, 40, v3, nat3.steps, nat3.set_filter, probe_steps=60) is None   # 2k > probe
    assert nat3.spmv == 0
    M2, _ = _clustered_top_matrix(150, 4)
    nat2 = _EmulatedNative(M2)
    v2 = torch.from_numpy(np.random.RandomState(2).randn(150))
    assert chebyshev_filtered_lanczos(nat2.matmat, 150, 5, v2, nat2.steps, nat2.set_filter) is None


def test_a_cut_above_the_wanted_eigenvalues_is_caught_not_returned():
    """count_target far below k puts the cut above lambda_k: pairs inside the damped band must never be returned."""
    n = 4000
    M, lam_true = _clustered_top_matrix(n, 5)
    v0 = torch.from_numpy(np.random.RandomState(3).randn(n))
    nat = _EmulatedNative(M)
    out = chebyshev_filtered_lanczos(nat.matmat, n, 20, v0, nat.steps, nat.set_filter, tol=1e-10, count_target=4, max_restarts=4)
    if out is not None:                                              # (only if the estimate happened to be low enough)
        assert np.allclose(np.sort(out[0].numpy())[::-1], lam_true[:20], atol=1e-9)
    assert nat.deg == 0
