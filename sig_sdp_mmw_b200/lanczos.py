"""Eigen-solvers over the library's symmetric mat-vec (sigsdp_solver_symv), replacing the
reference's ARPACK calls: eigsh(k=1, which='SA') for the gap (mmw.py:115) and svds(k=r)
for the final factor (mmw.py:215).  Thick-restart Lanczos with full (CGS2)
re-orthogonalisation; all vectors stay on the device (torch fp64 tensors, cuBLAS GEMV /
GEMM for the dense algebra), one host synchronisation per restart."""
import torch


def _dense_from_matvec(matmat, n, device):
    eye = torch.eye(n, dtype=torch.float64, device=device)
    M = matmat(eye)          # rows of the result are M e_i = columns of M (symmetric)
    return 0.5 * (M + M.T)


def eig_dense(matmat, n, k, which, device):
    """Small problems: materialise M through the mat-vec and call the dense solver."""
    M = _dense_from_matvec(matmat, n, device)
    lam, V = torch.linalg.eigh(M)
    if which == "LM":
        idx = torch.argsort(lam.abs(), descending=True)[:k]
        idx = torch.flip(idx, dims=[0])            # ascending |lambda| like svds
    elif which == "SA":
        idx = torch.arange(k, device=device)
    else:
        raise ValueError(which)
    return lam[idx], V[:, idx]


def thick_restart_lanczos(matmat, n, k, which, v0, ncv=None, tol=1e-13, max_restarts=500):
    """k extreme eigenpairs of a symmetric operator.

    matmat(X): X is (nvec, n) row-stacked vectors -> (nvec, n) of M x.
    which: 'LM' (largest magnitude, ascending |lambda| on return) or 'SA' (smallest algebraic).
    v0: start vector (n,), device fp64.
    Returns (lam (k,), V (n, k), info dict)."""
    dev = v0.device
    if ncv is None:
        ncv = max(2 * k + 20, 40)
    m = int(min(ncv, n))
    if m >= n or k >= m - 1:
        lam, V = eig_dense(matmat, n, k, which, dev)
        return lam, V, dict(restarts=0, matvecs=n, dense=True)
    Q = torch.zeros((m + 1, n), dtype=torch.float64, device=dev)   # basis, one vector per row
    Tm = torch.zeros((m, m), dtype=torch.float64, device=dev)
    Q[0] = v0 / torch.linalg.norm(v0)
    nkeep = 0
    matvecs = 0
    theta = S = None
    for restart in range(max_restarts):
        for j in range(nkeep, m):
            w = matmat(Q[j:j + 1])[0]
            matvecs += 1
            basis = Q[:j + 1]
            h = basis @ w
            w = w - basis.T @ h
            h2 = basis @ w                       # second Gram-Schmidt pass
            w = w - basis.T @ h2
            Tm[j, j] = h[j] + h2[j]
            beta = torch.linalg.norm(w)
            Q[j + 1] = w / beta
            if j + 1 < m:
                Tm[j, j + 1] = beta
                Tm[j + 1, j] = beta
        theta, S = torch.linalg.eigh(Tm)
        if which == "LM":
            order = torch.argsort(theta.abs(), descending=True)
        else:
            order = torch.arange(m, device=dev)
        want = order[:k]
        resid = (beta * S[m - 1, want]).abs()
        scale = theta.abs().max()
        if bool((resid <= tol * scale).all()):
            break
        # thick restart: keep the wanted Ritz vectors plus a buffer of the next best
        nk = int(min(k + max(8, (m - k) // 3), m - 2))
        keep = order[:nk]
        Q[:nk] = S[:, keep].T @ Q[:m]
        Q[nk] = Q[m]
        arrow = beta * S[m - 1, keep]
        Tm.zero_()
        Tm[torch.arange(nk), torch.arange(nk)] = theta[keep]
        Tm[nk, :nk] = arrow
        Tm[:nk, nk] = arrow
        nkeep = nk
    lam = theta[want]
    V = (S[:, want].T @ Q[:m]).T
    if which == "LM":
        o = torch.argsort(lam.abs())
        lam, V = lam[o], V[:, o]
    return lam, V, dict(restarts=restart, matvecs=matvecs, dense=False,
                        resid=float(resid.max()), scale=float(scale))
