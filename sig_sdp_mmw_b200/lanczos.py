"""Eigen-solvers over the library's symmetric mat-vec (sigsdp_solver_symv), replacing the
reference's ARPACK calls: eigsh(k=1, which='SA') for the gap (mmw.py:115) and svds(k=r)
for the final factor (mmw.py:215).  Thick-restart Lanczos with full (CGS2)
re-orthogonalisation.  All n-sized vectors stay on the device (torch fp64 tensors; cuBLAS
GEMV / GEMM for the dense algebra); the Lanczos basis has a fixed shape so every step is the
same short sequence of launches; the small projected eigenproblem (ncv x ncv) is solved on
the host once per restart, which is the only synchronisation."""
import numpy as np
import torch


_blas_ctl = None


def _small_eigh(Tm):
    """Projected (ncv x ncv) eigenproblem on the host, on ONE BLAS thread: at this size the
    threaded LAPACK is no faster, and its worker threads keep spinning after the call, which
    slows whatever the host does next on its cores (e.g. the next plan build, measured 2x)."""
    global _blas_ctl
    try:
        if _blas_ctl is None:
            from threadpoolctl import ThreadpoolController
            _blas_ctl = ThreadpoolController()
        with _blas_ctl.limit(limits=1, user_api="blas"):
            return np.linalg.eigh(Tm)
    except ImportError:
        return np.linalg.eigh(Tm)


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


def thick_restart_lanczos(matmat, n, k, which, v0, ncv=None, tol=1e-12, max_restarts=500, native_steps=None, keep_extra=None,
                          accept=None):
    """k extreme eigenpairs of a symmetric operator.

    matmat(X): X is (nvec, n) row-stacked vectors -> (nvec, n) of M x.
    which: 'LM' (largest magnitude, ascending |lambda| on return), 'LA' (largest algebraic, descending) or
        'SA' (smallest algebraic).
    accept(theta_want, resid, scale, ritz_rows): optional convergence test replacing `resid <= tol * scale`;
        ritz_rows() returns the (k, n) wanted Ritz vectors of the current basis.
    v0: start vector (n,), device fp64.
    native_steps(Q, m, j0, j1, al, be): optional; runs Lanczos steps j0..j1-1 in the library
        (sigsdp_solver_lanczos_steps) instead of the torch launches below.
    keep_extra: Ritz pairs kept at a restart beyond the k wanted ones (default max(8, (ncv - k) / 6)).
    Returns (lam (k,), V (n, k), info dict)."""
    dev = v0.device
    if ncv is None:
        # (cfg4, k = 30, measured on the B200: ncv 100 / keep k + 23 -> 60 ms, ncv 80 / keep k + 8 -> 52 ms: about the
        # same 710-750 mat-vecs, but every step re-orthogonalises against a shorter basis)
        ncv = max(2 * k + 20, 48)
    m = int(min(ncv, n))
    if m >= n or k >= m - 1:
        lam, V = eig_dense(matmat, n, k, which, dev)
        return lam, V, dict(restarts=0, matvecs=n, dense=True)
    # basis: one vector per row; rows beyond the current step are kept at zero so the
    # orthogonalisation can always use the whole array (fixed shapes, no slicing)
    Q = torch.zeros((m + 1, n), dtype=torch.float64, device=dev)
    Qt = Q.T
    al = torch.zeros(m, dtype=torch.float64, device=dev)
    be = torch.zeros(m, dtype=torch.float64, device=dev)
    Q[0] = v0 / torch.linalg.norm(v0)
    nkeep = 0
    matvecs = 0
    arrow = None
    theta_keep = None
    gen = torch.Generator(device="cpu").manual_seed(987654321)
    breakdowns = 0

    def steps(j0, j1):
        if native_steps is not None:
            native_steps(Q, m, j0, j1, al, be)
            return
        for j in range(j0, j1):
            w = matmat(Q[j:j + 1])[0]
            h = torch.mv(Q, w)
            w = torch.addmv(w, Qt, h, alpha=-1.0)
            h2 = torch.mv(Q, w)                       # second Gram-Schmidt pass
            w = torch.addmv(w, Qt, h2, alpha=-1.0)
            al[j] = h[j] + h2[j]
            beta = torch.linalg.norm(w)
            prev = be[j - 1] if j > 0 else torch.zeros((), dtype=torch.float64, device=dev)
            ok = beta > 1e-12 * torch.maximum(al[j].abs(), prev)
            be[j] = torch.where(ok, beta, torch.zeros_like(beta))
            Q[j + 1] = torch.where(ok, w / beta, torch.zeros_like(w))

    def cycle(j0):
        """Lanczos steps j0 .. m-1.  A zero beta_j marks a breakdown (q_0..q_j span an invariant
        subspace: X_avgd = I after one iteration, an exhausted component of a disconnected
        graph): the steps after it produced zeros, so q_{j+1} is replaced by a fresh random
        direction orthogonalised against the basis (beta_j stays 0: the projected matrix
        decouples there) and the cycle resumes from j + 1."""
        nonlocal breakdowns
        steps(j0, m)
        while True:
            be_h = be.cpu().numpy()
            brk = np.nonzero(be_h[j0:m] == 0.0)[0]
            if brk.size == 0:
                return
            j = j0 + int(brk[0])
            breakdowns += 1
            if j + 1 >= m:
                Q[m].zero_()
                return
            v = torch.randn(n, dtype=torch.float64, generator=gen).to(dev)
            for _ in range(2):
                v = torch.addmv(v, Qt[:, :j + 1], torch.mv(Q[:j + 1], v), alpha=-1.0)
            Q[j + 1] = v / torch.linalg.norm(v)
            Q[j + 2:].zero_()
            j0 = j + 1
            steps(j0, m)

    import time as _t
    t_cycle = t_host = 0.0
    t_all = _t.perf_counter()
    for restart in range(max_restarts):
        t0 = _t.perf_counter()
        cycle(nkeep)
        al_h, be_h = al.cpu().numpy(), be.cpu().numpy()   # synchronises: the cycle's kernels are done
        t_cycle += _t.perf_counter() - t0
        t0 = _t.perf_counter()
        matvecs += m - nkeep
        # projected matrix on the host: diag(theta_keep) with its arrow row, then the tridiagonal tail
        Tm = np.zeros((m, m))
        if nkeep:
            Tm[np.arange(nkeep), np.arange(nkeep)] = theta_keep
            Tm[nkeep, :nkeep] = arrow
            Tm[:nkeep, nkeep] = arrow
        for j in range(nkeep, m):
            Tm[j, j] = al_h[j]
            if j + 1 < m:
                Tm[j, j + 1] = Tm[j + 1, j] = be_h[j]
        theta, S = _small_eigh(Tm)
        order = np.argsort(-np.abs(theta)) if which == "LM" else np.argsort(-theta) if which == "LA" else np.arange(m)
        want = order[:k]
        beta_m = be_h[m - 1]
        resid = np.abs(beta_m * S[m - 1, want])
        scale = np.abs(theta).max()
        if accept is None:
            converged = bool(np.all(resid <= tol * scale))
        else:
            converged = bool(accept(theta[want], resid, scale,
                                    lambda: torch.from_numpy(np.ascontiguousarray(S[:, want].T)).to(dev) @ Q[:m]))
        if converged:
            t_host += _t.perf_counter() - t0
            break
        # thick restart: keep the wanted Ritz vectors plus a buffer of the next best
        nk = int(min(k + (max(8, (m - k) // 6) if keep_extra is None else keep_extra), m - 2))
        keep = order[:nk]
        Sk = torch.from_numpy(np.ascontiguousarray(S[:, keep].T)).to(dev)
        last = Q[m].clone()
        Qnew = Sk @ Q[:m]
        Q.zero_()
        Q[:nk] = Qnew
        Q[nk] = last
        theta_keep = theta[keep]
        arrow = beta_m * S[m - 1, keep]
        nkeep = nk
        t_host += _t.perf_counter() - t0
    Sw = torch.from_numpy(np.ascontiguousarray(S[:, want].T)).to(dev)
    lam = torch.from_numpy(np.ascontiguousarray(theta[want])).to(dev)
    V = (Sw @ Q[:m]).T
    if which == "LM":
        o = torch.argsort(lam.abs())
        lam, V = lam[o], V[:, o]
    if not converged:
        import warnings
        warnings.warn("thick-restart Lanczos stopped after %d restarts with residual %.3e > %.1e * %.3e: the returned "
                      "eigenpairs are not converged" % (max_restarts, float(resid.max()), tol, float(scale)), RuntimeWarning)
    return lam, V, dict(restarts=restart, matvecs=matvecs, dense=False, converged=converged, breakdowns=breakdowns,
                        resid=float(resid.max()), scale=float(scale), ncv=m, cycle_s=round(t_cycle, 4),
                        restart_s=round(t_host, 4), total_s=round(_t.perf_counter() - t_all, 4))


class _FilterUnusable(Exception):
    pass


def chebyshev_filtered_lanczos(matmat, n, k, v0, native_steps, set_filter, tol=1e-10, probe_steps=60, degree=8, ncv=None,
                               count_target=None, max_restarts=10):
    """The k largest-|lambda| eigenpairs of a symmetric M whose wanted eigenvalues sit at the TOP of a spectrum that
    is dense there (X_avgd / nit: 100k eigenvalues in [0.52, 1.55], lambda_30 - lambda_31 = 2e-4), by thick-restart
    Lanczos on p(M) = T_degree((M - c) / e) instead of M (polynomial filtering: Zhou & Saad's Chebyshev filter inside
    Lanczos).  On M itself the Krylov space needs ~710-750 steps, each re-orthogonalised against the basis (the
    expensive part); on p(M) ~120 steps of `degree` mat-vecs do, because the filter maps everything below `cut` into
    [-1, 1] and stretches the top.

      1. probe: `probe_steps` unfiltered Lanczos steps from v0.  Their Ritz values and quadrature weights
         n * S[0, i]^2 estimate how many eigenvalues lie above a level (Lanczos spectral density); `cut` is the level
         with ~3k above it (conservative: the run time hardly depends on it, a cut ABOVE lambda_k would lose wanted
         pairs), `lo` = theta_min - beta (a lower bound of the spectrum).
      2. thick-restart Lanczos on p(M), started from the sum of the probe's top-k Ritz vectors; a restart whose
         estimated p-residuals are small is checked against M ITSELF: Rayleigh-Ritz of M on the k Ritz vectors,
         ||M v - lambda v|| <= tol * max|lambda| (the plain solver's criterion, on true residuals), and every
         returned lambda must lie above `cut` (p is monotone only there: that makes the top k of p(M) the top k of M).

    Returns (lam, V, info) like thick_restart_lanczos(which='LM'), or None when the filter is not applicable (probe
    breakdown, 2k > probe_steps, spectrum not safely one-sided, checks failed, no progress after 4 restarts, no
    convergence within max_restarts): the caller then runs the unfiltered solver.  set_filter(degree, lo, cut) switches the operator behind native_steps (degree 0: off)."""
    import time as _t
    dev = v0.device
    t_all = _t.perf_counter()
    m1 = int(probe_steps)
    m = int(ncv if ncv is not None else max(2 * k + 20, 48))
    if native_steps is None or n < 4 * max(m, m1) or k >= m - 1 or 2 * k > m1:
        return None      # (2k > probe steps: too few Ritz values above the wanted ones to place the cut)
    target = float(count_target if count_target is not None else 3 * k)
    set_filter(0, 0.0, 1.0)
    Q1 = torch.zeros((m1 + 1, n), dtype=torch.float64, device=dev)
    al = torch.zeros(m1, dtype=torch.float64, device=dev)
    be = torch.zeros(m1, dtype=torch.float64, device=dev)
    Q1[0] = v0 / torch.linalg.norm(v0)
    native_steps(Q1, m1, 0, m1, al, be)
    al_h, be_h = al.cpu().numpy(), be.cpu().numpy()
    if not (np.all(np.isfinite(al_h)) and np.all(np.isfinite(be_h)) and np.all(be_h > 0.0)):
        return None
    Tm = np.diag(al_h) + np.diag(be_h[:-1], 1) + np.diag(be_h[:-1], -1)
    th, S = _small_eigh(Tm)
    lo = float(th[0] - be_h[-1])
    cum = np.cumsum((n * S[0] ** 2)[::-1])
    i = int(min(np.searchsorted(cum, target), m1 - k))       # (never below the probe's k-th Ritz value from the bottom)
    cut = float(th[::-1][i])
    # one-sided: everything below -cut would be a large-|lambda| pair the filter cannot see
    if not (cut > lo and lo > -cut and cut - lo > 1e-8 * max(abs(cut), abs(lo))):
        return None
    v1 = torch.from_numpy(np.ascontiguousarray(S[:, -k:].sum(axis=1))).to(dev) @ Q1[:m1]
    del Q1
    c, e = 0.5 * (lo + cut), 0.5 * (cut - lo)

    def pmat(X):   # the same operator for the torch path of the solver (breakdown resume is native-free)
        t0, t1 = X, (matmat(X) - c * X) / e
        for _ in range(degree - 1):
            t0, t1 = t1, 2.0 * (matmat(t1) - c * t1) / e - t0
        return t1

    found = {}
    checks = [0]
    calls = [0]

    def accept(theta_want, resid, scale, ritz_rows):
        calls[0] += 1
        if calls[0] >= 4 and not np.all(resid <= 1e-3 * scale):
            # a good cut converges within 2-3 restarts; this one is not getting anywhere (cut above the wanted
            # eigenvalues, or a spectrum the filter does not separate): stop paying for it
            raise _FilterUnusable("no progress after %d restarts" % calls[0])
        if not np.all(resid <= 10.0 * tol * scale):
            return False
        checks[0] += 1
        V = ritz_rows()                                   # (k, n)
        MV = matmat(V)
        H = V @ MV.T
        # (the k x k Rayleigh quotient is solved on the host: a device eigh of a 30 x 30 matrix is ~1.5 ms of launches)
        lam_h, W_h = _small_eigh((0.5 * (H + H.T)).cpu().numpy())
        lam = torch.from_numpy(lam_h).to(dev)
        W = torch.from_numpy(np.ascontiguousarray(W_h)).to(dev)
        U = W.T @ V
        R = W.T @ MV - lam[:, None] * U
        rn_h = torch.linalg.norm(R, dim=1).cpu().numpy()
        if lam_h.min() <= cut:
            raise _FilterUnusable("a returned eigenvalue is not above the cut")
        if np.all(rn_h <= tol * np.abs(lam_h).max()):
            found.update(lam=lam, U=U, resid=float(rn_h.max()), scale=float(np.abs(lam_h).max()))
            return True
        return False

    set_filter(degree, lo, cut)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            _, _, info = thick_restart_lanczos(pmat, n, k, "LA", v1, ncv=m, tol=tol, max_restarts=max_restarts,
                                               native_steps=native_steps, accept=accept)
    except _FilterUnusable:
        return None
    finally:
        set_filter(0, 0.0, 1.0)
    if not found or info.get("dense"):
        return None
    lam, V = found["lam"], found["U"].T
    o = torch.argsort(lam.abs())
    info = dict(info)
    info.update(filter=dict(degree=degree, lo=lo, cut=cut, probe_steps=m1, checks=checks[0]),
                matvecs=m1 + info["matvecs"] * degree + k * checks[0], lanczos_steps=m1 + info["matvecs"],
                resid=found["resid"], scale=found["scale"], total_s=round(_t.perf_counter() - t_all, 4))
    return lam[o], V[:, o], info
