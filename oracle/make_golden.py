"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the UNMODIFIED
reference (run in the build container; /root/reference is not on the GPU box).

    python -m oracle.make_golden            # rewrites tests/golden/

What is recorded (all produced by the reference's own code, scipy 1.18.1 /
numpy 2.3.5): the env-generated state, the seed that drives np.random (so Omega
can be regenerated bit-exactly with np.random.seed(seed); np.random.randn), and
through *observation-only* wrappers around mmw.expm_half_randsk,
scipy.special.softmax and scipy's _fragment_3_1/_exact_inf_norm: the per-iteration
L_accu/2 values, sketch output Y_h, dual weights Y, Taylor (m*, s, executed
terms), plus the gap rows, the returned X_half and rounding results.
The wrappers call straight through; the reference's arithmetic is untouched.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_harness import load_reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

CASES = [
    # name, env kwargs, Z, nit, eta, rank_radio, log_gap, seed, trace (per-iter dumps)
    dict(name="n75_z8", env=dict(cell_size=5, sta_density_per_1m2=75e-4, seed=0),
         Z=8, nit=30, eta=0.04, rank_radio=2, log_gap=True, seed=11, trace=True),
    dict(name="n75_z6_rr3", env=dict(cell_size=5, sta_density_per_1m2=75e-4, seed=3),
         Z=6, nit=12, eta=0.1, rank_radio=3, log_gap=False, seed=5, trace=True),
    dict(name="n300_z10", env=dict(cell_size=10, sta_density_per_1m2=75e-4, seed=1),
         Z=10, nit=150, eta=0.04, rank_radio=2, log_gap=True, seed=7, trace=False),
    dict(name="n500_z4_cfg1", env=dict(cell_size=10, sta_density_per_1m2=125e-4, seed=0),
         Z=4, nit=150, eta=0.04, rank_radio=2, log_gap=False, seed=0, trace=False),
    dict(name="n500_z13", env=dict(cell_size=10, sta_density_per_1m2=125e-4, seed=0),
         Z=13, nit=60, eta=0.04, rank_radio=2, log_gap=False, seed=2, trace=False),
    # the instance size of BASELINE configs[4] (1,000-node Monte-Carlo instances)
    dict(name="n1000_z8", env=dict(cell_size=20, sta_density_per_1m2=6.25e-3, seed=4),
         Z=8, nit=40, eta=0.04, rank_radio=2, log_gap=False, seed=9, trace=False),
    # the sketch widths of the benchmarked kernels: D = 32 (BASELINE cfg3 / cfg4: Z = 16,
    # rank_radio = 2; the two-chunk fp64 kernels with 16 lanes per row) and D = 64 (cfg2: Z = 8,
    # rank_radio = 8; the fp32 two-chunk kernels, and the 32-lane fp64 ones)
    dict(name="n300_z16_d32", env=dict(cell_size=10, sta_density_per_1m2=75e-4, seed=2),
         Z=16, nit=60, eta=0.04, rank_radio=2, log_gap=False, seed=13, trace=False),
    dict(name="n500_z8_d64", env=dict(cell_size=10, sta_density_per_1m2=125e-4, seed=1),
         Z=8, nit=40, eta=0.04, rank_radio=8, log_gap=False, seed=17, trace=False),
]


def csr_pack(prefix, M, out):
    M = sp.csr_matrix(M)
    M.sort_indices()
    out[prefix + "_indptr"] = M.indptr.astype(np.int32)
    out[prefix + "_indices"] = M.indices.astype(np.int32)
    out[prefix + "_data"] = M.data.astype(np.float64)
    out[prefix + "_shape"] = np.array(M.shape, dtype=np.int64)


def run_case(ref, c):
    import scipy.special
    import scipy.sparse.linalg._expm_multiply as em
    mmw = ref.mmw
    e = ref.env(**c["env"])
    state = e.generate_S_Q_hmax()
    out = {}
    csr_pack("S", state[0], out)
    csr_pack("Q", state[1], out)
    out["h_max"] = np.asarray(state[2], dtype=np.float64)
    for k in ("Z", "nit", "eta", "rank_radio", "seed"):
        out[k] = np.array(c[k])
    out["log_gap"] = np.array(int(c["log_gap"]))

    rec = dict(L=[], Yh=[], Y=[], ms=[], ninf=[0], nterms=[], a1=[])
    orig_sk = mmw.__dict__["expm_half_randsk"].__func__
    orig_softmax = scipy.special.softmax
    orig_frag = em._fragment_3_1
    orig_inf = em._exact_inf_norm
    in_main = dict(flag=False)

    def sk(L, D):
        in_main["flag"] = True
        rec["ninf"][0] = 0
        n0 = len(rec["ms"])
        ret = orig_sk(L, D)
        in_main["flag"] = False
        if len(rec["ms"]) == n0:          # ||A||_1 == 0 path
            rec["ms"].append((0, 1))
            rec["a1"].append(0.0)
        m_star, s = rec["ms"][-1]
        rec["nterms"].append((rec["ninf"][0] - s) // 2)
        Lc = sp.csr_matrix(L)
        Lc.sort_indices()
        rec["L"].append((Lc.indptr.copy(), Lc.indices.copy(), Lc.data.copy()))
        rec["Yh"].append(ret.copy())
        return ret

    def softmax(x, *a, **k):
        y = orig_softmax(x, *a, **k)
        rec["Y"].append(np.array(y))
        return y

    def frag(norm_info, n0, tol, m_max=55, ell=2):
        m, s = orig_frag(norm_info, n0, tol, m_max=m_max, ell=ell)
        if in_main["flag"]:
            rec["ms"].append((int(m), int(s)))
            rec["a1"].append(float(norm_info.onenorm()))
        return m, s

    def infn(A):
        if in_main["flag"]:
            rec["ninf"][0] += 1
        return orig_inf(A)

    mmw.expm_half_randsk = staticmethod(sk)
    scipy.special.softmax = softmax
    em._fragment_3_1 = frag
    em._exact_inf_norm = infn
    try:
        alg = mmw(nit=c["nit"], rank_radio=c["rank_radio"], eta=c["eta"], log_gap=c["log_gap"])
        np.random.seed(c["seed"])
        ok, X_half = alg.run_with_state(0, c["Z"], state)
    finally:
        mmw.expm_half_randsk = staticmethod(orig_sk)
        scipy.special.softmax = orig_softmax
        em._fragment_3_1 = orig_frag
        em._exact_inf_norm = orig_inf

    out["X_half"] = X_half
    out["m_star"] = np.array([m for m, s in rec["ms"]], dtype=np.int32)
    out["s_scale"] = np.array([s for m, s in rec["ms"]], dtype=np.int32)
    out["nterms"] = np.array(rec["nterms"], dtype=np.int32)
    out["a1norm"] = np.array(rec["a1"])
    if c["log_gap"]:
        out["gap"] = alg.LOGGED_NP_DATA["gap"][:, 3:].copy()
    Yh = np.stack(rec["Yh"])
    Y = np.stack(rec["Y"])
    out["Y_last"] = Y[-1]
    out["Yh_last"] = Yh[-1]
    out["L_last_indptr"] = rec["L"][-1][0].astype(np.int32)
    out["L_last_indices"] = rec["L"][-1][1].astype(np.int32)
    out["L_last_data"] = rec["L"][-1][2]
    # per-iteration scalar traces (cheap, always stored)
    out["Y_max"] = Y.max(axis=1)
    out["Y_sq"] = (Y * Y).sum(axis=1)
    out["Yh_fro"] = np.sqrt((Yh * Yh).sum(axis=(1, 2)))
    out["L_trace"] = np.array([sp.csr_matrix((d, i, p), shape=state[0].shape).diagonal().sum()
                               for (p, i, d) in rec["L"]])
    if c["trace"]:
        out["Y_all"] = Y
        out["Yh_all"] = Yh
        out["L_all_data"] = np.stack([d for (_, _, d) in rec["L"]]) \
            if len({len(d) for (_, _, d) in rec["L"]}) == 1 else np.zeros(0)

    # rounding goldens on the reference's own factor (sdp_solver.py:18-107)
    zs, rems, seeds = [], [], []
    for rs in range(4):
        np.random.seed(1000 + rs)
        z_vec, Zr, rem = alg.rounding(c["Z"], X_half, state)
        zs.append(z_vec)
        rems.append(rem)
        seeds.append(1000 + rs)
    out["round_seeds"] = np.array(seeds)
    out["round_z"] = np.stack(zs)
    out["round_rem"] = np.array(rems)
    # one single attempt too (exact stream position known)
    np.random.seed(2000)
    z1, _, rem1 = alg.rounding_one_attempt(c["Z"], X_half, state)
    out["round1_z"] = z1
    out["round1_rem"] = np.array(rem1)
    return out


def run_binary_search_case(ref):
    """Whole driver chain (binary_search_relaxation.run -> mmw.run_with_state -> rounding) of
    the unmodified reference on the n=75 topology, np.random.seed(77): probe sequence
    [left, right, mid, Z, remainder], bounds, final colouring."""
    import contextlib
    import io
    e = ref.env(cell_size=5, sta_density_per_1m2=75e-4, seed=0)
    state = e.generate_S_Q_hmax()
    bs = ref.binary_search_relaxation()
    alg = ref.mmw(nit=40, eta=0.04)
    bs.feasibility_check_alg = alg
    np.random.seed(77)
    with contextlib.redirect_stdout(io.StringIO()):
        z_vec, Z, rem = bs.run(state)
    np.savez_compressed(os.path.join(GOLD, "bs_n75.npz"), z_vec=z_vec, Z=Z, rem=rem,
                        per_it=bs.LOGGED_NP_DATA["bs_search_per_it"][:, 3:8],
                        bounds=bs.LOGGED_NP_DATA["bs_set_bounds"][0, 3:5], seed=77, nit=40, eta=0.04)
    print("bs_n75: Z=%d rem=%d probes=%d" % (Z, rem, bs.LOGGED_NP_DATA["bs_search_per_it"].shape[0]))


def run_evaluate_case(ref):
    """env.evaluate_sinr / evaluate_bler (env.py:198-232) of the unmodified reference on two
    topologies for a fixed random colouring."""
    out = {}
    for name, kw in [("a", dict(cell_size=5, sta_density_per_1m2=75e-4, seed=0)),
                     ("b", dict(cell_size=10, sta_density_per_1m2=75e-4, seed=1))]:
        e = ref.env(**kw)
        Z = 9
        z = np.random.RandomState(3).randint(Z, size=e.n_sta).astype(float)
        out[name + "_z"], out[name + "_Z"] = z, Z
        out[name + "_sinr"], out[name + "_bler"] = e.evaluate_sinr(z, Z), e.evaluate_bler(z, Z)
        out[name + "_kw"] = np.array([kw["cell_size"], kw["sta_density_per_1m2"], kw["seed"]])
    np.savez_compressed(os.path.join(GOLD, "evaluate_n75_n300.npz"), **out)


def run_r2_pins(ref):
    """rounding.py:56-66 of the unmodified reference (rand_rounding.get_interference /
    get_violation_pct) on fixed colourings of three topologies: H = S_gain^T with an empty
    diagonal (interference received by user k from same-slot users), I_max = h_max.  Pins
    oracle.conflict_counts and the device counter (sigsdp_round_conflicts)."""
    rr = ref.rand_rounding
    out = {}
    for tag, kw, Z in [("a", dict(cell_size=5, sta_density_per_1m2=75e-4, seed=0), 6),
                       ("b", dict(cell_size=10, sta_density_per_1m2=75e-4, seed=1), 9),
                       ("c", dict(cell_size=10, sta_density_per_1m2=125e-4, seed=0), 13)]:
        e = ref.env(**kw)
        S, Q, h = e.generate_S_Q_hmax()
        H = np.asarray(sp.csr_matrix(S).transpose().todense(), dtype=np.float64)
        np.fill_diagonal(H, 0.)
        H = np.asarray(H)
        rs = np.random.RandomState(21)
        for j in range(3):
            z = rs.randint(Z, size=S.shape[0])
            I = np.asarray(rr.get_interference(H, z)).ravel()
            pct = rr.get_violation_pct(I, np.asarray(h, dtype=np.float64))
            out["%s%d_z" % (tag, j)] = z.astype(np.int64)
            out["%s%d_I" % (tag, j)] = I
            out["%s%d_pct" % (tag, j)] = np.array(pct)
        out[tag + "_kw"] = np.array([kw["cell_size"], kw["sta_density_per_1m2"], kw["seed"]])
        out[tag + "_Z"] = np.array(Z)
    np.savez_compressed(os.path.join(GOLD, "r2_pins.npz"), **out)
    print("r2_pins: %d colourings" % (len([k for k in out if k.endswith("_z")])))


def main():
    ref = load_reference()
    os.makedirs(GOLD, exist_ok=True)
    only = set(sys.argv[1:])          # optional: names of the cases to (re)generate
    if not only or "bs_n75" in only:
        run_binary_search_case(ref)
    if not only or "evaluate" in only:
        run_evaluate_case(ref)
    if not only or "r2_pins" in only:
        run_r2_pins(ref)
    for c in CASES:
        if only and c["name"] not in only:
            continue
        out = run_case(ref, c)
        path = os.path.join(GOLD, c["name"] + ".npz")
        np.savez_compressed(path, **out)
        print("%-16s K=%d nnzS=%d  nterms(avg)=%.1f  m*max=%d  rem=%s  -> %s (%.0f KB)" % (
            c["name"], out["h_max"].size, out["S_data"].size, out["nterms"].mean(),
            out["m_star"].max(), out["round_rem"].tolist(), os.path.relpath(path, ROOT),
            os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
