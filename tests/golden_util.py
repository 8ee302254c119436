"""Helpers shared by the CPU and GPU parity tests: load the committed golden
fixtures (generated from the unmodified reference by oracle/make_golden.py)."""
import os

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
CASES = ["n75_z8", "n75_z6_rr3", "n300_z10", "n500_z4_cfg1", "n500_z13", "n1000_z8", "n300_z16_d32", "n500_z8_d64"]


def load_case(name):
    d = np.load(os.path.join(GOLD, name + ".npz"))
    g = {k: d[k] for k in d.files}
    S = sp.csr_matrix((g["S_data"], g["S_indices"], g["S_indptr"]), shape=tuple(g["S_shape"]))
    Q = sp.csr_matrix((g["Q_data"], g["Q_indices"], g["Q_indptr"]), shape=tuple(g["Q_shape"]))
    g["state"] = (S, Q, g["h_max"])
    for k in ("Z", "nit", "rank_radio", "seed"):
        g[k] = int(g[k])
    g["eta"] = float(g["eta"])
    g["log_gap"] = bool(int(g["log_gap"]))
    return g


def omega_stream(seed, K, D, nit):
    """The raw normal blocks the reference drew: np.random.seed(seed) then one
    randn(K, D) per iteration (mmw.py:226; no other draws on the deterministic
    expm branch)."""
    rs = np.random.RandomState(seed)
    return [rs.randn(K, D) for _ in range(nit)]


def load_r2_pins():
    """Colourings with the reference's own rand_rounding.get_interference / get_violation_pct
    outputs (rounding.py:56-66), see oracle/make_golden.py:run_r2_pins.  Yields
    (env kwargs, Z, z, I_ref, violation fraction)."""
    d = np.load(os.path.join(GOLD, "r2_pins.npz"))
    for tag in "abc":
        cs, dens, seed = d[tag + "_kw"]
        kw = dict(cell_size=int(cs), sta_density_per_1m2=float(dens), seed=int(seed))
        for j in range(3):
            yield kw, int(d[tag + "_Z"]), d["%s%d_z" % (tag, j)], d["%s%d_I" % (tag, j)], float(d["%s%d_pct" % (tag, j)])
