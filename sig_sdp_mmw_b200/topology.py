"""Sparse twin of the reference's topology generator (sim_src/env/env.py:12-196).

The reference materialises dense n_sta x n_ap matrices (env.py:137,145-153) and a dense
K x K association block (env.py:176-186), which stops it at a few 10^4 stations.  This
module produces the same `state = (S_gain, Q_asso, h_max)` from the same constructor
arguments and seed using a k-d tree over the access points, so only station/AP pairs
within radio reach are ever touched: O(n * reach) time and memory, 100k+ stations in
seconds.  For sizes the reference can build, the output matches it (same pattern, values
to rounding): see tests/test_host_logic.py (test_sparse_topology_reproduces_reference_state, test_sparse_evaluate_sinr_bler_matches_reference) against the committed reference fixtures.

Input generator only -- not on the solver's hot path."""
import math

import numpy as np
import scipy.sparse as sp
import scipy.stats
from scipy.spatial import cKDTree

NOISE_FLOOR_DBM = -94.0        # env.py:9,90-92


def _loss_db(fre_hz, dis):     # env.py:94-98 log-distance path loss, at least one metre
    L = 20.0 * math.log10(fre_hz / 1e6) + 16 - 28
    return L + 28 * np.log10(dis + 1)


def _polyanskiy(snr_dec, L, B, T):   # env.py:108-112 finite-blocklength error model
    nu = -L * math.log(2.0) + B * T * math.log(1 + snr_dec)
    do = math.sqrt(B * T * (1.0 - 1.0 / ((1.0 + snr_dec) ** 2)))
    return scipy.stats.norm.sf(nu / do)


def min_sinr_dec(packet_bit, bandwidth, slot_time, max_err, a=-5.0, b=30.0, tol=0.1):
    """env.py:114-134: bisection on the dB SINR whose block error equals max_err."""
    def err(x):
        return _polyanskiy(10.0 ** (x / 10.0), packet_bit, bandwidth, slot_time) / max_err - 1.0
    if err(a) * err(b) >= 0:
        raise ValueError("bisection bracket does not change sign")
    while (err(a) - err(b)) > tol:
        mid = (a + b) / 2
        if err(mid) == 0:
            return 10.0 ** (mid / 10.0)
        if err(a) * err(mid) < 0:
            b = mid
        else:
            a = mid
    return 10.0 ** (((a + b) / 2) / 10.0)


class sparse_env:
    """Same constructor as the reference's `env` (env.py:12)."""

    def __init__(self, cell_edge=20., cell_size=20, sta_density_per_1m2=5e-3, fre_Hz=4e9, txp_dbm_hi=5.,
                 txp_offset=2., min_s_n_ratio=0.1, packet_bit=800, bandwidth=5e6, slot_time=1.25e-4,
                 max_err=1e-5, seed=1):
        self.cell_edge, self.cell_size = cell_edge, cell_size
        self.grid_edge = cell_edge * cell_size
        self.n_ap = int(cell_size ** 2)
        self.n_sta = int(cell_size ** 2 * (sta_density_per_1m2 * cell_edge ** 2))
        self.fre_Hz, self.txp_offset, self.min_s_n_ratio = fre_Hz, txp_offset, min_s_n_ratio
        self.packet_bit, self.bandwidth, self.slot_time, self.max_err = packet_bit, bandwidth, slot_time, max_err
        off = cell_edge / 2.0
        x = np.linspace(0 + off, self.grid_edge - off, cell_size)
        xx, yy = np.meshgrid(x, x)
        self.ap_locs = np.array((xx.ravel(), yy.ravel())).T                      # env.py:50-54
        self.sta_locs = np.random.default_rng(seed).uniform(low=0., high=self.grid_edge,
                                                            size=(self.n_sta, 2))  # env.py:13,56-57
        self.min_sinr = min_sinr_dec(packet_bit, bandwidth, slot_time, max_err)

    def _pairs(self, threshold=True):
        """(sta, ap, received SNR) for every pair above min_s_n_ratio, and each station's AP."""
        n = self.n_sta
        tree = cKDTree(self.ap_locs)
        _, near = tree.query(self.sta_locs, k=1)
        d0 = np.sqrt(((self.sta_locs - self.ap_locs[near]) ** 2).sum(axis=1))
        smax = -_loss_db(self.fre_Hz, d0)
        # power control (env.py:136-142): own AP receives txp_offset * min_sinr
        txp = 10.0 * math.log10(self.min_sinr) - (smax - NOISE_FLOOR_DBM) + 10.0 * math.log10(self.txp_offset)
        # reach: rxpr >= min_s_n_ratio  <=>  loss <= txp - noise - 10 log10(min_s_n_ratio)
        L0 = 20.0 * math.log10(self.fre_Hz / 1e6) + 16 - 28
        budget = txp - NOISE_FLOOR_DBM - 10.0 * math.log10(self.min_s_n_ratio)
        reach = 10.0 ** ((budget - L0) / 28.0) - 1.0
        cand = tree.query_ball_point(self.sta_locs, reach * (1 + 1e-9) + 1e-9)
        cnt = np.fromiter((len(c) for c in cand), dtype=np.int64, count=n)
        ks = np.repeat(np.arange(n), cnt)
        aps = np.fromiter((a for c in cand for a in c), dtype=np.int64, count=int(cnt.sum()))
        diff = self.sta_locs[ks] - self.ap_locs[aps]
        dis = np.sqrt(diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1])
        rx_db = txp[ks] - _loss_db(self.fre_Hz, dis) - NOISE_FLOOR_DBM
        rx = 10 ** (rx_db / 10.)
        keep = rx >= self.min_s_n_ratio                                           # env.py:151
        ks, aps, rx = ks[keep], aps[keep], rx[keep]
        rxpr = sp.csr_matrix((rx, (ks, aps)), shape=(n, self.n_ap))
        # association = arg-max received power (env.py:177), first index on ties
        asso = np.asarray(rxpr.argmax(axis=1)).ravel()
        return rxpr, asso

    def generate_S_Q_hmax(self):
        """env.py:168-196 without dense intermediates."""
        rxpr, asso = self._pairs()
        n = self.n_sta
        member = sp.csr_matrix((np.ones(n), (asso, np.arange(n))), shape=(self.n_ap, n))   # AP -> its stations
        S_gain = (rxpr @ member).tocsr()          # S[k, j] = rxpr[k, asso_j]; one term per (k, j)
        S_gain.eliminate_zeros()
        S_gain.sort_indices()
        Q = (member.T @ member).tocsr()
        Q.setdiag(0.)
        Q.eliminate_zeros()
        Q.data[:] = 1.0
        Q.sort_indices()
        h_max = S_gain.diagonal() / self.min_sinr - 1.
        return S_gain, Q, h_max

    # ---- evaluating a colouring (env.py:198-232) ------------------------------------
    def _rx_power(self, stas, aps):
        """Un-thresholded received SNR of stations `stas` at APs `aps` (pairwise arrays),
        env.py:157-166 (_compute_state_real)."""
        tree_d0 = self._d0 if hasattr(self, "_d0") else None
        if tree_d0 is None:
            tree = cKDTree(self.ap_locs)
            _, near = tree.query(self.sta_locs, k=1)
            self._near = near
            d0 = np.sqrt(((self.sta_locs - self.ap_locs[near]) ** 2).sum(axis=1))
            smax = -_loss_db(self.fre_Hz, d0)
            self._txp = 10.0 * math.log10(self.min_sinr) - (smax - NOISE_FLOOR_DBM) + 10.0 * math.log10(self.txp_offset)
            self._d0 = d0
        diff = self.sta_locs[stas] - self.ap_locs[aps]
        dis = np.sqrt(diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1])
        return 10 ** ((self._txp[stas] - _loss_db(self.fre_Hz, dis) - NOISE_FLOOR_DBM) / 10.)

    def evaluate_sinr(self, z, Z, exact=None, floor_ratio=3e-2):
        """SINR of every station under the colouring z (env.py:198-224): signal = own link,
        interference = power received at the station's AP from every other station of the same
        slot, noise = 1; among the stations of one AP that share a slot only the strongest keeps
        its SINR, the others get 1e-3.
        exact=True sums the interference over ALL same-slot stations like the reference (dense,
        O(n^2 / Z)); exact=False only over stations whose power at that AP is above
        floor_ratio * min_s_n_ratio (k-d tree; what makes 100k stations tractable).  Default:
        exact for n <= 4000."""
        z = np.asarray(z).astype(np.int64)
        n = self.n_sta
        if exact is None:
            exact = n <= 4000
        self._rx_power(np.arange(1), np.arange(1))          # initialise power control / association
        asso = self._near
        own = self._rx_power(np.arange(n), asso)
        interf = np.zeros(n)
        if exact:
            for zz in range(Z):
                idx = np.nonzero(z == zz)[0]
                if idx.size == 0:
                    continue
                J, Kk = np.meshgrid(idx, idx, indexing="ij")       # power of j at k's AP
                M = self._rx_power(J.ravel(), asso[Kk.ravel()]).reshape(idx.size, idx.size)
                np.fill_diagonal(M, 0.0)
                interf[idx] = M.sum(axis=0)
        else:
            L0 = 20.0 * math.log10(self.fre_Hz / 1e6) + 16 - 28
            budget = self._txp - NOISE_FLOOR_DBM - 10.0 * math.log10(self.min_s_n_ratio * floor_ratio)
            reach = 10.0 ** ((budget - L0) / 28.0) - 1.0
            tree = cKDTree(self.ap_locs)
            cand = tree.query_ball_point(self.sta_locs, reach)
            cnt = np.fromiter((len(c) for c in cand), dtype=np.int64, count=n)
            js = np.repeat(np.arange(n), cnt)
            aps = np.fromiter((a for c in cand for a in c), dtype=np.int64, count=int(cnt.sum()))
            pw = self._rx_power(js, aps)
            # total power per (AP, slot), then subtract the station's own contribution
            tot = sp.csr_matrix((pw, (aps, z[js])), shape=(self.n_ap, Z))
            interf = np.asarray(tot[asso, z]).ravel() - own
            interf = np.maximum(interf, 0.0)
        sinr = own / (interf + 1.0)
        # one winner per (AP, slot)
        # one winner per (AP, slot): the arg-max, first index on exact ties (the reference's
        # masked argmax).  Stations of one AP have the same power-controlled signal up to
        # rounding, so their SINRs differ in the last bits only: the exact path reproduces the
        # reference's pick, the truncated path may pick another, physically equivalent, one.
        key = asso * Z + z
        order = np.lexsort((np.arange(n), -sinr, key))
        first = np.ones(n, dtype=bool)
        first[1:] = key[order][1:] != key[order][:-1]
        out = np.full(n, 1e-3)
        win = order[first]
        out[win] = sinr[win]
        return out

    def evaluate_bler(self, z, Z, **kw):
        """Finite-blocklength block error rate of every station (env.py:226-232)."""
        snr = self.evaluate_sinr(z, Z, **kw)
        B, T, L = self.bandwidth, self.slot_time, self.packet_bit
        nu = -L * math.log(2.) + B * T * np.log(1 + snr)
        do = np.sqrt(B * T * (1. - 1. / ((1. + snr) ** 2)))
        return scipy.stats.norm.sf(nu / do)
