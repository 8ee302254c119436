"""Outer search over the number of slots Z with the reference's surface
(sim_src/alg/binary_search_relaxation.py:8-71): `set_bounds`, `run`, `search`, the
`force_lower_bound` / `force_full_bound` switches and the `bs_*` log rows.  Pure control
flow around `feasibility_check_alg.run_with_state` / `.rounding`; shipped so the drivers'
whole call chain can be exercised where the reference tree is not present.  The solver's
graph plan is Z-independent and cached, so every probe after the first skips the set-up."""
import math

import numpy as np

from .stats import STATS_OBJECT


class binary_search_relaxation(STATS_OBJECT):
    def __init__(self):
        self.feasibility_check_alg = None
        self.force_lower_bound = False
        self.force_full_bound = False
        self.verbose = False

    def set_bounds(self, state):
        """Lower bound: largest association clique (max row length of Q_asso) + 1; upper bound:
        max degree of the symmetrised interference graph + 1 (:13-29)."""
        lb = int(np.max(np.diff(state[1].indptr))) + 1
        if self.force_lower_bound:
            return lb, lb
        if self.force_full_bound:
            return 1, state[0].shape[0]
        S = state[0] + state[0].transpose()
        S = S.tocsr()
        S.setdiag(0)          # stored zeros are kept: the reference counts them too (:22-25)
        ub = int(np.max(np.diff(S.indptr))) + 1
        return lb, ub

    def run(self, state):
        tic = self._get_tic()
        left, right = self.set_bounds(state)
        self._add_np_log("bs_set_bounds", 0, np.array([left, right, self._get_tim(tic)]))
        tic = self._get_tic()
        Z, z_vec, rem, it = self.search(left, right, state)
        self._add_np_log("bs_search", 0, np.array([left, right, Z, rem, it, self._get_tim(tic)]))
        return z_vec, Z, rem

    def search(self, left, right, state):
        it = 0
        alg = self.feasibility_check_alg
        while True:
            mid = math.floor(float(left + right) / 2.)
            tic = self._get_tic()
            _, gX = alg.run_with_state(it, mid, state)
            t_slv = self._get_tim(tic)
            tic = self._get_tic()
            z_vec, Z, rem = alg.rounding(mid, gX, state)
            t_rnd = self._get_tim(tic)
            self._add_np_log("bs_search_per_it", it, np.array([left, right, mid, Z, rem, t_slv, t_rnd]))
            it += 1
            done = False
            if left < right and rem > 0:
                left = mid + 1
            elif left + 1 < right and rem == 0:
                right = mid
            elif left + 1 == right and rem == 0:
                done = True
            elif left >= right and rem == 0:
                done = True
            elif left >= right and rem > 0:
                left += 1
                right += 1
            if self.verbose:
                self._printalltime(left, right, mid, Z, rem, "++++++++++++++++++++")
            if done:
                return Z, z_vec, rem, it
