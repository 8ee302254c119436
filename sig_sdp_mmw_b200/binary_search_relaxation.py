"""Outer search for the smallest number of slots Z that the solver + rounding can colour.

Same surface as the reference's driver (sim_src/alg/binary_search_relaxation.py:8-71:
`feasibility_check_alg`, `force_lower_bound`, `force_full_bound`, `set_bounds`, `run`,
`search`, log keys `bs_set_bounds`, `bs_search`, `bs_search_per_it`) so scripts written
against it work with this package where the reference tree is absent.  It is control flow
only; the window update is written as a small transition function and checked against the
reference's own probe log (tests/test_driver_chain.py).  The solver keeps its graph plan
across probes (it does not depend on Z), so only the first probe pays for the set-up."""
import numpy as np

from .stats import STATS_OBJECT


def next_window(left, right, mid, feasible):
    """One step of the search window [left, right] after probing `mid`.

    feasible (rounding left nobody unassigned): shrink from above while the window is wider
    than two values, otherwise the probe is the answer.  Infeasible: move the lower end past
    the probe, or -- once the window has collapsed -- slide it up by one slot.
    Returns (left, right, finished)."""
    if feasible:
        if right - left >= 2:
            return left, mid, False
        return left, right, True
    if left < right:
        return mid + 1, right, False
    return left + 1, right + 1, False


class binary_search_relaxation(STATS_OBJECT):
    def __init__(self):
        self.feasibility_check_alg = None
        self.force_lower_bound = False
        self.force_full_bound = False
        self.verbose = False

    def set_bounds(self, state):
        """(lower, upper) for Z.  Lower: the largest group of stations that share an access
        point needs one slot each, i.e. the longest row of Q_asso plus one.  Upper: one more than
        the largest degree of the symmetrised interference pattern (entries on the diagonal are
        stored and counted, as the reference does)."""
        S_gain, Q_asso = state[0], state[1]
        lower = int(np.diff(Q_asso.indptr).max()) + 1
        if self.force_lower_bound:
            return lower, lower
        if self.force_full_bound:
            return 1, S_gain.shape[0]
        sym = (S_gain + S_gain.T).tocsr()
        sym.setdiag(0)
        upper = int(np.diff(sym.indptr).max()) + 1
        return lower, upper

    def run(self, state):
        t = self._get_tic()
        lo, hi = self.set_bounds(state)
        self._add_np_log("bs_set_bounds", 0, np.array([lo, hi, self._get_tim(t)]))
        t = self._get_tic()
        Z, z_vec, rem, probes = self.search(lo, hi, state)
        self._add_np_log("bs_search", 0, np.array([lo, hi, Z, rem, probes, self._get_tim(t)]))
        return z_vec, Z, rem

    def search(self, left, right, state):
        solver = self.feasibility_check_alg
        probes = 0
        finished = False
        while not finished:
            mid = (left + right) // 2
            t = self._get_tic()
            _, gX = solver.run_with_state(probes, mid, state)
            us_solve = self._get_tim(t)
            t = self._get_tic()
            z_vec, Z, rem = solver.rounding(mid, gX, state)
            us_round = self._get_tim(t)
            self._add_np_log("bs_search_per_it", probes, np.array([left, right, mid, Z, rem, us_solve, us_round]))
            probes += 1
            left, right, finished = next_window(left, right, mid, rem == 0)
            if self.verbose:
                self._printalltime(left, right, mid, Z, rem, "+" * 20)
        return Z, z_vec, rem, probes
