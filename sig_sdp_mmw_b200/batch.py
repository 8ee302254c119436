"""Monte-Carlo sweeps: many independent small instances solved together, one thread block per
instance (sigsdp_batch_*).  The reference runs these sequentially in `for seed in
range(REPEAT)` loops (e.g. sim_script/journal_version/sim_all_bler.py:30-40); across GPUs
the instances are simply split between the ranks (no communication)."""
import numpy as np

from . import _lib

_PHI = 0x9E3779B97F4A7C15
_MASK = (1 << 64) - 1


def instance_seed(seed, i):
    """Philox key of instance i inside a batch launch (matches k_batch)."""
    return (int(seed) + _PHI * (i + 1)) & _MASK


def shard(n_items, rank, world):
    """Contiguous, balanced split of a batch between ranks: items [lo, hi) of rank `rank`."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class BatchSolver:
    """`states[i]` solved for `Zs[i]` slots; all instances share nit / eta / rank_radio / dtype."""

    def __init__(self, states, Zs, eta, rank_radio=2, dtype="float64", device=0, order=1):
        if np.isscalar(Zs):
            Zs = [int(Zs)] * len(states)
        assert len(Zs) == len(states) and len(states) > 0
        code = _lib.F64 if dtype in ("float64", "f64") else _lib.F32
        self.plans = [_lib.Plan(st, device=device, order=order) for st in states]
        self.solvers = [_lib.Solver(p, Z, Z * rank_radio, eta, code) for p, Z in zip(self.plans, Zs)]
        # one launch per lane-width class (instances with different Z may need different kernels);
        # every instance keeps the Omega stream of its global index
        groups = {}
        for i, s in enumerate(self.solvers):
            groups.setdefault(s.lanes, []).append(i)
        self.batches = [_lib.Batch([self.solvers[i] for i in idx], ids=idx) for idx in groups.values()]
        self.Zs, self.rank_radio = list(Zs), rank_radio

    def iterate(self, n_iters, seed=0, stream=None):
        for b in self.batches:
            b.iterate(n_iters, seed, stream)

    def total_terms(self):
        return sum(s.total_terms() for s in self.solvers)
