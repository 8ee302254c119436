"""ONE graph row-sharded across the GPUs of a box (BASELINE configs[3]; include/sigsdp_mmw.h,
"row sharding").

Rank r owns a contiguous range of rows of the locality-ordered pattern.  Inside
`sigsdp_solver_iterate` the ranks' persistent kernels talk to each other directly: the SpMM
epilogue stores the boundary rows of the new Taylor term into the neighbours' sketch blocks
through peer-mapped memory, and the grid barrier between the phases spans the GPUs and carries
the packed scalars (max / sums / norms / trace).  The host only sets the exchange up:

  RowShardRank   one process per GPU (torch.distributed): the IPC handles of the exchange
                 arenas are all-gathered once, then every rank just calls iterate().
  RowShardGroup  all shards driven by one process -- shards on ONE GPU (co-resident kernels on
                 separate streams: how the single-GPU test box exercises the whole exchange
                 protocol) or on several P2P-capable GPUs.

State fetches of a shard return the entries it owns and zeros elsewhere; `gather_*` sum them
over the ranks."""
import numpy as np

from . import _lib


def _sum_tuples(parts):
    out = [np.array(x, copy=True) for x in parts[0]]
    for p in parts[1:]:
        for o, x in zip(out, p):
            o += x
    return tuple(out)


class RowShardGroup:
    """`nranks` row shards of one (plan, Z, D, eta, dtype) in this process.  `plans`: one plan
    (all shards on its device) or one plan per rank (same state, different devices)."""

    def __init__(self, plans, Z, D, eta, nranks, dtype=_lib.F64, tiling=-1, max_blocks=None):
        import torch
        if isinstance(plans, _lib.Plan):
            plans = [plans] * nranks
        assert len(plans) == nranks
        same_dev = len({p.device for p in plans}) == 1
        if max_blocks is None:
            if same_dev and nranks > 1:
                sms = torch.cuda.get_device_properties(plans[0].device).multi_processor_count
                max_blocks = (2 * sms) // nranks          # co-resident: the ranks' kernels wait for each other
            else:
                max_blocks = 0
        self.plans, self.nranks = plans, nranks
        self.shards = [_lib.Solver(plans[r], Z, D, eta, dtype, tiling=tiling, rows=(r, nranks), max_blocks=max_blocks)
                       for r in range(nranks)]
        if nranks > 1:
            _lib.Solver.attach_local(self.shards)
        self.streams = [torch.cuda.Stream(device=p.device) for p in plans]
        self.torch = torch

    def iterate(self, n_iters, omega_dev_ptrs=None, seed=0):
        """omega_dev_ptrs: None, one device pointer (shards on one GPU) or one per rank."""
        if omega_dev_ptrs is None or isinstance(omega_dev_ptrs, int):
            omega_dev_ptrs = [omega_dev_ptrs] * self.nranks
        for s, st, om in zip(self.shards, self.streams, omega_dev_ptrs):
            s.iterate(n_iters, om, seed, st.cuda_stream)

    def synchronize(self):
        for st in self.streams:
            st.synchronize()

    def reset(self):
        self.synchronize()
        for s, st in zip(self.shards, self.streams):
            s.reset(st.cuda_stream)
        self.synchronize()

    def gather_dual(self):
        return _sum_tuples([s.dual() for s in self.shards])

    def gather_X(self, averaged=False):
        return _sum_tuples([s.X(averaged) for s in self.shards])

    def gather_L(self):
        return _sum_tuples([s.L() for s in self.shards])

    def gather_sketch(self):
        return _sum_tuples([(s.sketch(),) for s in self.shards])[0]

    def total_terms(self):
        return self.shards[0].total_terms()


class RowShardRank:
    """This process's shard of a graph row-sharded over a torch.distributed group (one process
    per GPU, NCCL or gloo for the handle exchange and the final gathers)."""

    def __init__(self, plan, Z, D, eta, dtype=_lib.F64, tiling=-1, group=None):
        import torch
        import torch.distributed as dist
        self.dist, self.torch, self.group = dist, torch, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.plan = plan
        self.solver = _lib.Solver(plan, Z, D, eta, dtype, tiling=tiling, rows=(self.rank, self.world))
        handles = [None] * self.world
        dist.all_gather_object(handles, self.solver.ipc_handle(), group=group)
        self.solver.attach_ipc(handles)
        self.barrier()

    def barrier(self):
        self.torch.cuda.synchronize()
        self.dist.barrier(group=self.group)

    def iterate(self, n_iters, omega_dev_ptr=None, seed=0, stream=None):
        self.solver.iterate(n_iters, omega_dev_ptr, seed, stream)

    def reset(self, stream=None):
        """Collective: every rank resets, then all ranks meet before anyone iterates again (a
        running kernel must not see a peer's epochs from before the reset)."""
        self.barrier()
        self.solver.reset(stream)
        self._completed = False
        self.barrier()

    def complete_averages(self):
        """All-reduce (sum, in place on the device, NCCL) of the running sums X_avgd and Y_avgd: a
        shard only ever writes the entries it owns, so afterwards every rank holds the complete
        sums and can run the final factor (mmw.py:202-216) or the gap log on its own solver.
        Once per solve: the foreign entries are no longer zero afterwards (reset() clears them)."""
        from .sharded import _DevView
        if getattr(self, "_completed", False):
            raise _lib.SigSdpError("complete_averages was already called since the last reset")
        self._completed = True
        torch, dist = self.torch, self.dist
        dev = torch.device("cuda", self.plan.device)
        for which in (_lib.ARR_X_AVGD, _lib.ARR_Y_AVGD):
            ptr, cnt = self.solver.device_array(which)
            t = torch.as_tensor(_DevView(ptr, cnt), device=dev)
            if dist.get_backend(self.group) == "nccl":
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            else:
                h = t.cpu()
                dist.all_reduce(h, op=dist.ReduceOp.SUM, group=self.group)
                t.copy_(h)

    def _allsum(self, arrays):
        torch, dist = self.torch, self.dist
        dev = torch.device("cuda", self.plan.device) if dist.get_backend(self.group) == "nccl" else torch.device("cpu")
        flat = torch.from_numpy(np.concatenate([np.ravel(a) for a in arrays])).to(dev)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat = flat.cpu().numpy()
        out, o = [], 0
        for a in arrays:
            out.append(flat[o:o + a.size].reshape(a.shape))
            o += a.size
        return tuple(out)

    def gather_dual(self):
        return self._allsum(self.solver.dual())

    def gather_X(self, averaged=False):
        return self._allsum(self.solver.X(averaged))

    def gather_L(self):
        return self._allsum(self.solver.L())

    def gather_sketch(self):
        return self._allsum((self.solver.sketch(),))[0]
