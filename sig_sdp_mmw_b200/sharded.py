"""One graph across the GPUs of a box by sharding the SKETCH COLUMNS.

exp(L/2) Omega is computed column by column: the D columns are independent linear solves
with the same sparse matrix, so each rank runs the Taylor SpMMs on its own D/N columns with
no exchange at all.  What couples the columns is the Gram step, X_kc = <y_k, y_c> / tr:
every rank produces the partial dot products over its columns and the partial ||y_k||^2, and
ONE all-reduce (sum) per iteration -- nnzL + n doubles, 24 MB at 100k nodes -- completes
them.  The dual / loss state (e_accu, Y, L_accu) is replicated: every rank computes the same
bits from the same reduced buffer.

Compared with the row partition + all-gather of the sketch block on every Taylor term that
SURVEY section 8(e) starts from (n D w bytes, 4-5 times per iteration), this moves less data
and needs no locality-aware partition; its limit is Amdahl on the replicated dual / loss
phases.

The all-reduce is NCCL through torch.distributed; the library only exposes the buffer
(sigsdp_solver_exchange_buffer)."""
import numpy as np

from . import _lib


class _DevView:
    """CUDA-array-interface wrapper so torch can alias a library-owned device buffer."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3,
                                         "strides": None}


def column_shard(D, rank, world, vec):
    """Columns [col0, col0 + Dl) of rank `rank`: contiguous, multiples of `vec` columns."""
    units = D // vec
    if units * vec != D or units < world:
        raise ValueError("sketch width %d cannot be split over %d ranks in units of %d columns" % (D, world, vec))
    base, extra = divmod(units, world)
    lo = rank * base + min(rank, extra)
    return lo * vec, (base + (1 if rank < extra else 0)) * vec


class ShardedSolver:
    """Rank `rank` of `world` for one (plan, Z, D, eta, dtype).  `reduce_fn(tensor)` must sum
    the tensor over the ranks in place (torch.distributed.all_reduce by default)."""

    def __init__(self, plan, Z, D, eta, rank, world, dtype=_lib.F64, reduce_fn=None, group=None):
        import torch
        vec = 2 if dtype == _lib.F64 else 4
        self.col0, self.Dl = column_shard(D, rank, world, vec)
        self.solver = _lib.Solver(plan, Z, self.Dl, eta, dtype, D_total=D, col0=self.col0)
        ptr, cnt = self.solver.exchange_buffer()
        self.buffer = torch.as_tensor(_DevView(ptr, cnt), device=torch.device("cuda", plan.device))
        if reduce_fn is None:
            import torch.distributed as dist

            def reduce_fn(t):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        self.reduce_fn = reduce_fn
        self.rank, self.world = rank, world

    def iterate(self, n_iters, omega_dev=None, seed=0, stream=None):
        """omega_dev: optional torch tensor (n_iters, n, D_total) of raw normals, identical on
        every rank."""
        with self._on(stream) as st:
            for i in range(n_iters):
                ptr = omega_dev[i].data_ptr() if omega_dev is not None else None
                self.solver.split_step(True, ptr, seed, st)
                self.reduce_fn(self.buffer)

    def finish(self, stream=None):
        """Complete the last Gram; call before reading any state from self.solver."""
        with self._on(stream) as st:
            self.solver.split_step(False, None, 0, st)

    def _on(self, stream):
        """Kernel and collective must be ordered on ONE stream: the all-reduce runs on torch's
        current stream, so a caller-supplied raw stream is made current for the duration."""
        import contextlib
        import torch

        @contextlib.contextmanager
        def ctx():
            if stream is None or stream == torch.cuda.current_stream().cuda_stream:
                yield torch.cuda.current_stream().cuda_stream
            else:
                ext = torch.cuda.ExternalStream(stream)
                with torch.cuda.stream(ext):
                    yield stream
        return ctx()
