"""class mmw: the reference's solver object (sim_src/alg/mmw.py:12-229) with the work done
by hand-written sm_100a kernels behind the C ABI of include/sigsdp_mmw.h.

Same constructor, attributes, methods, return values and LOGGED_NP_DATA keys as the
reference, so `binary_search_relaxation.feasibility_check_alg = mmw(...)` and the
sim_script drivers run unchanged.  Extra keyword arguments select what the reference
has no notion of (all default to the reference's behaviour):

  dtype   : "float64" (default) or "float32" sketch block
  omega   : "numpy"  -- draw np.random.randn(K, D) per iteration on numpy's global stream
                       exactly where the reference does (mmw.py:226), so a seeded run
                       reproduces the reference's Omega and everything downstream;
            "device" -- counter-based Philox normals generated inside the kernel
                       (throughput mode; statistically equivalent, different stream)
  device  : CUDA device index
  order   : 1 (default) renumber nodes inside the kernels for locality (transparent: all
            inputs and outputs stay in the caller's numbering), 0 keep the caller's order
  warm_start : True -- successive run_with_state calls on the SAME state (the probes of the binary
            search, binary_search_relaxation.py:44-71) start from the previous call's dual state (e_accu, Y)
            instead of e_accu = 0, Y = 1/C; not something the reference does, off by default
  row_shard : True -- the process is one rank of a torch.distributed job (one process per
            GPU) and the ranks solve ONE graph together, each owning a strip of rows
            (sig_sdp_mmw_b200/rowshard.py, BASELINE configs[3]); every rank passes the same
            state and gets the same X_half back.  `shard_group`: the process group (default
            world).

There is no CPU fallback: a missing library or CUDA device raises."""
import math

import time

import numpy as np

from . import _lib
from .lanczos import chebyshev_filtered_lanczos, thick_restart_lanczos
from .sdp_solver import sdp_solver
from .stats import STATS_OBJECT

# host staging budget for the numpy Omega stream (bytes per kernel launch)
_OMEGA_CHUNK_BYTES = 96 << 20


class _Laps:
    """SIGSDP_PLAN_TIMING=1: stage times of the final factor on stderr (synchronises; diagnosis only)."""

    def __init__(self, torch):
        import os
        self.on = os.environ.get("SIGSDP_PLAN_TIMING") is not None
        self.torch = torch
        self.t = time.perf_counter() if self.on else 0.0

    def __call__(self, what):
        if not self.on:
            return
        import sys
        self.torch.cuda.synchronize()
        t = time.perf_counter()
        sys.stderr.write("[fact] %-24s %8.2f ms\n" % (what, 1e3 * (t - self.t)))
        self.t = t


class mmw(STATS_OBJECT, sdp_solver):
    EIG_FILTER_MIN_NODES = 10000   # see eig_filter below

    def __init__(self, nit=100, rank_radio=2, alpha=1., eta=0.1, log_gap=False,
                 dtype="float64", omega="numpy", device=0, order=1, seed=0, row_shard=False, shard_group=None,
                 warm_start=False):
        sdp_solver.__init__(self, nit=nit, rank_radio=rank_radio, alpha=alpha)
        self.eta = eta
        self.LOG_GAP = log_gap
        self.dtype = dtype
        self.omega = omega
        self.device = device
        self.plan_order = order
        self.seed = seed
        self.mode = _lib.MODE_FUSED
        self.eig_tol = 1e-10          # relative residual of the final factor's eigenpairs
        # final factor by Lanczos on a Chebyshev-filtered operator (lanczos.chebyshev_filtered_lanczos): "auto" = from
        # EIG_FILTER_MIN_NODES nodes on, True / False = always / never; the plain solver is the fallback either way
        self.eig_filter = "auto"
        self.eig_filter_degree = 8
        self.row_shard = bool(row_shard)
        self.shard_group = shard_group
        self.warm_start = bool(warm_start)
        self.last_solver = None

    def run_with_state(self, bs_iteration, Z, state):
        tic = self._get_tic()
        ret = self._run(Z, state)
        tim = self._get_tim(tic)
        K = state[0].shape[0]
        self._add_np_log("mmw_all_it", bs_iteration, np.array([Z, K, tim]))
        return ret

    # ------------------------------------------------------------------ internals
    def _dtype_code(self):
        if self.dtype in ("float64", "f64", np.float64):
            return _lib.F64
        if self.dtype in ("float32", "f32", np.float32):
            return _lib.F32
        raise ValueError("dtype must be 'float64' or 'float32'")

    def _matmat(self, solver, torch, dev):
        n = solver.plan.n

        def mm(X):
            X = X.contiguous()
            Y = torch.empty_like(X)
            solver.symv(X.data_ptr(), Y.data_ptr(), X.shape[0], torch.cuda.current_stream().cuda_stream)
            return Y
        return mm

    @staticmethod
    def _native_steps(solver, torch):
        def steps(Q, m, j0, j1, al, be):
            solver.lanczos_steps(Q.data_ptr(), m, j0, j1, al.data_ptr(), be.data_ptr(),
                                 torch.cuda.current_stream().cuda_stream)
        return steps

    def _gap_row(self, solver, torch, dev):
        """mmw.py:79-117 for the state at the start of the next iteration."""
        e_max = solver.gap_prepare(torch.cuda.current_stream().cuda_stream)
        n = solver.plan.n
        g = torch.Generator(device="cpu").manual_seed(12345)
        v0 = torch.randn(n, dtype=torch.float64, generator=g).to(dev)
        lam, _, _ = thick_restart_lanczos(self._matmat(solver, torch, dev), n, 1, "SA", v0, ncv=40, tol=1e-10,
                                          native_steps=self._native_steps(solver, torch))
        lam_min = float(lam[0]) * n
        return np.array([e_max, lam_min, e_max - lam_min])

    def _final_factor(self, solver, Z, nit, torch, dev, stream, seed_offset=0):
        """X_half = U sqrt(|Lambda|) of the top-r |lambda| eigenpairs of X_avgd / nit
        (mmw.py:202-216, where svds does it), rows in the caller's numbering."""
        plan = solver.plan
        lap = _Laps(torch)
        K = plan.n
        rank = int(min(K - 1, (Z - 1) * self.rank_radio))
        solver.xavg_matrix(1.0 / nit, stream)
        lap("xavg_matrix")
        if self.omega == "numpy":
            v0 = torch.from_numpy(np.random.standard_normal(K)).to(dev)   # svds' ARPACK start vector
        else:
            g = torch.Generator(device=dev).manual_seed(int(self.seed) + 1 + seed_offset)
            v0 = torch.randn(K, dtype=torch.float64, generator=g, device=dev)
        perm = torch.from_numpy(plan.perm()).to(dev).long()
        lap("start vector, perm")
        out = None
        if self.eig_filter is True or (self.eig_filter == "auto" and K >= self.EIG_FILTER_MIN_NODES):
            out = chebyshev_filtered_lanczos(self._matmat(solver, torch, dev), K, rank, v0[perm], self._native_steps(solver, torch),
                                             solver.lanczos_filter, tol=self.eig_tol, degree=self.eig_filter_degree)
        if out is None:      # small graphs, or the filter declined (see chebyshev_filtered_lanczos): the plain solver
            out = thick_restart_lanczos(self._matmat(solver, torch, dev), K, rank, "LM", v0[perm], tol=self.eig_tol,
                                        native_steps=self._native_steps(solver, torch))
        lam, V, info = out
        lap("eigen-solver")
        self.last_eig_info = info
        X_half_int = V * torch.sqrt(lam.abs())[None, :]
        X_half = torch.empty_like(X_half_int)
        X_half[perm] = X_half_int                              # internal -> caller numbering
        self.last_singular_values = lam.abs().cpu().numpy()
        # device -> host through a pinned buffer (24 MB at cfg4: ~1 ms instead of ~4 from pageable memory)
        lap("scale, unpermute")
        host = torch.empty(X_half.shape, dtype=X_half.dtype, pin_memory=True)
        lap("pinned buffer")
        host.copy_(X_half, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        lap("factor to host")
        return host.numpy()      # (the array keeps the pinned block alive; it returns to torch's host cache afterwards)

    def run_many_with_states(self, Zs, states):
        """Monte-Carlo sweeps (the drivers' `for seed in range(REPEAT)` loops): every
        (Z, state) pair solved in ONE batched launch, one thread block per instance, Omega
        generated on device.  Returns the list of X_half factors (same as calling
        run_with_state on each with omega="device")."""
        import torch
        from .batch import BatchSolver
        if not torch.cuda.is_available():
            raise _lib.SigSdpError("no CUDA device: sig_sdp_mmw_b200 has no CPU fallback")
        tic = self._get_tic()
        dev = torch.device("cuda", self.device)
        nit = int(self.nit)
        keep_omega = self.omega
        self.omega = "device"
        try:
            with torch.cuda.device(dev):
                stream = torch.cuda.current_stream().cuda_stream
                bs = BatchSolver(states, Zs, self.eta, rank_radio=self.rank_radio, dtype=self.dtype,
                                 device=self.device, order=self.plan_order)
                bs.iterate(nit, seed=self.seed, stream=stream)
                torch.cuda.current_stream().synchronize()
                out = [self._final_factor(sol, Z, nit, torch, dev, stream, seed_offset=i)
                       for i, (sol, Z) in enumerate(zip(bs.solvers, bs.Zs))]
        finally:
            self.omega = keep_omega
        self.last_batch = bs
        self._add_np_log("mmw_batch", 0, np.array([len(states), nit, self._get_tim(tic)]))
        return out

    def _run(self, Z, state):
        import torch
        if not torch.cuda.is_available():
            raise _lib.SigSdpError("no CUDA device: sig_sdp_mmw_b200 has no CPU fallback")
        sp_tic = self._get_tic()
        K = state[0].shape[0]
        D = Z * self.rank_radio                                   # mmw.py:180
        shard = None
        sharded = False
        if self.row_shard:
            import torch.distributed as dist
            if not (dist.is_available() and dist.is_initialized()):
                raise _lib.SigSdpError("row_shard=True needs an initialised torch.distributed process group (one process per GPU)")
            sharded = dist.get_world_size(self.shard_group) > 1       # a one-rank group is just the plain solver
        plan = self._plan_for(state, collective=sharded, group=self.shard_group)
        if sharded:
            from .rowshard import RowShardRank
            if self.LOG_GAP:
                raise _lib.SigSdpError("LOG_GAP is not available on a row-sharded solve")
            shard = RowShardRank(plan, Z, D, self.eta, dtype=self._dtype_code(), group=self.shard_group)
            solver = shard.solver
        else:
            solver = _lib.Solver(plan, Z, D, self.eta, self._dtype_code(), self.mode)
            prev = self.last_solver
            if self.warm_start and prev is not None and prev.plan is plan and prev.rows is None and prev.D_total == prev.D:
                solver.warm_start(prev)
        self.last_solver = solver
        dev = torch.device("cuda", plan.device)
        self._add_np_log("mmw_state_process", 0, np.array([Z, K, self._get_tim(sp_tic)]))

        nit = int(self.nit)
        self.N_STEP = 0
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream().cuda_stream
            step = 1 if self.LOG_GAP else nit
            if self.omega == "numpy" and not self.LOG_GAP:
                step = max(1, min(nit, _OMEGA_CHUNK_BYTES // max(1, K * D * 8)))
            done = 0
            ring, slot = None, 0
            wall_tic = self._get_tic()
            while done < nit:
                cnt = min(step, nit - done)
                if self.LOG_GAP:
                    self._add_np_log("gap", done, self._gap_row(solver, torch, dev))
                if self.omega == "numpy":
                    # the reference's stream: one randn(K, D) per iteration (mmw.py:226), continued natively on all host
                    # cores (bit for bit numpy's numbers, numpy's global state advanced as numpy would: _lib.numpy_randn_into)
                    # into one of two pinned buffers, so the draw of the next chunk runs while this one is copied and
                    # iterated on
                    if ring is None:
                        ring = [(torch.empty((step, K, D), dtype=torch.float64, pin_memory=True),
                                 torch.empty((step, K, D), dtype=torch.float64, device=dev), torch.cuda.Event()) for _ in range(2)]
                    host, om_d, ev = ring[slot]
                    slot ^= 1
                    ev.synchronize()                               # the copy that last read this pinned buffer is done
                    _lib.numpy_randn_into(host.numpy()[:cnt])
                    om_d[:cnt].copy_(host[:cnt], non_blocking=True)
                    ev.record()
                    solver.iterate(cnt, om_d.data_ptr(), 0, stream)  # (stream order keeps om_d alive until its kernel is done)
                elif self.omega == "device":
                    solver.iterate(cnt, None, self.seed, stream)
                else:
                    raise ValueError("omega must be 'numpy' or 'device'")
                done += cnt
                self.N_STEP = done
            torch.cuda.current_stream().synchronize()
            ring = None
            wall_us = self._get_tim(wall_tic)

            # per-iteration phase logs, device-timed (mmw.py:142,170,197,200)
            cnt = min(nit, 8192)
            pt = solver.phase_times(cnt) if self.mode == _lib.MODE_FUSED else np.zeros((cnt, 4))
            if not pt.any():
                pt = np.full((cnt, 4), wall_us / max(nit, 1) / 4.0)
            its = np.arange(nit - cnt, nit)
            zk = np.tile(np.array([float(Z), float(K)]), (cnt, 1))
            self._add_np_log_rows("mmw_dual", its, np.column_stack((zk, pt[:, 0])))
            self._add_np_log_rows("mmw_loss", its, np.column_stack((zk, pt[:, 1])))
            self._add_np_log_rows("mmw_expm", its, np.column_stack((zk, pt[:, 2] + pt[:, 3])))
            self._add_np_log_rows("mmw_per_it", its, np.column_stack((zk, pt.sum(axis=1))))

            # final factor (mmw.py:202-216): top-r |lambda| eigenpairs of X_avgd / nit
            tic_xavg = self._get_tic()
            if shard is not None:
                shard.complete_averages()      # every rank now holds the whole running sum
            X_half = self._final_factor(solver, Z, nit, torch, dev, stream)
            self._add_np_log("mmw_xavg", 0, np.array([Z, K, self._get_tim(tic_xavg)]))
        return True, X_half
