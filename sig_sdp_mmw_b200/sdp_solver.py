"""Solver base class and randomized rounding with the reference's surface
(sim_src/alg/sdp_solver.py:9-107).  The arithmetic runs in libsigsdp_mmw.so:

  * sigsdp_round_project  (device): randv gX^T, per-user slot preference order, ||gX_k||
  * sigsdp_round_greedy_device (device): the sequential feasibility assignment, reproduced exactly by
    rounds of mutually non-interacting users (sigsdp_round_greedy: the same pass sequentially on the host)
  * sigsdp_round_conflicts(device): same-slot interference / violation / asso-conflict counts

The random draws (np.random.randn for the directions, np.random.randint for the
leftovers) stay on numpy's global stream exactly where the reference draws them
(sdp_solver.py:48,105), so a seeded run reproduces the reference's colouring."""
import ctypes as C

import numpy as np

from . import _lib


def _digest_vec(v):
    """(length, sigsdp_checksum) of an array's bytes: a position-weighted 64-bit checksum computed on the plan
    builder's own host threads (~0.5 ms for the 48 MB of a 100k-node state, against 4.4 ms for a single-threaded
    numpy reduction; a threaded BLAS dot would leave its workers spinning on the cores the builder needs next)."""
    v = np.ascontiguousarray(v)
    return (int(v.size), _lib.checksum(v))


def _digest(M):
    return (int(M.nnz),) + _digest_vec(M.data)[1:] + _digest_vec(M.indices)[1:] + _digest_vec(M.indptr)[1:]


class sdp_solver:
    def __init__(self, nit=100, rank_radio=2, alpha=1.):
        self.nit = nit
        self.rank_radio = rank_radio
        self.alpha = alpha  # unused, as in the reference (sdp_solver.py:13)

    # set by subclasses / callers
    device = 0
    plan_order = 0
    # the greedy pass of the rounding: "device" (rounds of mutually non-interacting users, sigsdp_round_greedy_device),
    # "host" (the same pass sequentially in native code, sigsdp_round_greedy) or "auto": both give the same colours
    # bit for bit; the device wins on large graphs (100k nodes: 6.6 ms against 30 ms), the host on small ones where
    # the number of rounds (82-206), not the work, sets the time
    greedy = "auto"
    GREEDY_DEVICE_MIN_NODES = 50000

    def run_with_state(self, bs_iteration, Z, state):
        pass

    # ---- plan cache: the graph plan is Z-independent and reused across the binary search
    def _plan_for(self, state, collective=False, group=None):
        """The plan of `state`, rebuilt when the state's content changes.  The key is a position-
        weighted checksum of every buffer (values AND structure: an in-place permutation of the
        values, or a changed pattern with the same nnz, changes it), about 0.5 ms at 100k nodes
        against a ~50 ms plan build; the keyed objects are kept alive so id() cannot be reused."""
        S, Q, h = state
        key = (id(S), id(Q), S.shape, self.device, self.plan_order) + _digest(S) + _digest(Q) + _digest_vec(np.asarray(h))
        cache = self.__dict__.setdefault("_plan_cache", {})
        if cache.get("key") != key:
            if collective:      # every rank of the job calls this with the same state (row-sharded solve)
                cache["plan"] = _lib.Plan.collective(state, self.device, self.plan_order, group)
            else:
                cache["plan"] = _lib.Plan(state, device=self.device, order=self.plan_order)
            cache["key"] = key
            cache["objects"] = (S, Q, h)
        return cache["plan"]

    def invalidate_plan(self):
        """Forget the cached graph plan (call after changing a state in place in a way the checksum
        cannot see, or to release its device memory)."""
        self.__dict__.pop("_plan_cache", None)

    def rounding(self, Z, gX, state, nattempt=10):
        z_vec = None
        remainder = None
        for n in range(nattempt):
            z_vec, Z, remainder = self.rounding_one_attempt(Z, gX, state)
            if remainder == 0:
                return z_vec, Z, remainder
        return z_vec, Z, remainder

    def rounding_one_attempt(self, Z, gX, state):
        import torch
        lib = _lib.load()
        plan = self._plan_for(state)
        K = state[0].shape[0]
        gX = np.ascontiguousarray(np.asarray(gX, dtype=np.float64))
        D = gX.shape[1]
        randv = np.random.randn(Z, D)                                  # sdp_solver.py:48
        randv = randv / np.linalg.norm(randv, axis=1, keepdims=True)   # :49
        dev = torch.device("cuda", plan.device)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream().cuda_stream
            gX_d = torch.from_numpy(gX).to(dev)
            rv_d = torch.from_numpy(np.ascontiguousarray(randv)).to(dev)
            pref_d = torch.empty((K, Z), dtype=torch.int32, device=dev)
            norm_d = torch.empty(K, dtype=torch.float64, device=dev)
            _lib.check(lib.sigsdp_round_project(plan.handle, gX_d.data_ptr(), D, rv_d.data_ptr(), Z,
                                                pref_d.data_ptr(), norm_d.data_ptr(), stream))
            # visit order argsort(-||gX_k||) (:52); stable so equal norms keep index order
            rank_d = torch.argsort(-norm_d, stable=True).to(torch.int32)
            rem = C.c_int64()
            on_device = self.greedy == "device" or (self.greedy == "auto" and K >= self.GREEDY_DEVICE_MIN_NODES)
            if on_device:
                # the sequential feasibility pass (:70-101) as rounds of mutually non-interacting users on the
                # device: same result, no n x Z preference table over PCIe (sigsdp_round_greedy_device)
                z_d = torch.empty(K, dtype=torch.int32, device=dev)
                rounds = C.c_int64()
                _lib.check(lib.sigsdp_round_greedy_device(plan.handle, Z, rank_d.data_ptr(), pref_d.data_ptr(), z_d.data_ptr(),
                                                          C.byref(rem), C.byref(rounds), stream))
                z_int = z_d.cpu().numpy()
                self.last_greedy_rounds = int(rounds.value)
            else:
                rank = rank_d.cpu().numpy()
                pref = pref_d.cpu().numpy()
        if not on_device:
            z_int = np.empty(K, np.int32)
            S, Q, h = plan._S, plan._Q, plan._h
            _lib.check(lib.sigsdp_round_greedy(K, Z, _lib._p(S[0], C.c_int32), _lib._p(S[1], C.c_int32), _lib._p(S[2], C.c_double),
                                               _lib._p(Q[0], C.c_int32), _lib._p(Q[1], C.c_int32), _lib._p(Q[2], C.c_double),
                                               _lib._p(h, C.c_double), _lib._p(rank, C.c_int32),
                                               _lib._p(np.ascontiguousarray(pref), C.c_int32), _lib._p(z_int, C.c_int32),
                                               C.byref(rem)))
        z_vec = z_int.astype(np.float64)
        not_assigned = z_int < 0
        z_vec[not_assigned] = 0.0
        remainder = int(rem.value)
        if remainder:
            z_vec[not_assigned] = np.random.randint(Z, size=remainder)  # :104-105
        return z_vec, Z, remainder

    def conflict_counts(self, z_vec, state, return_interference=False):
        """rounding.py:56-66 on device: (#users whose same-slot interference exceeds h_max,
        #association pairs sharing a slot[, I])."""
        import torch
        plan = self._plan_for(state)
        dev = torch.device("cuda", plan.device)
        with torch.cuda.device(dev):
            z_d = torch.from_numpy(np.ascontiguousarray(np.asarray(z_vec).astype(np.int32))).to(dev)
            I_d = torch.empty(plan.n, dtype=torch.float64, device=dev) if return_interference else None
            counts = (C.c_int64 * 2)()
            _lib.check(_lib.load().sigsdp_round_conflicts(plan.handle, z_d.data_ptr(),
                                                          I_d.data_ptr() if I_d is not None else None, counts,
                                                          torch.cuda.current_stream().cuda_stream))
        if return_interference:
            return int(counts[0]), int(counts[1]), I_d.cpu().numpy()
        return int(counts[0]), int(counts[1])

    def argmax_colours(self, Z, gX, randv_raw=None):
        """rounding.py:46-54 on the solver's factor: colour_k = argmax_z <randv_z, g_k>, the
        first preference of rounding_one_attempt."""
        import torch
        lib = _lib.load()
        gX = np.ascontiguousarray(np.asarray(gX, dtype=np.float64))
        K, D = gX.shape
        if randv_raw is None:
            randv_raw = np.random.randn(Z, D)
        randv = randv_raw / np.linalg.norm(randv_raw, axis=1, keepdims=True)
        plan = self.__dict__.get("_plan_cache", {}).get("plan")
        if plan is None or plan.n != K:
            raise RuntimeError("argmax_colours needs a plan: call run_with_state/rounding on the state first")
        dev = torch.device("cuda", plan.device)
        with torch.cuda.device(dev):
            gX_d = torch.from_numpy(gX).to(dev)
            rv_d = torch.from_numpy(np.ascontiguousarray(randv)).to(dev)
            pref_d = torch.empty((K, Z), dtype=torch.int32, device=dev)
            norm_d = torch.empty(K, dtype=torch.float64, device=dev)
            _lib.check(lib.sigsdp_round_project(plan.handle, gX_d.data_ptr(), D, rv_d.data_ptr(), Z,
                                                pref_d.data_ptr(), norm_d.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream))
            return pref_d[:, 0].cpu().numpy().astype(np.int64)


class rand_sdp_solver(sdp_solver):
    """The reference's random baseline (sdp_solver.py:109-114): unit-norm random rows instead of an
    SDP factor, rounded by the same greedy pass.  The drivers use it as the "rand" arm."""

    def run_with_state(self, bs_iteration, Z, state):
        K = state[0].shape[0]
        randv = np.random.randn(K, Z * self.rank_radio)
        return True, randv / np.linalg.norm(randv, axis=1, keepdims=True)
