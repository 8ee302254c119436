"""ctypes binding of libsigsdp_mmw.so (include/sigsdp_mmw.h).  There is no CPU
fallback: if the library is missing or a call fails, this raises."""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SIGSDP_LIB") or os.path.join(_HERE, "libsigsdp_mmw.so")

F64, F32 = 0, 1
ARR_X_AVGD, ARR_X, ARR_Y_AVGD, ARR_Y = 1, 2, 3, 4
MODE_FUSED, MODE_STEPWISE = 0, 1

_lib = None


class SigSdpError(RuntimeError):
    pass


def _p(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype)) if arr is not None else None


def load():
    """Load the shared library (build it first with ``python -m sig_sdp_mmw_b200.build``
    or ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SigSdpError("%s not found: build it with `python -m sig_sdp_mmw_b200.build` "
                          "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    i32p, i64p, f64p, vp = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double), C.c_void_p
    lib.sigsdp_last_error.restype = C.c_char_p
    lib.sigsdp_version.restype = C.c_int
    lib.sigsdp_device_count.restype = C.c_int
    sigs = {
        "sigsdp_plan_create": [C.c_int64, i32p, i32p, f64p, i32p, i32p, f64p, f64p, C.c_int, C.c_int, C.POINTER(vp)],
        "sigsdp_plan_builds_on_device": [C.c_int64, C.c_int],
        "sigsdp_plan_image_size": [vp, i64p],
        "sigsdp_plan_image": [vp, vp],
        "sigsdp_plan_create_from_image": [C.c_int64, i32p, i32p, f64p, i32p, i32p, f64p, f64p, vp, C.c_int64, C.c_int, C.POINTER(vp)],
        "sigsdp_plan_info": [vp, i64p],
        "sigsdp_checksum": [vp, C.c_int64, C.POINTER(C.c_uint64)],
        "sigsdp_numpy_standard_normal": [C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double),
                                         C.c_int64, f64p],
        "sigsdp_plan_edges": [vp, i32p, i32p, f64p, f64p, i32p, i32p],
        "sigsdp_plan_vectors": [vp, f64p, f64p],
        "sigsdp_plan_perm": [vp, i32p],
        "sigsdp_solver_create": [vp, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(vp)],
        "sigsdp_solver_create_tiled": [vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.POINTER(vp)],
        "sigsdp_solver_create_sharded": [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.POINTER(vp)],
        "sigsdp_solver_split_step": [vp, C.c_int, vp, C.c_uint64, vp],
        "sigsdp_solver_exchange_buffer": [vp, C.POINTER(vp), i64p],
        "sigsdp_solver_device_array": [vp, C.c_int, C.POINTER(vp), i64p],
        "sigsdp_solver_create_rows": [vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)],
        "sigsdp_solver_shard_info": [vp, i64p],
        "sigsdp_plan_row_partition": [vp, C.c_int, C.c_int, C.c_int, C.c_int, i64p, i64p, i64p, i64p],
        "sigsdp_solver_shard_arena": [vp, C.POINTER(vp), i64p],
        "sigsdp_solver_shard_ipc_handle": [vp, vp],
        "sigsdp_solver_shard_attach_ipc": [vp, vp],
        "sigsdp_solver_shard_attach_local": [C.POINTER(vp), C.c_int],
        "sigsdp_solver_set_X": [vp, C.c_int, f64p, f64p, f64p],
        "sigsdp_solver_reset": [vp, vp],
        "sigsdp_solver_warm_start": [vp, vp, vp],
        "sigsdp_solver_set_mode": [vp, C.c_int],
        "sigsdp_solver_info": [vp, i64p],
        "sigsdp_solver_iterate": [vp, C.c_int, vp, C.c_uint64, vp],
        "sigsdp_solver_get_dual": [vp, f64p, f64p, f64p],
        "sigsdp_solver_get_X": [vp, C.c_int, f64p, f64p, f64p],
        "sigsdp_solver_get_L": [vp, f64p, f64p, f64p],
        "sigsdp_solver_get_sketch": [vp, f64p],
        "sigsdp_solver_get_history": [vp, C.c_int, i32p, i32p, i32p, f64p, f64p],
        "sigsdp_solver_total_terms": [vp, i64p],
        "sigsdp_solver_debug_cycles": [vp, i64p],
        "sigsdp_solver_get_phase_times": [vp, C.c_int, f64p],
        "sigsdp_solver_xavg_matrix": [vp, C.c_double, vp],
        "sigsdp_solver_gap_prepare": [vp, f64p, vp],
        "sigsdp_solver_symv": [vp, vp, vp, C.c_int, vp],
        "sigsdp_solver_get_matrix": [vp, f64p],
        "sigsdp_solver_lanczos_steps": [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp],
        "sigsdp_solver_lanczos_filter": [vp, C.c_int, C.c_double, C.c_double],
        "sigsdp_plan_pattern": [vp, i32p, i32p],
        "sigsdp_plan_tile_stats": [vp, C.c_int, C.c_int, C.c_int, i64p],
        "sigsdp_debug_normals": [C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_int, f64p],
        "sigsdp_batch_create": [C.POINTER(vp), C.c_int, C.POINTER(vp)],
        "sigsdp_batch_create_ids": [C.POINTER(vp), i64p, C.c_int, C.POINTER(vp)],
        "sigsdp_batch_iterate": [vp, C.c_int, C.c_uint64, vp],
        "sigsdp_round_project": [vp, vp, C.c_int, vp, C.c_int, vp, vp, vp],
        "sigsdp_round_greedy": [C.c_int64, C.c_int, i32p, i32p, f64p, i32p, i32p, f64p, f64p, i32p, i32p, i32p, i64p],
        "sigsdp_round_conflicts": [vp, vp, vp, i64p, vp],
        "sigsdp_round_greedy_device": [vp, C.c_int, vp, vp, vp, i64p, i64p, vp],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.sigsdp_plan_destroy.argtypes = [vp]
    lib.sigsdp_plan_destroy.restype = None
    lib.sigsdp_solver_destroy.argtypes = [vp]
    lib.sigsdp_solver_destroy.restype = None
    lib.sigsdp_batch_destroy.argtypes = [vp]
    lib.sigsdp_batch_destroy.restype = None
    lib.sigsdp_batch_blocks_per_instance.argtypes = [vp]
    lib.sigsdp_batch_blocks_per_instance.restype = C.c_int
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise SigSdpError("libsigsdp_mmw error %d: %s" % (rc, load().sigsdp_last_error().decode()))


def checksum(arr):
    """sigsdp_checksum of a contiguous numpy array's bytes."""
    arr = np.ascontiguousarray(arr)
    out = C.c_uint64()
    check(load().sigsdp_checksum(arr.ctypes.data_as(C.c_void_p), int(arr.nbytes), C.byref(out)))
    return int(out.value)


_randn_lock = threading.Lock()


def numpy_randn_into(out):
    """Fills the float64 array `out` (C order) with the numbers np.random.standard_normal(out.shape) would return next
    on numpy's GLOBAL legacy stream and leaves that stream where numpy would have left it -- bit for bit, on all host
    cores (sigsdp_numpy_standard_normal).  Falls back to numpy itself when the global generator is not MT19937."""
    if out.dtype != np.float64 or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous float64 array")
    with _randn_lock:      # get_state ... set_state is one draw: two threads of this process must not interleave there
        st = np.random.get_state()
        if st[0] != "MT19937" or out.size == 0:
            out[...] = np.random.standard_normal(out.shape)
            return out
        key = np.ascontiguousarray(st[1], dtype=np.uint32).copy()
        pos, hg, g = C.c_int32(int(st[2])), C.c_int32(int(st[3])), C.c_double(float(st[4]))
        check(load().sigsdp_numpy_standard_normal(key.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(pos), C.byref(hg), C.byref(g),
                                                  int(out.size), out.ctypes.data_as(C.POINTER(C.c_double))))
        np.random.set_state(("MT19937", key, pos.value, hg.value, g.value))
    return out


def csr_arrays(M, canonicalize=True):
    """(indptr int32, indices int32, data float64) of a scipy matrix as canonical CSR
    (sorted, duplicates summed) without touching the caller's object.  canonicalize=False skips
    scipy's O(nnz) sortedness scan: the library validates the structure itself and reports
    unsorted / duplicate indices, upon which the caller retries with canonicalize=True."""
    import scipy.sparse as sp
    if not (sp.isspmatrix_csr(M) or isinstance(M, getattr(sp, "csr_array", ()))):
        M = sp.csr_matrix(M)
        canonicalize = True
    if canonicalize and not M.has_canonical_format:
        M = M.copy()
        M.sum_duplicates()
    return (np.ascontiguousarray(M.indptr, dtype=np.int32), np.ascontiguousarray(M.indices, dtype=np.int32),
            np.ascontiguousarray(M.data, dtype=np.float64))


class Plan:
    """Graph plan (Z-independent), see sigsdp_plan_create."""

    def __init__(self, state, device=0, order=0, image=None):
        """image: a plan image (Plan.image()) of the same state built elsewhere: skips the host build."""
        lib = load()
        S, Q, h = state
        n = S.shape[0]
        if S.shape != (n, n) or Q.shape != (n, n) or np.asarray(h).shape != (n,):
            raise ValueError("state must be (S_gain n x n, Q_asso n x n, h_max (n,))")
        self._h = np.ascontiguousarray(np.asarray(h, dtype=np.float64))
        self.handle = C.c_void_p()
        if image is not None:
            self._S = csr_arrays(S, True)
            self._Q = csr_arrays(Q, True)
            img = np.ascontiguousarray(image, dtype=np.uint8)
            check(lib.sigsdp_plan_create_from_image(n, _p(self._S[0], C.c_int32), _p(self._S[1], C.c_int32), _p(self._S[2], C.c_double),
                                                    _p(self._Q[0], C.c_int32), _p(self._Q[1], C.c_int32), _p(self._Q[2], C.c_double),
                                                    _p(self._h, C.c_double), img.ctypes.data_as(C.c_void_p), img.size, device,
                                                    C.byref(self.handle)))
            self._read_info()
            return
        for canonicalize in (False, True):
            self._S = csr_arrays(S, canonicalize)
            self._Q = csr_arrays(Q, canonicalize)
            rc = lib.sigsdp_plan_create(n, _p(self._S[0], C.c_int32), _p(self._S[1], C.c_int32), _p(self._S[2], C.c_double),
                                        _p(self._Q[0], C.c_int32), _p(self._Q[1], C.c_int32), _p(self._Q[2], C.c_double),
                                        _p(self._h, C.c_double), device, order, C.byref(self.handle))
            if rc == 0 or canonicalize or b"sorted and duplicate-free" not in lib.sigsdp_last_error():
                break      # (non-canonical input: second pass on a canonicalised copy)
        check(rc)
        self._read_info()

    def _read_info(self):
        info = (C.c_int64 * 8)()
        check(load().sigsdp_plan_info(self.handle, info))
        self.n, self.E_g, self.E_a, self.nnz, self.nnzT, self.device, self.order, self.max_row = [int(x) for x in info]

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                load().sigsdp_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def image(self):
        """The built plan as a flat uint8 array (sigsdp_plan_image), for Plan(state, image=...) in another process."""
        nb = C.c_int64()
        check(load().sigsdp_plan_image_size(self.handle, C.byref(nb)))
        buf = np.empty(int(nb.value), np.uint8)
        check(load().sigsdp_plan_image(self.handle, buf.ctypes.data_as(C.c_void_p)))
        return buf

    @staticmethod
    def collective(state, device, order=1, group=None, min_world=4):
        """One plan per rank of a torch.distributed job on one box.  With few ranks every rank builds its own (the
        cores are shared: 8 each at 2 ranks); from `min_world` ranks on, rank 0 builds it ONCE on all host cores and
        the image travels by broadcast (NCCL: pinned host -> device -> peers -> pinned host), which the other ranks
        import -- instead of 8 builds side by side on 2 cores each.  A graph large enough for the device-side builder
        (sigsdp_plan_builds_on_device) is always built per rank: each GPU builds its own copy, the host cores only
        stage the inputs and run the locality ordering (one thread per rank)."""
        import torch
        import torch.distributed as dist
        world = dist.get_world_size(group)
        if world < min_world or load().sigsdp_plan_builds_on_device(int(state[0].shape[0]), int(device)) == 1:
            return Plan(state, device=device, order=order)
        rank = dist.get_rank(group)
        nccl = dist.get_backend(group) == "nccl"
        dev = torch.device("cuda", device) if nccl else torch.device("cpu")
        src = dist.get_global_rank(group, 0) if group is not None else 0
        lib = load()
        if rank == 0:
            keep = os.environ.pop("LOCAL_WORLD_SIZE", None)     # the other ranks wait: rank 0 may use every core
            try:
                plan = Plan(state, device=device, order=order)
            finally:
                if keep is not None:
                    os.environ["LOCAL_WORLD_SIZE"] = keep
            nb = C.c_int64()
            check(lib.sigsdp_plan_image_size(plan.handle, C.byref(nb)))
            size = torch.tensor([int(nb.value)], dtype=torch.int64, device=dev)
            dist.broadcast(size, src=src, group=group)
            img = torch.empty(int(nb.value), dtype=torch.uint8, pin_memory=nccl)
            check(lib.sigsdp_plan_image(plan.handle, C.c_void_p(img.data_ptr())))
            dist.broadcast(img.to(dev, non_blocking=True) if nccl else img, src=src, group=group)
            return plan
        size = torch.zeros(1, dtype=torch.int64, device=dev)
        dist.broadcast(size, src=src, group=group)
        img = torch.empty(int(size.item()), dtype=torch.uint8, device=dev)
        dist.broadcast(img, src=src, group=group)
        if nccl:
            host = torch.empty(img.numel(), dtype=torch.uint8, pin_memory=True)
            host.copy_(img)
            torch.cuda.current_stream().synchronize()
            img = host
        return Plan(state, device=device, order=order, image=img.numpy())

    def edges(self):
        gi = np.empty(self.E_g, np.int32); gj = np.empty(self.E_g, np.int32)
        tij = np.empty(self.E_g); tji = np.empty(self.E_g)
        ai = np.empty(self.E_a, np.int32); aj = np.empty(self.E_a, np.int32)
        check(load().sigsdp_plan_edges(self.handle, _p(gi, C.c_int32), _p(gj, C.c_int32), _p(tij, C.c_double),
                                       _p(tji, C.c_double), _p(ai, C.c_int32), _p(aj, C.c_int32)))
        return gi, gj, tij, tji, ai, aj

    def vectors(self):
        a = np.empty(self.n); b = np.empty(self.n)
        check(load().sigsdp_plan_vectors(self.handle, _p(a, C.c_double), _p(b, C.c_double)))
        return a, b

    def perm(self):
        p = np.empty(self.n, np.int32)
        check(load().sigsdp_plan_perm(self.handle, _p(p, C.c_int32)))
        return p

    def tile_stats(self, max_rows, ucap, nnzcap):
        """Row-tile statistics for the given caps (host only): dict of tiles, runs (bulk copies),
        staged rows, the largest tile's staged rows / non-zeros, nnz."""
        out = np.zeros(6, np.int64)
        check(load().sigsdp_plan_tile_stats(self.handle, max_rows, ucap, nnzcap, _p(out, C.c_int64)))
        return dict(zip(["tiles", "runs", "staged_rows", "umax", "nnzmax", "nnz"], out.tolist()))

    def row_partition(self, nranks, max_rows=0, ucap=0, nnzcap=0):
        """The row partition of a row-sharded solver (host only): dict of row0 (nranks + 1), and per rank the
        rows pushed per Taylor term, foreign rows read, association edges owned."""
        row0 = np.zeros(nranks + 1, np.int64); send = np.zeros(nranks, np.int64)
        recv = np.zeros(nranks, np.int64); owned = np.zeros(nranks, np.int64)
        check(load().sigsdp_plan_row_partition(self.handle, int(nranks), int(max_rows), int(ucap), int(nnzcap), _p(row0, C.c_int64),
                                               _p(send, C.c_int64), _p(recv, C.c_int64), _p(owned, C.c_int64)))
        return dict(row0=row0, send=send, recv=recv, owned_asso=owned)

    def pattern(self):
        rp = np.empty(self.n + 1, np.int32); col = np.empty(self.nnz, np.int32)
        check(load().sigsdp_plan_pattern(self.handle, _p(rp, C.c_int32), _p(col, C.c_int32)))
        return rp, col


class Solver:
    """MMW state for one (plan, Z, D, eta, dtype), see sigsdp_solver_create."""

    def __init__(self, plan, Z, D, eta, dtype=F64, mode=MODE_FUSED, tiling=-1, D_total=None, col0=0,
                 rows=None, max_blocks=0):
        """rows=(rank, nranks): a row shard (sigsdp_solver_create_rows); attach its peers before
        iterating (attach_local / ipc_handle + attach_ipc)."""
        lib = load()
        self.plan = plan
        self.handle = C.c_void_p()
        if D_total is None:
            D_total = D
        self.rows = tuple(int(x) for x in rows) if rows is not None else None
        if self.rows is not None:
            check(lib.sigsdp_solver_create_rows(plan.handle, int(Z), int(D), float(eta), int(dtype), int(tiling),
                                                self.rows[0], self.rows[1], int(max_blocks), C.byref(self.handle)))
        else:
            check(lib.sigsdp_solver_create_sharded(plan.handle, int(Z), int(D_total), int(col0), int(D), float(eta),
                                                   int(dtype), int(tiling), C.byref(self.handle)))
        self.D_total, self.col0 = int(D_total), int(col0)
        if mode != MODE_FUSED:
            check(lib.sigsdp_solver_set_mode(self.handle, mode))
        self.Z, self.D = int(Z), int(D)
        i = self.info()
        self.Dp, self.C, self.grid, self.threads, self.lanes = i["Dp"], i["C"], i["grid"], i["threads"], i["lanes"]
        self.tile_rows, self.smem = i["tile_rows"], i["smem"]

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                load().sigsdp_solver_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def info(self):
        a = (C.c_int64 * 12)()
        check(load().sigsdp_solver_info(self.handle, a))
        keys = ["n", "Z", "D", "Dp", "C", "iters", "dtype", "grid", "threads", "lanes", "tile_rows", "smem"]
        return dict(zip(keys, [int(x) for x in a]))

    def reset(self, stream=None):
        check(load().sigsdp_solver_reset(self.handle, stream))

    def iterate(self, n_iters, omega_dev_ptr=None, seed=0, stream=None):
        check(load().sigsdp_solver_iterate(self.handle, int(n_iters), omega_dev_ptr, int(seed), stream))

    def warm_start(self, src, stream=None):
        """Initial dual state (e_accu, Y, soft-max shift) from another solver of the same plan."""
        check(load().sigsdp_solver_warm_start(self.handle, src.handle, stream))

    def split_step(self, do_iter=True, omega_dev_ptr=None, seed=0, stream=None):
        check(load().sigsdp_solver_split_step(self.handle, int(bool(do_iter)), omega_dev_ptr, int(seed), stream))

    # ---- row shards
    def shard_info(self):
        a = (C.c_int64 * 12)()
        check(load().sigsdp_solver_shard_info(self.handle, a))
        keys = ["rank", "nranks", "row_lo", "row_hi", "tile_lo", "tile_hi", "halo_send_rows", "halo_recv_rows",
                "arena_bytes", "n_inc", "n_inc_owned", "attached"]
        return dict(zip(keys, [int(x) for x in a]))

    def ipc_handle(self):
        buf = (C.c_ubyte * 64)()
        check(load().sigsdp_solver_shard_ipc_handle(self.handle, buf))
        return bytes(buf)

    def attach_ipc(self, handles):
        """handles: the ranks' 64-byte IPC handles in rank order (this rank's own entry is ignored)."""
        blob = b"".join(handles)
        assert len(blob) == 64 * self.rows[1]
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        check(load().sigsdp_solver_shard_attach_ipc(self.handle, buf))

    @staticmethod
    def attach_local(shards):
        arr = (C.c_void_p * len(shards))(*[s.handle for s in shards])
        check(load().sigsdp_solver_shard_attach_local(arr, len(shards)))

    def set_X(self, averaged, diag, gain, asso):
        d = np.ascontiguousarray(diag, dtype=np.float64); g = np.ascontiguousarray(gain, dtype=np.float64)
        a = np.ascontiguousarray(asso, dtype=np.float64)
        assert d.size == self.plan.n and g.size == self.plan.E_g and a.size == self.plan.E_a
        check(load().sigsdp_solver_set_X(self.handle, int(averaged), _p(d, C.c_double), _p(g, C.c_double), _p(a, C.c_double)))

    def exchange_buffer(self):
        """(device pointer, number of doubles) of the per-iteration all-reduce buffer."""
        ptr = C.c_void_p()
        cnt = C.c_int64()
        check(load().sigsdp_solver_exchange_buffer(self.handle, C.byref(ptr), C.byref(cnt)))
        return int(ptr.value), int(cnt.value)

    def device_array(self, which):
        """(device pointer, number of doubles) of a state array in the internal layout (ARR_*)."""
        ptr = C.c_void_p()
        cnt = C.c_int64()
        check(load().sigsdp_solver_device_array(self.handle, int(which), C.byref(ptr), C.byref(cnt)))
        return int(ptr.value), int(cnt.value)

    def dual(self):
        Y = np.empty(self.C); e = np.empty(self.C); Yb = np.empty(self.C)
        check(load().sigsdp_solver_get_dual(self.handle, _p(Y, C.c_double), _p(e, C.c_double), _p(Yb, C.c_double)))
        return Y, e, Yb

    def X(self, averaged=False):
        p = self.plan
        d = np.empty(p.n); g = np.empty(p.E_g); a = np.empty(p.E_a)
        check(load().sigsdp_solver_get_X(self.handle, int(averaged), _p(d, C.c_double), _p(g, C.c_double), _p(a, C.c_double)))
        return d, g, a

    def L(self):
        p = self.plan
        d = np.empty(p.n); g = np.empty(p.E_g); a = np.empty(p.E_a)
        check(load().sigsdp_solver_get_L(self.handle, _p(d, C.c_double), _p(g, C.c_double), _p(a, C.c_double)))
        return d, g, a

    def sketch(self):
        Yh = np.empty((self.plan.n, self.D))
        check(load().sigsdp_solver_get_sketch(self.handle, _p(Yh, C.c_double)))
        return Yh

    def history(self, count):
        m = np.empty(count, np.int32); s = np.empty(count, np.int32); nt = np.empty(count, np.int32)
        a1 = np.empty(count); mu = np.empty(count)
        check(load().sigsdp_solver_get_history(self.handle, count, _p(m, C.c_int32), _p(s, C.c_int32), _p(nt, C.c_int32),
                                               _p(a1, C.c_double), _p(mu, C.c_double)))
        # scipy's _fragment_3_1 only uses the ||A||_1 rule while condition (3.13) holds,
        # ||A||_1 <= 2 l p_max (p_max + 3) theta_55 / (55 n0) = 63.36 / D (_expm_multiply.py:519-531); beyond
        # it the reference estimates ||A^p||_1^(1/p) with a randomised norm estimator and may pick a cheaper
        # (m*, s).  The kernels always use the ||A||_1 rule (conservative: never fewer terms); `cond313`
        # says, per iteration, whether both sides provably took the same branch.
        return dict(m_star=m, s=s, nterms=nt, a1norm=a1, mu=mu, cond313=a1 <= 63.36 / self.D_total)

    def phase_times(self, count):
        t = np.empty((count, 4))
        check(load().sigsdp_solver_get_phase_times(self.handle, count, _p(t, C.c_double)))
        return t

    def xavg_matrix(self, scale, stream=None):
        check(load().sigsdp_solver_xavg_matrix(self.handle, float(scale), stream))

    def gap_prepare(self, stream=None):
        v = C.c_double()
        check(load().sigsdp_solver_gap_prepare(self.handle, C.byref(v), stream))
        return float(v.value)

    def symv(self, x_ptr, y_ptr, nvec=1, stream=None):
        check(load().sigsdp_solver_symv(self.handle, x_ptr, y_ptr, int(nvec), stream))

    def lanczos_steps(self, Q_ptr, m, j0, j1, al_ptr, be_ptr, stream=None):
        check(load().sigsdp_solver_lanczos_steps(self.handle, Q_ptr, int(m), int(j0), int(j1), al_ptr, be_ptr, stream))

    def lanczos_filter(self, degree, lo=0.0, cut=1.0):
        """Chebyshev filter of the operator behind lanczos_steps (degree < 2: off), see sigsdp_solver_lanczos_filter."""
        check(load().sigsdp_solver_lanczos_filter(self.handle, int(degree), float(lo), float(cut)))

    def matrix_values(self):
        v = np.empty(self.plan.nnz)
        check(load().sigsdp_solver_get_matrix(self.handle, _p(v, C.c_double)))
        return v

    def sync_wait_ns(self):
        """Nanoseconds the fused kernel's leader thread spent in team barriers since reset."""
        a = (C.c_int64 * 8)()
        check(load().sigsdp_solver_debug_cycles(self.handle, a))
        return int(a[4])

    def barrier_breakdown_ns(self):
        """Row shards: leader's barrier nanoseconds since reset as (total, local wait, send, peer wait)."""
        a = (C.c_int64 * 8)()
        check(load().sigsdp_solver_debug_cycles(self.handle, a))
        return int(a[4]), int(a[5]), int(a[6]), int(a[7])

    def total_terms(self):
        v = C.c_int64()
        check(load().sigsdp_solver_total_terms(self.handle, C.byref(v)))
        return int(v.value)


class Batch:
    """Independent instances advanced together, one thread block each (sigsdp_batch_*)."""

    def __init__(self, solvers, ids=None):
        self.solvers = list(solvers)
        arr = (C.c_void_p * len(self.solvers))(*[s.handle for s in self.solvers])
        self.handle = C.c_void_p()
        if ids is None:
            check(load().sigsdp_batch_create(arr, len(self.solvers), C.byref(self.handle)))
        else:
            ida = np.ascontiguousarray(ids, dtype=np.int64)
            check(load().sigsdp_batch_create_ids(arr, _p(ida, C.c_int64), len(self.solvers), C.byref(self.handle)))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                load().sigsdp_batch_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def iterate(self, n_iters, seed=0, stream=None):
        check(load().sigsdp_batch_iterate(self.handle, int(n_iters), int(seed), stream))

    def blocks_per_instance(self):
        """Thread blocks each instance got in the last iterate (> 1 when the batch under-fills the GPU)."""
        return int(load().sigsdp_batch_blocks_per_instance(self.handle))


def debug_normals(seed, it, n, D, dtype=F64):
    out = np.empty((n, D))
    check(load().sigsdp_debug_normals(int(seed), int(it), n, D, dtype, _p(out, C.c_double)))
    return out
