"""CPU tests (no GPU): the C-ABI library loads and exports every symbol the header
declares; the native host logic (graph plan, sequential rounding pass) matches the
oracle; the sparse topology generator reproduces the reference's fixtures."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import mmw_oracle as orc
from sig_sdp_mmw_b200 import _lib
from sig_sdp_mmw_b200.topology import sparse_env
from tests.golden_util import CASES, ROOT, load_case


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "sigsdp_mmw.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(sigsdp_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 25
    lib = C.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert _lib.load().sigsdp_version() >= 100


def test_binding_covers_header():
    hdr = open(os.path.join(ROOT, "include", "sigsdp_mmw.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(sigsdp_[a-z0-9_]+)\s*\(", hdr))
    src = open(os.path.join(ROOT, "sig_sdp_mmw_b200", "_lib.py")).read()
    assert not [n for n in names if n not in src]


@pytest.mark.parametrize("name", CASES)
def test_plan_matches_oracle(name):
    g = load_case(name)
    p = orc.build_problem(g["Z"], g["state"])
    pl = _lib.Plan(g["state"], device=-1)
    gi, gj, tij, tji, ai, aj = pl.edges()
    assert (pl.n, pl.E_g, pl.E_a) == (p.K, p.E_g, p.E_a)
    assert pl.nnz == p.K + 2 * (p.E_g + p.E_a)
    np.testing.assert_array_equal(gi, p.gi)
    np.testing.assert_array_equal(gj, p.gj)
    np.testing.assert_array_equal(ai, p.ai)
    np.testing.assert_array_equal(aj, p.aj)
    np.testing.assert_array_equal(tij, p.tij)
    np.testing.assert_array_equal(tji, p.tji)
    S_sum, tn = pl.vectors()
    np.testing.assert_allclose(S_sum, p.S_sum, rtol=1e-14)
    np.testing.assert_allclose(tn, np.sqrt(np.asarray(p.T.multiply(p.T).sum(axis=1)).ravel()), rtol=1e-14)
    assert pl.nnzT == p.T.nnz
    # pattern is the symmetric union, diagonal included, columns ascending
    rp, col = pl.pattern()
    U = sp.csr_matrix((np.ones(pl.nnz), col, rp), shape=(pl.n, pl.n))
    assert (U != U.T).nnz == 0
    assert (U.diagonal() == 1).all()
    assert all(np.all(np.diff(col[rp[k]:rp[k + 1]]) > 0) for k in range(pl.n))


@pytest.mark.parametrize("order", [1, 16])
def test_plan_locality_order_is_a_relabelling(order):
    g = load_case("n300_z10")
    p0 = _lib.Plan(g["state"], device=-1, order=0)
    p1 = _lib.Plan(g["state"], device=-1, order=order)
    perm = p1.perm()
    assert sorted(perm.tolist()) == list(range(p0.n))
    for a, b in zip(p0.edges(), p1.edges()):      # edge lists stay in the caller's numbering
        np.testing.assert_array_equal(a, b)
    for a, b in zip(p0.vectors(), p1.vectors()):
        np.testing.assert_array_equal(a, b)
    rp0, c0 = p0.pattern()
    rp1, c1 = p1.pattern()
    U0 = sp.csr_matrix((np.ones(p0.nnz), c0, rp0), shape=(p0.n, p0.n))
    U1 = sp.csr_matrix((np.ones(p1.nnz), c1, rp1), shape=(p0.n, p0.n))
    P = sp.csr_matrix((np.ones(p0.n), (np.arange(p0.n), perm)), shape=(p0.n, p0.n))   # new <- old
    assert abs(P @ U0 @ P.T - U1).nnz == 0


def test_tile_stats_respect_the_caps():
    """Row tiles (HostTiles, plan_host.h): every row lands in exactly one tile, the staged rows
    cover at least the distinct columns and no tile exceeds the caps."""
    g = load_case("n1000_z8")
    pl = _lib.Plan(g["state"], device=-1, order=1)
    rp, col = pl.pattern()
    st = pl.tile_stats(32, 200, 1024)
    assert st["nnz"] == pl.nnz == rp[-1]
    assert st["tiles"] >= -(-pl.n // 32)
    assert st["umax"] <= 200 and st["nnzmax"] <= 1024
    assert 0 < st["runs"] <= st["staged_rows"]
    one_row = pl.tile_stats(1, 200, 1024)
    assert one_row["tiles"] == pl.n and one_row["nnzmax"] == int(np.diff(rp).max())
    with pytest.raises(RuntimeError):
        pl.tile_stats(32, 4, 1024)      # a single row has more neighbours than the cap


def _ring_state(n, hops=(1, 2, 5), seed=0):
    """Circulant interference graph with an association matching: cheap to build at any n."""
    rng = np.random.default_rng(seed)
    rows, cols = [], []
    for hp in hops:
        i = np.arange(n)
        rows += [i, (i + hp) % n]
        cols += [(i + hp) % n, i]
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    S = sp.csr_matrix((rng.uniform(0.1, 1.0, rows.size), (rows, cols)), shape=(n, n))
    S.sum_duplicates()
    a = np.arange(0, n - 1, 2)
    Q = sp.csr_matrix((np.ones(2 * a.size), (np.r_[a, a + 1], np.r_[a + 1, a])), shape=(n, n))
    S = S - S.multiply(Q)          # a pair is either a gain edge or an association edge
    S.eliminate_zeros()
    S.sort_indices()
    Q.sort_indices()
    return sp.csr_matrix(S), Q, rng.uniform(1.0, 2.0, n)


def test_plan_and_tiles_do_not_depend_on_host_thread_count(monkeypatch):
    """The plan stages and the tile builder run on all host cores (plan_host.cpp); their output
    must be the same whatever the thread count, including the 16 fixed row ranges large graphs
    are tiled in."""
    state = _ring_state(40000)
    out = []
    for threads in ("1", "3", None):
        if threads is None:
            monkeypatch.delenv("SIGSDP_HOST_THREADS", raising=False)
        else:
            monkeypatch.setenv("SIGSDP_HOST_THREADS", threads)
        pl = _lib.Plan(state, device=-1, order=1)
        out.append((pl.pattern(), pl.edges(), pl.vectors(), pl.perm(), pl.tile_stats(64, 300, 2048)))
    for other in out[1:]:
        for a, b in zip(out[0][0] + out[0][1] + out[0][2], other[0] + other[1] + other[2]):
            np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(out[0][3], other[3])
        assert out[0][4] == other[4]
    st = out[0][4]
    assert st["tiles"] >= 40000 // 64 and st["umax"] <= 300


def test_plan_rejects_bad_input():
    g = load_case("n75_z8")
    S, Q, h = g["state"]
    Qbad = Q.tolil()
    i, j = sp.triu(Q, 1).nonzero()
    Qbad[i[0], j[0]] = 0
    with pytest.raises(_lib.SigSdpError, match="symmetric"):
        _lib.Plan((S, Qbad.tocsr(), h), device=-1)
    Qd = Q.tolil()
    Qd[3, 3] = 1
    with pytest.raises(_lib.SigSdpError, match="diagonal"):
        _lib.Plan((S, Qd.tocsr(), h), device=-1)
    with pytest.raises(ValueError):
        _lib.Plan((S, Q, h[:-1]), device=-1)
    pl = _lib.Plan(g["state"], device=-1)
    with pytest.raises(_lib.SigSdpError, match="host-only"):
        _lib.Solver(pl, 4, 8, 0.1)


def _greedy(g, randv_raw):
    S, Q, h = g["state"]
    Z, gX = g["Z"], g["X_half"]
    K = S.shape[0]
    randv = randv_raw / np.linalg.norm(randv_raw, axis=1, keepdims=True)
    rank = np.argsort(-np.linalg.norm(gX, axis=1)).astype(np.int32)
    pref = np.ascontiguousarray(np.argsort(-(randv @ gX.T), axis=0).T.astype(np.int32))
    Sa, Qa = _lib.csr_arrays(S), _lib.csr_arrays(Q)
    z = np.empty(K, np.int32)
    rem = C.c_int64()
    _lib.check(_lib.load().sigsdp_round_greedy(
        K, Z, _lib._p(Sa[0], C.c_int32), _lib._p(Sa[1], C.c_int32), _lib._p(Sa[2], C.c_double),
        _lib._p(Qa[0], C.c_int32), _lib._p(Qa[1], C.c_int32), _lib._p(Qa[2], C.c_double),
        _lib._p(np.ascontiguousarray(h), C.c_double), _lib._p(rank, C.c_int32), _lib._p(pref, C.c_int32),
        _lib._p(z, C.c_int32), C.byref(rem)))
    return z, int(rem.value)


@pytest.mark.parametrize("name", CASES)
def test_native_greedy_pass_matches_reference_rounding(name):
    """sigsdp_round_greedy fed the same preference order reproduces the reference's
    rounding_one_attempt result stored in the fixture (np.random.seed(2000))."""
    g = load_case(name)
    rs = np.random.RandomState(2000)
    randv_raw = rs.randn(g["Z"], g["X_half"].shape[1])
    z, rem = _greedy(g, randv_raw)
    assert rem == int(g["round1_rem"])
    ok = z >= 0
    np.testing.assert_array_equal(z[ok].astype(float), g["round1_z"][ok])
    if rem:
        fill = rs.randint(g["Z"], size=rem)
        np.testing.assert_array_equal(fill.astype(float), g["round1_z"][~ok])


def test_native_greedy_pass_infeasible_Z():
    g = load_case("n500_z4_cfg1")     # Z=4 is below the association bound: leftovers expected
    z, rem = _greedy(g, np.random.RandomState(1).randn(g["Z"], g["X_half"].shape[1]))
    assert rem > 0 and (z < 0).sum() == rem


@pytest.mark.parametrize("name,kw", [
    ("n75_z8", dict(cell_size=5, sta_density_per_1m2=75e-4, seed=0)),
    ("n75_z6_rr3", dict(cell_size=5, sta_density_per_1m2=75e-4, seed=3)),
    ("n300_z10", dict(cell_size=10, sta_density_per_1m2=75e-4, seed=1)),
    ("n500_z13", dict(cell_size=10, sta_density_per_1m2=125e-4, seed=0)),
])
def test_sparse_topology_reproduces_reference_state(name, kw):
    g = load_case(name)
    S, Q, h = g["state"]
    S2, Q2, h2 = sparse_env(**kw).generate_S_Q_hmax()
    assert S2.shape == S.shape and S2.nnz == S.nnz
    np.testing.assert_array_equal(S2.indptr, S.indptr)
    np.testing.assert_array_equal(S2.indices, S.indices)
    np.testing.assert_allclose(S2.data, S.data, rtol=1e-12)
    assert abs(Q2 - Q).nnz == 0
    np.testing.assert_allclose(h2, h, rtol=1e-12)


def test_sparse_topology_large_is_bounded_degree():
    e = sparse_env(cell_size=60, sta_density_per_1m2=6.25e-3, seed=0)
    S, Q, h = e.generate_S_Q_hmax()
    assert S.shape[0] == 9000
    deg = np.diff((S + S.T).tocsr().indptr)
    assert deg.max() < 200 and 10 < deg.mean() < 60
    assert (h > 0).all()


@pytest.mark.parametrize("name", ["a", "b"])
def test_sparse_evaluate_sinr_bler_matches_reference(name):
    """env.evaluate_sinr / evaluate_bler twins against fixtures from the unmodified reference."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "evaluate_n75_n300.npz"))
    cs, rho, seed = g[name + "_kw"]
    e = sparse_env(cell_size=int(cs), sta_density_per_1m2=float(rho), seed=int(seed))
    z, Z = g[name + "_z"], int(g[name + "_Z"])
    np.testing.assert_allclose(e.evaluate_sinr(z, Z, exact=True), g[name + "_sinr"], rtol=1e-12)
    np.testing.assert_allclose(e.evaluate_bler(z, Z, exact=True), g[name + "_bler"], rtol=1e-9, atol=1e-300)
    # the truncated (k-d tree) evaluation drops far-field power below floor_ratio x the link
    # threshold (a few % of the interference at the default 3e-2, < 1e-3 at 1e-3); which of
    # several equal-SINR stations of an AP wins a slot may differ
    approx = e.evaluate_sinr(z, Z, exact=False)
    np.testing.assert_allclose(np.sort(approx), np.sort(g[name + "_sinr"]), rtol=5e-2)
    fine = e.evaluate_sinr(z, Z, exact=False, floor_ratio=1e-3)
    np.testing.assert_allclose(np.sort(fine), np.sort(g[name + "_sinr"]), rtol=2e-3)


def test_plan_image_round_trip_and_rejection():
    """sigsdp_plan_image / sigsdp_plan_create_from_image: a plan rebuilt from the image of another is the same plan;
    an image of a different state or a damaged one is refused."""
    from sig_sdp_mmw_b200.topology import sparse_env
    state = sparse_env(cell_size=12, sta_density_per_1m2=75e-4, seed=3).generate_S_Q_hmax()
    a = _lib.Plan(state, device=-1, order=1)
    img = a.image()
    b = _lib.Plan(state, device=-1, order=1, image=img)
    assert (a.n, a.E_g, a.E_a, a.nnz, a.nnzT, a.order, a.max_row) == (b.n, b.E_g, b.E_a, b.nnz, b.nnzT, b.order, b.max_row)
    for x, y in zip(a.pattern() + a.edges() + a.vectors() + (a.perm(),), b.pattern() + b.edges() + b.vectors() + (b.perm(),)):
        np.testing.assert_array_equal(x, y)
    assert a.tile_stats(32, 343, 2048) == b.tile_stats(32, 343, 2048)
    other = sparse_env(cell_size=11, sta_density_per_1m2=75e-4, seed=3).generate_S_Q_hmax()
    with pytest.raises(_lib.SigSdpError):
        _lib.Plan(other, device=-1, order=1, image=img)          # image of another graph
    with pytest.raises(_lib.SigSdpError):
        _lib.Plan(state, device=-1, order=1, image=img[:-8])     # truncated
    bad = img.copy()
    bad[:8] = 0
    with pytest.raises(_lib.SigSdpError):
        _lib.Plan(state, device=-1, order=1, image=bad)          # no magic


def test_checksum_sees_values_positions_and_length_whatever_the_thread_count(monkeypatch):
    """The plan cache's key (sdp_solver._plan_for): equal buffers give equal checksums on any number of host
    threads; a changed value, two swapped values, a truncation or an appended zero give another."""
    rs = np.random.RandomState(0)
    a = rs.rand(300_001)
    ref = _lib.checksum(a)
    for nt in ("1", "3", "16"):
        monkeypatch.setenv("SIGSDP_HOST_THREADS", nt)
        assert _lib.checksum(a.copy()) == ref
    monkeypatch.delenv("SIGSDP_HOST_THREADS")
    b = a.copy(); b[123456] = np.nextafter(b[123456], 2.0)
    c = a.copy(); c[[10, 200_000]] = c[[200_000, 10]]
    seen = {ref, _lib.checksum(b), _lib.checksum(c), _lib.checksum(a[:-1]), _lib.checksum(np.append(a, 0.0))}
    assert len(seen) == 5
    idx = np.arange(11, dtype=np.int32)                      # 44 bytes: five words and a 4-byte tail
    swapped = idx.copy(); swapped[[9, 10]] = swapped[[10, 9]]
    assert _lib.checksum(idx) != _lib.checksum(swapped)
    assert _lib.checksum(np.zeros(0)) == _lib.checksum(np.zeros(0, np.int32))
    # and the cache key built from it tells apart two states that differ in one stored value
    import importlib
    sdp_solver = importlib.import_module("sig_sdp_mmw_b200.sdp_solver")
    S = sp.random(50, 50, density=0.2, random_state=rs, format="csr")
    S2 = S.copy(); S2.data[7] *= 2.0
    assert sdp_solver._digest(S) != sdp_solver._digest(S2)
    assert sdp_solver._digest(S) == sdp_solver._digest(S.copy())


def test_plan_is_destroyed_with_its_python_object():
    """A plan owns a device slab and host arrays of the graph's size: dropping the Python object must free them
    (a Monte-Carlo sweep builds one plan per drop; a leak there also fragments the device pool and slows every
    later set-up)."""
    g = load_case(CASES[0])
    plan = _lib.Plan(g["state"], device=-1, order=1)
    assert plan.handle
    plan.__del__()
    assert plan.handle is None
    plan.__del__()                                   # idempotent
    sol_cls = _lib.Solver
    assert "plan" in sol_cls.__init__.__code__.co_names or "plan" in sol_cls.__init__.__code__.co_varnames   # a solver keeps its plan alive


@pytest.mark.parametrize("seed,warm,shape", [
    (0, 0, (7,)), (1, 1, (1000, 3)), (3, 0, (1,)), (4, 5, (2,)), (5, 2, (12345, 7)),
    (6, 0, (3200001,)),              # several chunks, odd count: the last pair's second number stays cached
    (7, 0, (411775,)), (8, 1, (823552,)),   # counts around one chunk's worth of accepted pairs
    (9, 3, (3, 50000, 8)),
])
def test_native_numpy_normal_stream_is_numpys_bit_for_bit(seed, warm, shape, monkeypatch):
    """sigsdp_numpy_standard_normal continues np.random's GLOBAL legacy stream (MT19937 -> doubles -> polar method,
    mmw.py:226 draws np.random.randn(K, D) per iteration): same numbers as numpy, and numpy's state afterwards is what
    numpy's own draw would have left (next normals AND next uniforms agree), whatever the host thread count."""
    def numpys():
        np.random.seed(seed)
        if warm:
            np.random.randn(warm)        # an odd warm-up leaves a cached second normal behind
        return np.random.standard_normal(shape), np.random.randn(5), np.random.rand(3)

    def ours():
        np.random.seed(seed)
        if warm:
            np.random.randn(warm)
        out = np.empty(shape)
        _lib.numpy_randn_into(out)
        return out, np.random.randn(5), np.random.rand(3)
    ref = numpys()
    for nt in (None, "1", "3"):
        if nt:
            monkeypatch.setenv("SIGSDP_HOST_THREADS", nt)
        got = ours()
        for a, b in zip(ref, got):
            assert np.array_equal(a, b)


def test_native_numpy_normal_stream_rejects_bad_buffers_and_other_generators():
    with pytest.raises(ValueError):
        _lib.numpy_randn_into(np.empty(4, np.float32))
    with pytest.raises(ValueError):
        _lib.numpy_randn_into(np.empty((4, 4))[:, ::2])
    out = np.empty(0)
    assert _lib.numpy_randn_into(out) is out                                  # nothing drawn, state untouched
    lib = _lib.load()
    key = np.zeros(624, np.uint32)
    pos, hg, g = C.c_int32(700), C.c_int32(0), C.c_double(0.0)                # pos outside the twister block
    rc = lib.sigsdp_numpy_standard_normal(key.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(pos), C.byref(hg), C.byref(g), 4,
                                          np.empty(4).ctypes.data_as(C.POINTER(C.c_double)))
    assert rc != 0


@pytest.mark.timeout(120)
def test_worker_pool_survives_fork_and_concurrent_callers():
    """The parallel stages run on a pool of sleeping workers: a forked child (which has none of the parent's threads)
    must start its own instead of waiting for workers that do not exist, and host threads calling at the same time
    must all get the right answer (one uses the pool, the others plain threads)."""
    import threading
    a = np.random.RandomState(0).rand(2_000_000)
    ref = _lib.checksum(a)                                   # the pool exists now
    pid = os.fork()
    if pid == 0:
        ok = _lib.checksum(a) == ref
        b = np.empty(300000)
        np.random.seed(1)
        _lib.numpy_randn_into(b)
        np.random.seed(1)
        ok = ok and np.array_equal(b, np.random.standard_normal(300000))
        os._exit(0 if ok else 3)
    _, status = os.waitpid(pid, 0)
    assert os.WIFEXITED(status) and os.WEXITSTATUS(status) == 0
    assert _lib.checksum(a) == ref
    res = []

    def work():
        for _ in range(40):
            res.append(_lib.checksum(a) == ref)
    th = [threading.Thread(target=work) for _ in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert len(res) == 160 and all(res)
