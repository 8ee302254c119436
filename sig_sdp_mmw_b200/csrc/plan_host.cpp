// Host-side graph plan builder and the sequential rounding pass (native C++).
// Follows mmw.py:26-41 (_process_state), mmw.py:52-57 (edge lists) and
// sdp_solver.py:70-101 (greedy feasibility pass) of the reference; written against
// flat CSR arrays instead of scipy objects.
#include "plan_host.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <pthread.h>
#include <cmath>
#include <cstdio>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <numeric>
#include <queue>
#include <thread>

#include "../../include/sigsdp_mmw.h"

namespace sigsdp {
namespace {
struct StageTimer {
    bool on = getenv("SIGSDP_PLAN_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void lap(const char* what) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[plan] %-24s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};


// first problem of rows [r0, r1) of a CSR structure, or nullptr
const char* check_csr_rows(int64_t n, const int32_t* p, const int32_t* idx, int64_t r0, int64_t r1) {
    for (int64_t r = r0; r < r1; ++r) {
        if (p[r + 1] < p[r]) return "indptr not monotone";
        for (int32_t q = p[r]; q < p[r + 1]; ++q) {
            if (idx[q] < 0 || idx[q] >= n) return "column index out of range";
            if (q > p[r] && idx[q] <= idx[q - 1]) return "indices must be sorted and duplicate-free per row";
        }
    }
    return nullptr;
}

// value of the CSR entry (r, c) or 0
inline double csr_at(const int32_t* p, const int32_t* idx, const double* x, int32_t r, int32_t c) {
    const int32_t* b = idx + p[r];
    const int32_t* e = idx + p[r + 1];
    const int32_t* it = std::lower_bound(b, e, c);
    if (it == e || *it != c) return 0.0;
    return x[it - idx];
}

struct Ent {
    int32_t col;
    int32_t kind;  // 0 diag, 1 gain, 2 asso
    double tf, tb;
};

// fn(begin, end) over [0, n) on the host's cores (contiguous chunks; fn must only touch
// its own rows' outputs)
// host threads for the builders: all cores, shared fairly when several ranks of one launch
// (torchrun's LOCAL_WORLD_SIZE) build their plans on the same box at the same time
unsigned host_threads() {
    unsigned nt = std::thread::hardware_concurrency();
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) nt /= (unsigned)std::max(1, atoi(e));
    if (const char* e = getenv("SIGSDP_HOST_THREADS")) nt = (unsigned)std::max(1, atoi(e));
    return std::max(1u, std::min(nt, 32u));
}

std::atomic<int> g_side_threads{0};   // sequential helper threads running next to the parallel stages

// Worker pool behind the parallel stages.  A set-up runs ~20 short parallel stages (checksums, checks, copies, merges) and
// the numpy stream two per chunk; starting and joining 16 threads for each costs 0.15-0.3 ms a time, a fifth of the set-up
// at 100k nodes.  The workers are started once and sleep on a condition variable between jobs.  One job at a time: a second
// caller (another host thread, or a task that itself calls a parallel stage) falls back to plain threads.  The pool lives
// on the heap and is never destroyed (workers asleep at exit die with the process); a forked child starts a fresh one.
class WorkerPool {
public:
    static WorkerPool* get() {
        WorkerPool* p = g_pool.load(std::memory_order_acquire);
        if (p) return p;
        static std::mutex create_mu;
        std::lock_guard<std::mutex> lock(create_mu);
        p = g_pool.load(std::memory_order_acquire);
        if (!p) {
            static bool hooked = false;
            if (!hooked) {
                pthread_atfork(nullptr, nullptr, [] { g_pool.store(nullptr); });   // the child has none of the parent's threads
                hooked = true;
            }
            p = new WorkerPool();
            g_pool.store(p, std::memory_order_release);
        }
        return p;
    }
    // task(i) for i in [0, ntasks) on `width` threads, the caller being one of them; false (nothing run) when the pool is busy
    bool try_run(unsigned width, int64_t ntasks, const std::function<void(int64_t)>& task) {
        if (t_inside || !run_mu.try_lock()) return false;
        std::lock_guard<std::mutex> run_lock(run_mu, std::adopt_lock);
        const unsigned helpers = (unsigned)std::min<int64_t>(width > 0 ? width - 1 : 0, std::max<int64_t>(ntasks - 1, 0));
        {
            std::unique_lock<std::mutex> lk(m);
            while (workers.size() < helpers) {
                const unsigned id = (unsigned)workers.size();
                workers.emplace_back([this, id] { worker_loop(id); });
            }
            job = &task;
            job_tasks = ntasks;
            job_helpers = helpers;
            next.store(0);
            pending = helpers;
            ++epoch;
        }
        cv_work.notify_all();
        t_inside = true;
        for (int64_t i = next.fetch_add(1); i < ntasks; i = next.fetch_add(1)) task(i);
        t_inside = false;
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&] { return pending == 0; });
        job = nullptr;
        return true;
    }

private:
    void worker_loop(unsigned id) {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            cv_work.wait(lk, [&] { return epoch != seen; });
            seen = epoch;
            if (id >= job_helpers) continue;   // not needed for this job
            const std::function<void(int64_t)>* f = job;
            const int64_t n = job_tasks;
            lk.unlock();
            t_inside = true;
            for (int64_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) (*f)(i);
            t_inside = false;
            lk.lock();
            if (--pending == 0) cv_done.notify_one();
        }
    }
    static std::atomic<WorkerPool*> g_pool;
    static thread_local bool t_inside;
    std::mutex run_mu, m;
    std::condition_variable cv_work, cv_done;
    std::vector<std::thread> workers;
    const std::function<void(int64_t)>* job = nullptr;
    int64_t job_tasks = 0;
    unsigned job_helpers = 0, pending = 0;
    uint64_t epoch = 0;
    std::atomic<int64_t> next{0};
};
std::atomic<WorkerPool*> WorkerPool::g_pool{nullptr};
thread_local bool WorkerPool::t_inside = false;

// task(i), i in [0, ntasks), on up to `width` threads
void run_tasks(unsigned width, int64_t ntasks, const std::function<void(int64_t)>& task) {
    if (ntasks <= 0) return;
    if (width <= 1 || ntasks == 1) {
        for (int64_t i = 0; i < ntasks; ++i) task(i);
        return;
    }
    if (WorkerPool::get()->try_run(width, ntasks, task)) return;
    std::atomic<int64_t> next{0};
    auto work = [&] { for (int64_t i = next.fetch_add(1); i < ntasks; i = next.fetch_add(1)) task(i); };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < width && (int64_t)t < ntasks; ++t) th.emplace_back(work);
    work();
    for (auto& x : th) x.join();
}

template <class F>
void parallel_rows(int64_t n, F fn, int64_t min_parallel = 4096) {
    const unsigned nt = (unsigned)std::max(1, (int)host_threads() - g_side_threads.load());
    if (n < min_parallel || nt == 1) {
        fn(0, n);
        return;
    }
    const int64_t chunk = (n + nt - 1) / nt;
    const int64_t nchunks = (n + chunk - 1) / chunk;
    run_tasks(nt, nchunks, [&](int64_t t) { fn(t * chunk, std::min<int64_t>(n, t * chunk + chunk)); });
}

// Clustered BFS ordering: grow clusters of ~cluster nodes breadth-first, visiting the
// graph cluster by cluster so that the rows of one CTA tile are a compact patch of
// the (geometric) interference graph and share their neighbours.  It walks the INPUTS (row u of
// S_gain and of Q_asso: every pair of the union pattern is in the row of at least one of its two
// ends), so it can start right after validation and run on its own thread while the other
// cores build T and the union pattern; the tiles it gives are as compact as those of a walk over
// the symmetric pattern (cfg4: 414.8k staged rows / 39.1k copy runs against 421.6k / 41.3k).
void locality_order_inputs(int64_t n, const int32_t* Sp, const int32_t* Si, const int32_t* Qp, const int32_t* Qi,
                           int cluster, std::vector<int32_t>& perm) {
    perm.clear();
    perm.reserve(n);
    std::vector<uint8_t> seen(n, 0);
    std::vector<int32_t> frontier;  // nodes adjacent to finished clusters, FIFO
    frontier.reserve(n);
    size_t fhead = 0;
    std::vector<int32_t> q;
    q.reserve(cluster + 8);
    // seen: 0 new, 1 placed in a cluster, 2 waiting in the frontier (queued once: a node's later occurrences in the
    // FIFO would be skipped anyway, because the first one is popped first and places it)
    auto visit = [&](int32_t v) {
        if ((uint32_t)v >= (uint32_t)n) return;   // (the walk starts before the index checks of validate_state finish)
        const uint8_t sv = seen[v];
        if (sv == 1) return;
        if ((int)q.size() < cluster) {
            seen[v] = 1;
            q.push_back(v);
            __builtin_prefetch(Sp + v);   // its row pointers now, its rows when it is two pops away:
            __builtin_prefetch(Qp + v);   // the walk is bound by the latency of these scattered reads
        } else if (sv == 0) {
            seen[v] = 2;
            frontier.push_back(v);
        }
    };
    for (int64_t start = 0; start < n; ++start) {
        if (seen[start]) continue;
        frontier.push_back((int32_t)start);
        while (fhead < frontier.size()) {
            const int32_t seed = frontier[fhead++];
            if (seen[seed] == 1) continue;
            q.clear();
            q.push_back(seed);
            seen[seed] = 1;
            size_t qh = 0;
            while (qh < q.size()) {
                const int32_t u = q[qh++];
                if (qh + 1 < q.size()) {
                    const int32_t u2 = q[qh + 1];
                    __builtin_prefetch(Si + Sp[u2]);
                    __builtin_prefetch(Si + Sp[u2] + 16);
                    __builtin_prefetch(Qi + Qp[u2]);
                }
                // (stored zeros are walked like any entry: this is an ordering heuristic, and not
                // touching the value arrays halves the memory traffic of the one sequential stage)
                for (int32_t e = Sp[u]; e < Sp[u + 1]; ++e) visit(Si[e]);
                for (int32_t e = Qp[u]; e < Qp[u + 1]; ++e) visit(Qi[e]);
            }
            for (int32_t u : q) perm.push_back(u);
        }
    }
}

}  // namespace


namespace {
constexpr int64_t IMAGE_MAGIC = 0x5347504c414e3031ll;   // "SGPLAN01"
struct ImageField {
    void* ptr;
    size_t bytes;
};
// the arrays of a plan in image order; sizes from the header fields already set in P
template <class PlanT, class F>
void image_fields(PlanT& P, F&& f) {
    f(P.rowptr, (size_t)P.n + 1); f(P.col, (size_t)P.nnz); f(P.eid, (size_t)P.nnz);
    f(P.tfwd, (size_t)P.nnz); f(P.tbwd, (size_t)P.nnz);
    f(P.gi, (size_t)P.E_g); f(P.gj, (size_t)P.E_g); f(P.ai, (size_t)P.E_a); f(P.aj, (size_t)P.E_a);
    f(P.tij, (size_t)P.E_g); f(P.tji, (size_t)P.E_g);
    f(P.S_sum, (size_t)P.n); f(P.tnorm, (size_t)P.n); f(P.h_max, (size_t)P.n);
    f(P.perm, (size_t)P.n); f(P.iperm, (size_t)P.n); f(P.dpos, (size_t)P.n); f(P.apos, (size_t)P.E_a);
}
inline size_t pad8(size_t b) { return (b + 7) & ~(size_t)7; }
}  // namespace

// position-weighted checksum of a buffer (64-bit words, mixed, odd weights; wraps): integer arithmetic, so the value
// does not depend on how the words are dealt to the threads
uint64_t checksum_bytes(const void* data, size_t bytes) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    const int64_t nw = (int64_t)(bytes / 8);
    std::atomic<uint64_t> total{0};
    auto mix = [](uint64_t w, uint64_t i) {
        w *= 0xff51afd7ed558ccdull;
        w ^= w >> 32;
        return w * (2 * i + 1);
    };
    parallel_for(nw, [&](int64_t b, int64_t e) {
        uint64_t s = 0;
        for (int64_t i = b; i < e; ++i) {
            uint64_t w;
            std::memcpy(&w, p + 8 * i, 8);
            s += mix(w, (uint64_t)i);
        }
        total.fetch_add(s);
    }, 1 << 17);
    uint64_t tail = 0;
    std::memcpy(&tail, p + 8 * nw, bytes - 8 * (size_t)nw);
    return total.load() + mix(tail ^ (uint64_t)bytes, (uint64_t)nw);
}

size_t host_plan_image_bytes(const HostPlan& P) {
    size_t tot = 8 * sizeof(int64_t);
    image_fields(P, [&](const auto& v, size_t cnt) { tot += pad8(cnt * sizeof(v[0])); });
    return tot;
}

void host_plan_to_image(const HostPlan& P, unsigned char* buf) {
    int64_t hdr[8] = {IMAGE_MAGIC, P.n, P.E_g, P.E_a, P.nnz, P.nnzT, P.max_row, P.order};
    std::memcpy(buf, hdr, sizeof(hdr));
    size_t off = sizeof(hdr);
    std::vector<CopySeg> segs;
    image_fields(P, [&](const auto& v, size_t cnt) {
        const size_t b = cnt * sizeof(v[0]);
        segs.push_back(CopySeg{buf + off, v.data(), b});
        std::memset(buf + off + b, 0, pad8(b) - b);   // (equal plans give equal images)
        off += pad8(b);
    });
    parallel_copy(segs);
}

bool host_plan_from_image(const unsigned char* buf, size_t bytes, int64_t n, HostPlan& P, std::string& err) {
    int64_t hdr[8];
    if (bytes < sizeof(hdr)) { err = "plan image too short"; return false; }
    std::memcpy(hdr, buf, sizeof(hdr));
    if (hdr[0] != IMAGE_MAGIC || hdr[1] != n || hdr[2] < 0 || hdr[3] < 0 || hdr[4] != n + 2 * (hdr[2] + hdr[3])) {
        err = "plan image does not describe a plan of this state";
        return false;
    }
    P = HostPlan();
    P.n = hdr[1]; P.E_g = hdr[2]; P.E_a = hdr[3]; P.nnz = hdr[4]; P.nnzT = hdr[5]; P.max_row = (int)hdr[6]; P.order = (int)hdr[7];
    if (host_plan_image_bytes(P) != bytes) { err = "plan image has the wrong size"; return false; }
    size_t off = sizeof(hdr);
    std::vector<CopySeg> segs;
    image_fields(P, [&](auto& v, size_t cnt) {
        v.resize(cnt);
        segs.push_back(CopySeg{v.data(), buf + off, cnt * sizeof(v[0])});
        off += pad8(cnt * sizeof(v[0]));
    });
    parallel_copy(segs);
    if (P.rowptr[0] != 0 || P.rowptr[n] != P.nnz) { err = "plan image is inconsistent"; return false; }
    return true;
}

void shard_cut_points(const HostPlan& P, const HostTiles* ht, int nranks, std::vector<int32_t>& row0,
                      std::vector<int32_t>& tile0) {
    const int nb = ht ? ht->ntiles : (int)P.n;
    auto brow = [&](int b) { return ht ? ht->trow[b] : b; };
    row0.assign(nranks + 1, 0);
    tile0.assign(nranks + 1, 0);
    for (int r = 1; r < nranks; ++r) {
        const double target = (double)P.nnz * r / nranks;
        int lo = tile0[r - 1], hi = nb;
        while (lo < hi) {   // first boundary whose cumulative non-zeros reach the target
            const int mid = (lo + hi) / 2;
            if ((double)P.rowptr[brow(mid)] < target) lo = mid + 1; else hi = mid;
        }
        tile0[r] = std::max(lo, tile0[r - 1]);
        row0[r] = brow(tile0[r]);
    }
    tile0[nranks] = nb;
    row0[nranks] = (int32_t)P.n;
}

void shard_halo(const HostPlan& P, const std::vector<int32_t>& row0, int rank, ShardHalo& out) {
    const int64_t n = P.n;
    const int32_t lo = row0[rank], hi = row0[rank + 1];
    auto owner = [&](int32_t row) { return (int)(std::upper_bound(row0.begin() + 1, row0.end(), row) - (row0.begin() + 1)); };
    out = ShardHalo();
    out.pmask.assign(n, 0);
    std::vector<uint8_t> seen(n, 0);
    std::vector<int32_t> for_e, for_p;
    for (int32_t k = lo; k < hi; ++k) {
        unsigned m = 0;
        for (int32_t p = P.rowptr[k]; p < P.rowptr[k + 1]; ++p) {
            const int32_t c = P.col[p];
            const bool own_c = c >= lo && c < hi;
            if (!own_c) {
                m |= 1u << owner(c);
                if (!seen[c]) {
                    seen[c] = 1;
                    ++out.recv;
                }
            }
            if (P.eid[p] >= P.E_g) {
                if (k < c) {
                    out.inc_e.push_back(P.eid[p] - (int32_t)P.E_g);
                    out.inc_p.push_back(p);
                } else if (!own_c) {
                    for_e.push_back(P.eid[p] - (int32_t)P.E_g);
                    for_p.push_back(p);
                }
            }
        }
        out.pmask[k] = (uint8_t)m;
        out.send += __builtin_popcount(m);
    }
    out.n_inc_owned = (int)out.inc_e.size();
    out.inc_e.insert(out.inc_e.end(), for_e.begin(), for_e.end());
    out.inc_p.insert(out.inc_p.end(), for_p.begin(), for_p.end());
}

void side_thread_begin() { g_side_threads.fetch_add(1); }
void side_thread_end() { g_side_threads.fetch_sub(1); }

void parallel_for(int64_t n, const std::function<void(int64_t, int64_t)>& fn, int64_t min_parallel) {
    parallel_rows(n, fn, min_parallel);
}

void parallel_copy(const std::vector<CopySeg>& segs) {
    constexpr size_t CH = (size_t)1 << 20;
    std::vector<CopySeg> tasks;
    for (const CopySeg& s : segs)
        for (size_t o = 0; o < s.bytes; o += CH)
            tasks.push_back(CopySeg{static_cast<char*>(s.dst) + o, static_cast<const char*>(s.src) + o, std::min(CH, s.bytes - o)});
    const unsigned nt = std::min<unsigned>(host_threads(), (unsigned)std::max<size_t>(1, tasks.size()));
    run_tasks(nt, (int64_t)tasks.size(), [&](int64_t i) { std::memcpy(tasks[i].dst, tasks[i].src, tasks[i].bytes); });
}

int validate_pointers(int64_t n, const int32_t* Sp, const int32_t* Si, const double* Sx, const int32_t* Qp, const int32_t* Qi,
                      const double* Qx, const double* h_max, std::string& err) {
    if (n <= 1 || n > (int64_t)1 << 30) {
        err = "n must be in [2, 2^30]";
        return SIGSDP_EINVAL;
    }
    if (!Sp || !Si || !Sx || !Qp || !Qi || !Qx || !h_max) {
        err = "null input array";
        return SIGSDP_EINVAL;
    }
    if (Sp[0] != 0 || Qp[0] != 0) {
        err = std::string(Sp[0] != 0 ? "S_gain" : "Q_asso") + ": indptr[0] != 0";
        return SIGSDP_EINVAL;
    }
    for (int64_t r = 0; r < n; ++r)
        if (Sp[r + 1] < Sp[r] || Qp[r + 1] < Qp[r]) {
            err = std::string(Sp[r + 1] < Sp[r] ? "S_gain" : "Q_asso") + ": indptr not monotone";
            return SIGSDP_EINVAL;
        }
    return SIGSDP_OK;
}

int validate_state(int64_t n, const int32_t* Sp, const int32_t* Si, const double* Sx, const int32_t* Qp, const int32_t* Qi,
                   const double* Qx, const double* h_max, std::string& err) {
    if (n <= 1 || n > (int64_t)1 << 30) {
        err = "n must be in [2, 2^30]";
        return SIGSDP_EINVAL;
    }
    if (!Sp || !Si || !Sx || !Qp || !Qi || !Qx || !h_max) {
        err = "null input array";
        return SIGSDP_EINVAL;
    }
    if (Sp[0] != 0 || Qp[0] != 0) {
        err = std::string(Sp[0] != 0 ? "S_gain" : "Q_asso") + ": indptr[0] != 0";
        return SIGSDP_EINVAL;
    }
    // Structure checks on all host cores; every chunk keeps its first problem and the one in
    // the lowest chunk is reported (the same message a sequential scan would give, up to which
    // of two broken matrices is named first: S_gain wins).  Q_asso must be symmetric with an
    // empty diagonal (the reference counts E_asso = nnz(Q)/2 and indexes triu(Q,1), mmw.py:57-59).
    {
        struct Issue { int64_t row = -1; std::string msg; };
        std::vector<Issue> issues(64);
        // the monotone check must hold everywhere before any row content is dereferenced
        for (int64_t r = 0; r < n; ++r)
            if (Sp[r + 1] < Sp[r] || Qp[r + 1] < Qp[r]) {
                err = std::string(Sp[r + 1] < Sp[r] ? "S_gain" : "Q_asso") + ": indptr not monotone";
                return SIGSDP_EINVAL;
            }
        std::atomic<int> slot{0};
        auto note = [&](int64_t row, std::string msg) {
            const int k = slot.fetch_add(1);
            if (k < (int)issues.size()) {
                issues[k].row = row;
                issues[k].msg = std::move(msg);
            }
        };
        for (int pass = 0; pass < 3; ++pass) {
            parallel_rows(n, [&](int64_t r0, int64_t r1) {
                if (pass == 0) {
                    if (const char* m = check_csr_rows(n, Sp, Si, r0, r1)) note(r0, std::string("S_gain: ") + m);
                } else if (pass == 1) {
                    if (const char* m = check_csr_rows(n, Qp, Qi, r0, r1)) note(r0, std::string("Q_asso: ") + m);
                } else {
                    for (int64_t i = r0; i < r1; ++i)
                        for (int32_t q = Qp[i]; q < Qp[i + 1]; ++q) {
                            if (Qx[q] == 0.0) continue;
                            const int32_t j = Qi[q];
                            if (j == i) {
                                note(r0, "Q_asso has a non-zero diagonal entry");
                                return;
                            }
                            if (csr_at(Qp, Qi, Qx, j, (int32_t)i) == 0.0) {
                                note(r0, "Q_asso is not structurally symmetric");
                                return;
                            }
                        }
                }
            });
            const int cnt = std::min<int>(slot.load(), (int)issues.size());
            if (cnt > 0) {
                int best = 0;
                for (int k = 1; k < cnt; ++k)
                    if (issues[k].row < issues[best].row) best = k;
                err = issues[best].msg;
                return SIGSDP_EINVAL;
            }
        }
    }
    return SIGSDP_OK;
}

void locality_order_of_inputs(int64_t n, const int32_t* Sp, const int32_t* Si, const int32_t* Qp, const int32_t* Qi, int cluster,
                              std::vector<int32_t>& perm) {
    g_side_threads.fetch_add(1);   // (runs on a thread of its own next to the callers' parallel stages)
    locality_order_inputs(n, Sp, Si, Qp, Qi, cluster, perm);
    g_side_threads.fetch_sub(1);
}

int build_host_plan(int64_t n, const int32_t* Sp, const int32_t* Si, const double* Sx,
                    const int32_t* Qp, const int32_t* Qi, const double* Qx, const double* h_max,
                    int order, HostPlan& P, std::string& err) {
    StageTimer tm;
    {
        const int rc = validate_pointers(n, Sp, Si, Sx, Qp, Qi, Qx, h_max, err);
        if (rc != SIGSDP_OK) return rc;
    }
    // The locality ordering is the one sequential stage (~11-14 ms at 100k nodes): it only needs the
    // inputs, so it runs on its own thread (the parallel stages leave it a core) while the other
    // cores check the inputs and build T, the union pattern and the edge ids; it is joined in front
    // of the renumbering.  (It guards its own index reads: the index checks run next to it.)
    std::vector<int32_t> bfs_perm;
    std::thread bfs_thread;
    if (order != 0) {
        g_side_threads.fetch_add(1);
        bfs_thread = std::thread([&, n] {
            StageTimer bt;
            locality_order_inputs(n, Sp, Si, Qp, Qi, order > 1 ? order : 64, bfs_perm);
            g_side_threads.fetch_sub(1);
            bt.lap("  (clustered BFS thread)");
        });
    }
    struct Joiner {   // every early return below must not leave the thread running on dead locals
        std::thread& t;
        ~Joiner() { if (t.joinable()) t.join(); }
    } bfs_joiner{bfs_thread};
    {
        const int rc = validate_state(n, Sp, Si, Sx, Qp, Qi, Qx, h_max, err);
        if (rc != SIGSDP_OK) return rc;
    }
    P = HostPlan();
    P.n = n;
    P.order = order;
    tm.lap("validate");
    // ---- T = S^T with association pairs and the diagonal zeroed (mmw.py:28-33)
    std::vector<int32_t> Tp(n + 1, 0);
    hvec<int32_t> Ti;
    hvec<double> Tx;
    std::vector<uint8_t> keep(Sp[n], 0);   // S entries that survive into T (= T^T entries of their own row)
    {
        // S entry (j, i) becomes T[i][j] unless i == j, the value is zero, or Q[i][j] != 0.
        // Q is symmetric (checked above), so Q[i][j] != 0 <=> Q[j][i] != 0: merge S row j with
        // Q row j (both sorted) instead of searching Q row i.
        // Transpose = counting sort with one histogram per thread: thread t owns the S rows
        // [j0_t, j1_t); its entries of T row i go behind those of the threads before it, so every
        // T row comes out sorted by column whatever the thread count.
        const unsigned nt = (unsigned)std::max(1, (int)host_threads() - g_side_threads.load());
        const unsigned nth = n < 4096 ? 1u : nt;
        const int64_t chunk = (n + nth - 1) / nth;
        std::vector<std::vector<int32_t>> cnt(nth);
        auto run_threads = [&](auto fn) {
            if (nth == 1) { fn(0u); return; }
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nth; ++t) th.emplace_back([=] { fn(t); });
            for (auto& x : th) x.join();
        };
        run_threads([&](unsigned t) {
            std::vector<int32_t>& c = cnt[t];
            c.assign(n, 0);
            const int64_t j0 = t * chunk, j1 = std::min<int64_t>(n, j0 + chunk);
            for (int64_t j = j0; j < j1; ++j) {
                int32_t qq = Qp[j];
                const int32_t qe = Qp[j + 1];
                for (int32_t q = Sp[j]; q < Sp[j + 1]; ++q) {
                    const int32_t i = Si[q];
                    if (Sx[q] == 0.0 || i == j) continue;
                    while (qq < qe && Qi[qq] < i) ++qq;
                    if (qq < qe && Qi[qq] == i && Qx[qq] != 0.0) continue;
                    keep[q] = 1;
                    c[i]++;
                }
            }
        });
        // per row: exclusive scan over the threads (its start inside the row), total into Tp
        parallel_rows(n, [&](int64_t i0, int64_t i1) {
            for (int64_t i = i0; i < i1; ++i) {
                int32_t run = 0;
                for (unsigned t = 0; t < nth; ++t) {
                    const int32_t c = cnt[t][i];
                    cnt[t][i] = run;
                    run += c;
                }
                Tp[i + 1] = run;
            }
        });
        for (int64_t i = 0; i < n; ++i) Tp[i + 1] += Tp[i];
        Ti.resize(Tp[n]);
        Tx.resize(Tp[n]);
        run_threads([&](unsigned t) {
            std::vector<int32_t>& c = cnt[t];
            const int64_t j0 = t * chunk, j1 = std::min<int64_t>(n, j0 + chunk);
            for (int64_t j = j0; j < j1; ++j)
                for (int32_t q = Sp[j]; q < Sp[j + 1]; ++q) {
                    if (!keep[q]) continue;
                    const int32_t i = Si[q];
                    const int32_t f = Tp[i] + c[i]++;
                    Ti[f] = (int32_t)j;
                    Tx[f] = Sx[q];
                }
        });
    }
    P.nnzT = Tp[n];
    tm.lap("  T = S^T filtered");

    // T^T needs no second transpose: row i of T^T is row i of S restricted to the kept entries
    // ---- S_sum = T 1 and sqrt((T o T) 1) (mmw.py:34-39)
    std::vector<double> S_sum(n), tnorm(n);
    parallel_rows(n, [&](int64_t i0, int64_t i1) {
        for (int64_t i = i0; i < i1; ++i) {
            double s = 0.0, s2 = 0.0;
            for (int32_t q = Tp[i]; q < Tp[i + 1]; ++q) {
                s += Tx[q];
                s2 += Tx[q] * Tx[q];
            }
            S_sum[i] = s;
            tnorm[i] = std::sqrt(s2);
        }
    });

    // ---- union pattern: diag + gain (T + T^T != 0) + asso, per row, columns ascending.
    // Each row is a 3-way merge of sorted lists; rows are independent, so they are built on
    // all host cores: pass 1 counts, pass 2 fills.
    auto build_row = [&](int64_t i, std::vector<Ent>& row) -> bool {
        row.clear();
        int32_t a = Tp[i], ae = Tp[i + 1], b = Sp[i], be = Sp[i + 1], q = Qp[i], qe = Qp[i + 1];
        bool diag_done = false;
        while (true) {
            while (q < qe && Qx[q] == 0.0) ++q;
            while (b < be && !keep[b]) ++b;
            const int32_t ca = a < ae ? Ti[a] : INT32_MAX, cb = b < be ? Si[b] : INT32_MAX;
            const int32_t cq = q < qe ? Qi[q] : INT32_MAX, cd = diag_done ? INT32_MAX : (int32_t)i;
            const int32_t c = std::min(std::min(ca, cb), std::min(cq, cd));
            if (c == INT32_MAX) break;
            int hits = 0;
            Ent e{c, 1, 0.0, 0.0};
            if (ca == c) { e.tf = Tx[a]; ++a; hits |= 1; }
            if (cb == c) { e.tb = Sx[b]; ++b; hits |= 1; }
            if (cq == c) { ++q; hits |= 2; e.kind = 2; }
            if (cd == c) { diag_done = true; hits |= 4; e.kind = 0; }
            if (hits == 1) {
                if (e.tf + e.tb == 0.0) continue;   // eliminate_zeros on T + T^T (mmw.py:54)
            } else if (hits != 2 && hits != 4) {
                return false;   // a pair that is both gain and asso (or on the diagonal)
            }
            row.push_back(e);
        }
        return true;
    };
    tm.lap("  row sums");
    std::vector<int32_t> rowptr(n + 1, 0);
    std::vector<int64_t> g_ut(n + 1, 0), a_ut(n + 1, 0);
    std::vector<int> bad(1, 0);
    parallel_rows(n, [&](int64_t r0, int64_t r1) {
        std::vector<Ent> row;
        for (int64_t i = r0; i < r1; ++i) {
            if (!build_row(i, row)) {
                bad[0] = 1;
                continue;
            }
            rowptr[i + 1] = (int32_t)row.size();
            int64_t g = 0, a = 0;
            for (const Ent& e : row) {
                g += e.col > i && e.kind == 1;
                a += e.col > i && e.kind == 2;
            }
            g_ut[i + 1] = g;
            a_ut[i + 1] = a;
        }
    });
    if (bad[0]) {
        err = "a node pair is both a gain edge and an association edge";
        return SIGSDP_EINVAL;
    }
    int64_t total = 0;
    for (int64_t i = 0; i < n; ++i) {
        P.max_row = std::max<int>(P.max_row, rowptr[i + 1]);
        total += rowptr[i + 1];
        if (total > (int64_t)INT32_MAX) {
            err = "pattern too large for int32 indices";
            return SIGSDP_EINVAL;
        }
        rowptr[i + 1] = (int32_t)total;
        g_ut[i + 1] += g_ut[i];
        a_ut[i + 1] += a_ut[i];
    }
    P.E_g = g_ut[n];
    P.E_a = a_ut[n];
    P.nnz = total;
    tm.lap("union pattern");

    // ---- node numbering of the kernels: the locality order (joined here) or the caller's
    std::vector<int32_t> perm(n), iperm(n);
    if (order != 0) {
        bfs_thread.join();
        perm.swap(bfs_perm);
        tm.lap("  clustered BFS (join)");
    } else {
        std::iota(perm.begin(), perm.end(), 0);
    }
    for (int64_t k = 0; k < n; ++k) iperm[perm[k]] = (int32_t)k;
    std::vector<int32_t> rp(n + 1, 0);
    for (int64_t k = 0; k < n; ++k) rp[k + 1] = rp[k] + (rowptr[perm[k] + 1] - rowptr[perm[k]]);

    // ---- rows of the union pattern written straight in the kernels' numbering (one pass: no
    // separate renumbering), edge ids in the reference's order (row-major upper triangle of the
    // CALLER's numbering, mmw.py:56-57): the entry whose caller row < caller column owns the id,
    // its mirror is marked and resolved in the pass below
    hvec<int32_t> col(P.nnz), eid(P.nnz);
    hvec<double> tfwd(P.nnz), tbwd(P.nnz);
    P.gi.resize(P.E_g); P.gj.resize(P.E_g); P.tij.resize(P.E_g); P.tji.resize(P.E_g);
    P.ai.resize(P.E_a); P.aj.resize(P.E_a);
    constexpr int32_t MIRROR = -2;
    parallel_rows(n, [&](int64_t k0, int64_t k1) {
        std::vector<Ent> row;
        std::vector<uint64_t> ord;
        std::vector<int32_t> ids;
        for (int64_t k = k0; k < k1; ++k) {
            const int64_t i = perm[k];
            if (k + 2 < k1) {   // the next rows' inputs are scattered over memory: start fetching them
                const int64_t i2 = perm[k + 2];
                __builtin_prefetch(Si + Sp[i2]);
                __builtin_prefetch(Sx + Sp[i2]);
                __builtin_prefetch(Ti.data() + Tp[i2]);
                __builtin_prefetch(Tx.data() + Tp[i2]);
                __builtin_prefetch(keep.data() + Sp[i2]);
            }
            build_row(i, row);
            int64_t g = g_ut[i], a = a_ut[i];
            const int m = (int)row.size();
            ids.resize(m);
            ord.resize(m);
            for (int t = 0; t < m; ++t) {
                const Ent& e = row[t];
                int32_t id = -1;
                if (e.col > i) {
                    if (e.kind == 1) {
                        P.gi[g] = (int32_t)i; P.gj[g] = e.col; P.tij[g] = e.tf; P.tji[g] = e.tb;
                        id = (int32_t)g++;
                    } else {
                        P.ai[a] = (int32_t)i; P.aj[a] = e.col;
                        id = (int32_t)(P.E_g + a++);
                    }
                } else if (e.col < i) {
                    id = MIRROR;
                }
                ids[t] = id;
                ord[t] = ((uint64_t)(uint32_t)iperm[e.col] << 32) | (uint32_t)t;
            }
            if (order != 0) std::sort(ord.begin(), ord.end());
            int32_t w = rp[k];
            for (int t = 0; t < m; ++t, ++w) {
                const int src = (int)(ord[t] & 0xffffffffu);
                col[w] = (int32_t)(ord[t] >> 32);
                eid[w] = ids[src];
                tfwd[w] = row[src].tf;
                tbwd[w] = row[src].tb;
            }
        }
    });
    tm.lap("  fill rows");
    {
        // the pattern is symmetric: a mirror entry (k, c) takes the id of (c, k), found by binary
        // search in row c (rows are independent: all host cores; only mirror slots are written,
        // only owner slots are read)
        std::atomic<int> asym{0};
        parallel_rows(n, [&](int64_t k0, int64_t k1) {
            for (int64_t k = k0; k < k1; ++k)
                for (int32_t q = rp[k]; q < rp[k + 1]; ++q) {
                    if (eid[q] != MIRROR) continue;
                    const int32_t c = col[q];
                    const int32_t* b = col.data() + rp[c];
                    const int32_t* e = col.data() + rp[c + 1];
                    const int32_t* it = std::lower_bound(b, e, (int32_t)k);
                    if (it == e || *it != k) {
                        asym.store(1);
                        return;
                    }
                    eid[q] = eid[it - col.data()];
                }
        });
        if (asym.load()) {
            err = "internal: asymmetric union pattern";
            return SIGSDP_EINVAL;
        }
    }
    tm.lap("edge ids");

    P.rowptr.swap(rp);
    P.col.swap(col);
    P.eid.swap(eid);
    P.tfwd.swap(tfwd);
    P.tbwd.swap(tbwd);
    P.S_sum.resize(n);
    P.tnorm.resize(n);
    P.h_max.resize(n);
    for (int64_t k = 0; k < n; ++k) {
        P.S_sum[k] = S_sum[perm[k]];
        P.tnorm[k] = tnorm[perm[k]];
        P.h_max[k] = h_max[perm[k]];
    }
    P.perm.swap(perm);
    P.iperm.swap(iperm);
    tm.lap("node vectors");
    P.dpos.assign(n, -1);
    P.apos.assign(P.E_a, -1);
    parallel_rows(n, [&](int64_t k0, int64_t k1) {   // every slot has exactly one writer
        for (int64_t k = k0; k < k1; ++k)
            for (int32_t q = P.rowptr[k]; q < P.rowptr[k + 1]; ++q) {
                if (P.eid[q] < 0) P.dpos[k] = q;
                else if (P.eid[q] >= P.E_g && k < P.col[q]) P.apos[P.eid[q] - P.E_g] = q;
            }
    });
    tm.lap("diag / asso positions");
    return SIGSDP_OK;
}

namespace {
// tiles of the rows [ra, rb), tile-local numbering of runs; lcol is written in place
struct TilePart {
    std::vector<int32_t> trow_end, ucnt, rend, runs;
    int umax = 0, nnzmax = 0;
    bool ok = true;
};
void build_tiles_range(const HostPlan& P, int64_t ra, int64_t rb, int max_rows, int ucap, int nnzcap, int run_gap,
                       uint16_t* lcol, TilePart& out) {
    // distinct columns of the tile under construction: one bit per node (12.5 KB at 100k nodes: cache resident);
    // the sorted list a finished tile needs falls out of a scan over the words between its smallest and largest
    // column instead of a sort (1.7 of the 4 ms this routine took at cfg4)
    std::vector<uint64_t> bits(((size_t)P.n + 63) / 64, 0);
    std::vector<int32_t> local(P.n, 0), cols, fresh;
    int64_t r0 = ra;
    while (r0 < rb) {
        cols.clear();
        int32_t cmin = INT32_MAX, cmax = -1;
        int64_t r1 = r0;
        while (r1 < rb && r1 - r0 < max_rows) {
            // distinct columns row r1 would add
            fresh.clear();
            for (int32_t q = P.rowptr[r1]; q < P.rowptr[r1 + 1]; ++q) {
                const int32_t c = P.col[q];
                uint64_t& w = bits[(size_t)c >> 6];
                const uint64_t m = (uint64_t)1 << (c & 63);
                if (!(w & m)) {
                    w |= m;
                    fresh.push_back(c);
                }
            }
            const bool fits = (int)(cols.size() + fresh.size()) <= ucap && P.rowptr[r1 + 1] - P.rowptr[r0] <= nnzcap;
            if (!fits) {
                for (int32_t c : fresh) bits[(size_t)c >> 6] &= ~((uint64_t)1 << (c & 63));
                break;
            }
            for (int32_t c : fresh) {
                cmin = std::min(cmin, c);
                cmax = std::max(cmax, c);
            }
            cols.insert(cols.end(), fresh.begin(), fresh.end());
            ++r1;
        }
        if (r1 == r0) {   // a single row exceeds the caps
            out.ok = false;
            return;
        }
        {   // cols, ascending; the bits are cleared on the way
            size_t k = 0;
            for (size_t wi = (size_t)cmin >> 6; wi <= ((size_t)cmax >> 6); ++wi) {
                uint64_t w = bits[wi];
                bits[wi] = 0;
                while (w) {
                    cols[k++] = (int32_t)(wi * 64 + (size_t)__builtin_ctzll(w));
                    w &= w - 1;
                }
            }
        }
        // runs of (nearly) consecutive columns, one bulk copy each: columns separated by at most
        // `gap` unneeded rows share a run (fewer, larger copies at the price of a few extra rows);
        // the gap shrinks until the copied rows fit the cap
        int gap = run_gap;
        int copied = 0;
        for (;; gap /= 2) {
            copied = 0;
            for (size_t i = 0; i < cols.size();) {
                size_t j = i + 1;
                while (j < cols.size() && cols[j] - cols[j - 1] <= gap + 1) ++j;
                copied += cols[j - 1] - cols[i] + 1;
                i = j;
            }
            if (copied <= ucap || gap == 0) break;
        }
        int slot = 0;
        for (size_t i = 0; i < cols.size();) {
            size_t j = i + 1;
            while (j < cols.size() && cols[j] - cols[j - 1] <= gap + 1) ++j;
            const int len = cols[j - 1] - cols[i] + 1;
            for (size_t k = i; k < j; ++k) local[cols[k]] = slot + (cols[k] - cols[i]);
            out.runs.push_back(cols[i]);
            out.runs.push_back(slot);
            out.runs.push_back(len);
            out.runs.push_back(0);
            slot += len;
            i = j;
        }
        for (int32_t q = P.rowptr[r0]; q < P.rowptr[r1]; ++q) lcol[q] = (uint16_t)local[P.col[q]];
        out.rend.push_back((int32_t)(out.runs.size() / 4));
        out.ucnt.push_back((int32_t)copied);
        out.trow_end.push_back((int32_t)r1);
        out.umax = std::max<int>(out.umax, copied);
        out.nnzmax = std::max<int>(out.nnzmax, P.rowptr[r1] - P.rowptr[r0]);
        r0 = r1;
    }
}
}  // namespace

void build_tiles(const HostPlan& P, int max_rows, int ucap, int nnzcap, HostTiles& T) {
    int run_gap = 2;
    if (const char* e = getenv("SIGSDP_RUN_GAP")) run_gap = std::max(0, atoi(e));
    const int64_t n = P.n;
    StageTimer tm;
    T = HostTiles();
    T.max_rows = max_rows;
    T.ucap = std::min(ucap, 65535);
    T.nnzcap = nnzcap;
    T.lcol.resize(P.nnz + 16);      // (not zero-filled: the range threads below touch their parts first)
    std::fill(T.lcol.begin() + P.nnz, T.lcol.end(), (uint16_t)0);
    // Large graphs are cut into a FIXED number of row ranges tiled independently on the host
    // cores (fixed, not the core count: the tiling, and with it the order of the per-block
    // partial sums, is the same on every machine); a range boundary only ends a tile early.
    const int nchunks = n >= 32768 ? 16 : 1;
    std::vector<TilePart> parts(nchunks);
    tm.lap("  tiles: lcol alloc");
    run_tasks(std::min<unsigned>(host_threads(), (unsigned)nchunks), nchunks, [&](int64_t c) {
        build_tiles_range(P, n * c / nchunks, n * (c + 1) / nchunks, max_rows, T.ucap, nnzcap, run_gap, T.lcol.data(), parts[c]);
    });
    tm.lap("  tiles: ranges");
    T.trow.push_back(0);
    T.rptr.push_back(0);
    for (const TilePart& pt : parts) {
        if (!pt.ok) return;   // T.ok stays false: the caller falls back to the gather kernels
        const int32_t run0 = (int32_t)(T.runs.size() / 4);
        T.runs.insert(T.runs.end(), pt.runs.begin(), pt.runs.end());
        for (int32_t e : pt.rend) T.rptr.push_back(run0 + e);
        T.ucnt.insert(T.ucnt.end(), pt.ucnt.begin(), pt.ucnt.end());
        T.trow.insert(T.trow.end(), pt.trow_end.begin(), pt.trow_end.end());
        T.umax = std::max(T.umax, pt.umax);
        T.nnzmax = std::max(T.nnzmax, pt.nnzmax);
    }
    const int t = (int)T.ucnt.size();
    T.ntiles = t;
    T.trec.resize((size_t)t * 8);
    for (int i = 0; i < t; ++i) {
        int32_t* r = T.trec.data() + (size_t)i * 8;
        r[0] = T.trow[i];
        r[1] = T.trow[i + 1];
        r[2] = P.rowptr[T.trow[i]];
        r[3] = P.rowptr[T.trow[i + 1]];
        r[4] = T.rptr[i];
        r[5] = T.rptr[i + 1];
        r[6] = T.ucnt[i];
        r[7] = 0;
    }
    T.ok = true;
    tm.lap("  tiles: merge");
}

int round_greedy_host(int64_t n, int Z, const int32_t* Sp, const int32_t* Si, const double* Sx,
                      const int32_t* Qp, const int32_t* Qi, const double* Qx, const double* h_max,
                      const int32_t* rank, const int32_t* pref, int32_t* z_vec, int64_t* remainder,
                      std::string& err) {
    if (n <= 0 || Z <= 0 || !Sp || !Si || !Sx || !Qp || !Qi || !Qx || !h_max || !rank || !pref || !z_vec || !remainder) {
        err = "bad argument";
        return SIGSDP_EINVAL;
    }
    // gain_sum[z][c]: interference already committed into slot z at user c; asso_sum alike
    std::vector<double> gain_sum((size_t)Z * n, 0.0), asso_sum((size_t)Z * n, 0.0);
    std::vector<int32_t> slot(n, -1);
    int64_t rem = 0;
    for (int64_t kk = 0; kk < n; ++kk) {
        int32_t k = rank[kk];
        if (k < 0 || k >= n) {
            err = "rank entry out of range";
            return SIGSDP_EINVAL;
        }
        int32_t chosen = -1;
        for (int zz = 0; zz < Z && chosen < 0; ++zz) {
            int32_t z = pref[(size_t)k * Z + zz];
            if (z < 0 || z >= Z) {
                err = "pref entry out of range";
                return SIGSDP_EINVAL;
            }
            double* gs = gain_sum.data() + (size_t)z * n;
            double* as = asso_sum.data() + (size_t)z * n;
            bool vio = gs[k] > h_max[k];  // S[k,k] is zeroed (sdp_solver.py:33)
            for (int32_t q = Sp[k]; q < Sp[k + 1] && !vio; ++q) {
                int32_t c = Si[q];
                if (c == k || Sx[q] == 0.0 || slot[c] != z) continue;
                vio = gs[c] + Sx[q] > h_max[c];
            }
            if (vio) continue;
            vio = as[k] >= 1.0;
            for (int32_t q = Qp[k]; q < Qp[k + 1] && !vio; ++q) {
                int32_t c = Qi[q];
                if (slot[c] != z) continue;
                vio = as[c] + Qx[q] >= 1.0;
            }
            if (vio) continue;
            for (int32_t q = Sp[k]; q < Sp[k + 1]; ++q)
                if (Si[q] != k) gs[Si[q]] += Sx[q];
            for (int32_t q = Qp[k]; q < Qp[k + 1]; ++q) as[Qi[q]] += Qx[q];
            chosen = z;
        }
        slot[k] = chosen;
        if (chosen < 0) ++rem;
    }
    for (int64_t k = 0; k < n; ++k) z_vec[k] = slot[k];
    *remainder = rem;
    return SIGSDP_OK;
}

}  // namespace sigsdp
