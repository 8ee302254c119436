// Graph plan built on the DEVICE (same result as build_host_plan in plan_host.cpp, array for array):
// mmw._process_state (mmw.py:26-41) and the edge-list set-up (mmw.py:52-57) as sorts, scans and
// one-thread-per-row merges, so that the only host work left on the critical path of a large
// graph's set-up is the (sequential) locality ordering, which runs next to it on its own thread.
#include <cub/cub.cuh>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "../../include/sigsdp_mmw.h"
#include "plan_device.h"

namespace sigsdp {
namespace {

constexpr int MIRROR = -2;
constexpr unsigned long long DROPPED = ~0ull;

// S entry (j, i) becomes T[i][j] unless i == j, the value is zero, or Q[j][i] != 0 (mmw.py:28-33)
__global__ void k_keep(int n, const int* Sp, const int* Si, const double* Sx, const int* Qp, const int* Qi, const double* Qx,
                       unsigned char* keep, unsigned long long* key, unsigned long long* nkept) {
    unsigned long long cnt = 0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int qb = Qp[j], qe = Qp[j + 1];
        for (int q = Sp[j]; q < Sp[j + 1]; ++q) {
            const int i = Si[q];
            bool k = Sx[q] != 0.0 && i != j;
            if (k && qb < qe) {
                int b = qb, e = qe;
                while (b < e) {
                    const int m = (b + e) >> 1;
                    if (Qi[m] < i) b = m + 1; else e = m;
                }
                if (b < qe && Qi[b] == i && Qx[b] != 0.0) k = false;
            }
            keep[q] = k ? 1 : 0;
            key[q] = k ? (((unsigned long long)(unsigned)i << 32) | (unsigned)j) : DROPPED;
            cnt += k;
        }
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(nkept, cnt);
}
// T in CSR from the sorted (row, col) keys: Tp by binary search, Ti = low word
__global__ void k_t_rows(int n, long long nT, const unsigned long long* key, int* Tp, int* Ti) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nT; t += (long long)gridDim.x * blockDim.x)
        Ti[t] = (int)(key[t] & 0xffffffffull);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x) {
        const unsigned long long want = (unsigned long long)(unsigned)i << 32;
        long long b = 0, e = nT;
        while (b < e) {
            const long long m = (b + e) >> 1;
            if (key[m] < want) b = m + 1; else e = m;
        }
        Tp[i] = (int)b;
    }
}
// S_sum = T 1 and sqrt((T o T) 1), summed in column order like the host builder (mmw.py:34-39)
__global__ void k_t_sums(int n, const int* Tp, const double* Tx, double* S_sum, double* tnorm) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double s = 0.0, s2 = 0.0;
        for (int q = Tp[i]; q < Tp[i + 1]; ++q) {
            s = __dadd_rn(s, Tx[q]);
            s2 = __dadd_rn(s2, __dmul_rn(Tx[q], Tx[q]));   // (no fma: the host builder rounds the product)
        }
        S_sum[i] = s;
        tnorm[i] = sqrt(s2);
    }
}

// One row of the union pattern: 4-way merge of T row i, the kept part of S row i (= column i of T), Q row i and the
// diagonal, columns ascending.  emit(col, kind, tf, tb) with kind 0 diag, 1 gain, 2 asso; returns false when a pair is
// both gain and asso / on the diagonal.
struct RowSrc {
    const int *Tp, *Ti, *Sp, *Si, *Qp, *Qi;
    const double *Tx, *Sx, *Qx;
    const unsigned char* keep;
};
template <class Emit>
__device__ __forceinline__ bool merge_row(const RowSrc& R, int i, Emit&& emit) {
    int a = R.Tp[i], ae = R.Tp[i + 1], b = R.Sp[i], be = R.Sp[i + 1], q = R.Qp[i], qe = R.Qp[i + 1];
    bool diag_done = false;
    const int BIG = 0x7fffffff;
    for (;;) {
        while (q < qe && R.Qx[q] == 0.0) ++q;
        while (b < be && !R.keep[b]) ++b;
        const int ca = a < ae ? R.Ti[a] : BIG, cb = b < be ? R.Si[b] : BIG;
        const int cq = q < qe ? R.Qi[q] : BIG, cd = diag_done ? BIG : i;
        const int c = min(min(ca, cb), min(cq, cd));
        if (c == BIG) break;
        int hits = 0, kind = 1;
        double tf = 0.0, tb = 0.0;
        if (ca == c) { tf = R.Tx[a]; ++a; hits |= 1; }
        if (cb == c) { tb = R.Sx[b]; ++b; hits |= 1; }
        if (cq == c) { ++q; hits |= 2; kind = 2; }
        if (cd == c) { diag_done = true; hits |= 4; kind = 0; }
        if (hits == 1) {
            if (tf + tb == 0.0) continue;   // eliminate_zeros on T + T^T (mmw.py:54)
        } else if (hits != 2 && hits != 4) {
            return false;
        }
        emit(c, kind, tf, tb);
    }
    return true;
}
__global__ void k_union_count(int n, RowSrc R, int* len, int* gcnt, int* acnt, int* flags /* [0] bad, [1] max row */) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int l = 0, g = 0, a = 0;
        const bool ok = merge_row(R, i, [&](int c, int kind, double, double) {
            ++l;
            g += c > i && kind == 1;
            a += c > i && kind == 2;
        });
        if (!ok) flags[0] = 1;
        len[i] = l;
        gcnt[i] = g;
        acnt[i] = a;
        atomicMax(&flags[1], l);
    }
}
// rows in the caller's numbering; the entry with row < col owns the edge id (reference order: row-major upper
// triangle, mmw.py:56-57), its mirror is marked and resolved by k_mirror
__global__ void k_union_fill(int n, RowSrc R, const int* rowptr, const int* g0, const int* a0, int E_g, int* col, int* eid, double* tfwd,
                             double* tbwd, int* gi, int* gj, double* tij, double* tji, int* ai, int* aj) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int w = rowptr[i], g = g0[i], a = a0[i];
        merge_row(R, i, [&](int c, int kind, double tf, double tb) {
            int id = -1;
            if (c > i) {
                if (kind == 1) {
                    gi[g] = i; gj[g] = c; tij[g] = tf; tji[g] = tb;
                    id = g++;
                } else {
                    ai[a] = i; aj[a] = c;
                    id = E_g + a++;
                }
            } else if (c < i) {
                id = MIRROR;
            }
            col[w] = c;
            eid[w] = id;
            tfwd[w] = tf;
            tbwd[w] = tb;
            ++w;
        });
    }
}
__global__ void k_mirror(int n, const int* rowptr, const int* col, int* eid, int* flags) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        for (int q = rowptr[i]; q < rowptr[i + 1] && col[q] < i; ++q) {
            const int c = col[q];
            int b = rowptr[c], e = rowptr[c + 1];
            while (b < e) {
                const int m = (b + e) >> 1;
                if (col[m] < i) b = m + 1; else e = m;
            }
            if (b >= rowptr[c + 1] || col[b] != i) {
                flags[0] = 2;   // asymmetric union pattern: cannot happen for validated inputs
                continue;
            }
            eid[q] = eid[b];
        }
}
// renumbering: new row k = old row perm[k]; its entries keyed by the new column, to be sorted per row
__global__ void k_new_len(int n, const int* perm, const int* rowptr, int* newlen) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) newlen[k] = rowptr[perm[k] + 1] - rowptr[perm[k]];
}
__global__ void k_new_keys(int n, const int* perm, const int* iperm, const int* rowptr, const int* col, const int* rp, int* keys, int* src) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int o = perm[k];
        int w = rp[k];
        for (int q = rowptr[o]; q < rowptr[o + 1]; ++q, ++w) {
            keys[w] = iperm[col[q]];
            src[w] = q;
        }
    }
}
__global__ void k_gather(long long nnz, const int* src, const int* eid0, const double* tf0, const double* tb0, int* eid, double* tfwd, double* tbwd) {
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < nnz; p += (long long)gridDim.x * blockDim.x) {
        const int q = src[p];
        eid[p] = eid0[q];
        tfwd[p] = tf0[q];
        tbwd[p] = tb0[q];
    }
}
__global__ void k_node_vectors(int n, const int* perm, const double* S0, const double* t0, const double* h0, double* S_sum, double* tnorm, double* h_max) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int o = perm ? perm[k] : k;
        S_sum[k] = S0[o];
        tnorm[k] = t0[o];
        h_max[k] = h0[o];
    }
}
__global__ void k_positions(int n, int E_g, const int* rowptr, const int* col, const int* eid, int* dpos, int* apos) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
        for (int q = rowptr[k]; q < rowptr[k + 1]; ++q) {
            if (eid[q] < 0) dpos[k] = q;
            else if (eid[q] >= E_g && k < col[q]) apos[eid[q] - E_g] = q;
        }
}

struct DevTimer {   // SIGSDP_PLAN_TIMING=1: stage times on stderr (synchronises: only for diagnosis)
    bool on = getenv("SIGSDP_PLAN_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void lap(const char* what) {
        if (!on) return;
        cudaDeviceSynchronize();
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[dplan] %-26s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// Temporaries of one build come from a process-wide workspace that is kept for the next build (a bump allocator over one
// cudaMalloc block): a binary search or a benchmark loop builds plans of the same size over and over, and fresh
// stream-ordered allocations of ~300 MB cost 6-28 ms per plan at 100k nodes (pool growth) against ~2 ms of kernels.
// What does not fit goes to cudaMallocAsync for this build and the block is re-sized afterwards.  Builds are serialised
// by the caller (the API's staging lock).
struct Workspace {
    char* base = nullptr;
    size_t cap = 0;
    int device = -1;
};
Workspace g_ws;
struct Tmp {
    cudaStream_t st;
    size_t used = 0, need = 0;
    std::vector<void*> overflow;
    Tmp(cudaStream_t s, int device) : st(s) {
        if (g_ws.device != device) {
            if (g_ws.base) cudaFree(g_ws.base);
            g_ws = Workspace();
            g_ws.device = device;
        }
    }
    template <class U> cudaError_t get(U** p, size_t count) {
        const size_t bytes = ((count ? count : 1) * sizeof(U) + 255) & ~(size_t)255;
        need += bytes;
        if (used + bytes <= g_ws.cap) {
            *p = reinterpret_cast<U*>(g_ws.base + used);
            used += bytes;
            return cudaSuccess;
        }
        void* q = nullptr;
        cudaError_t e = cudaMallocAsync(&q, bytes, st);
        if (e == cudaSuccess) overflow.push_back(q);
        *p = static_cast<U*>(q);
        return e;
    }
    ~Tmp() {
        for (void* p : overflow) cudaFreeAsync(p, st);
        if (need <= g_ws.cap) return;
        size_t limit = (size_t)4096 << 20;
        if (const char* e = getenv("SIGSDP_PLAN_WORKSPACE_MB")) limit = (size_t)std::max(0, atoi(e)) << 20;
        if (g_ws.base) cudaFree(g_ws.base);   // (synchronises: every kernel of this build is done)
        g_ws.base = nullptr;
        g_ws.cap = 0;
        const size_t want = need + need / 16;
        if (want <= limit && cudaMalloc((void**)&g_ws.base, want) == cudaSuccess) g_ws.cap = want;
        else cudaGetLastError();
    }
};

#define DK(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            err = std::string(#call) + ": " + cudaGetErrorString(e_);                    \
            return SIGSDP_ECUDA;                                                         \
        }                                                                                \
    } while (0)

}  // namespace

int build_device_plan(int64_t n64, const int32_t* Sp, const int32_t* Si, const double* Sx, const int32_t* Qp, const int32_t* Qi,
                      const double* Qx, const double* h_max, int order, const DevicePlanAlloc& alloc, HostPlan& P, DevicePlanArrays& D,
                      std::string& err) {
    {
        const int rc = validate_pointers(n64, Sp, Si, Sx, Qp, Qi, Qx, h_max, err);
        if (rc != SIGSDP_OK) return rc;
    }
    if ((int64_t)Sp[n64] + Qp[n64] + 4 * n64 > (int64_t)0x7ffffff0) {
        err = "pattern too large for int32 indices";
        return SIGSDP_EINVAL;
    }
    DevTimer tm;
    const int n = (int)n64;
    const long long nS = Sp[n], nQ = Qp[n];
    cudaStream_t st = (cudaStream_t)0;
    // the locality ordering: sequential, on its own thread, next to everything below (it guards its own index reads:
    // the index checks of validate_state run beside it)
    std::vector<int32_t> perm_h;
    std::thread bfs;
    if (order != 0) bfs = std::thread([&] { locality_order_of_inputs(n, Sp, Si, Qp, Qi, order > 1 ? order : 64, perm_h); });
    struct Joiner {
        std::thread& t;
        ~Joiner() { if (t.joinable()) t.join(); }
    } joiner{bfs};
    {
        const int rc = validate_state(n64, Sp, Si, Sx, Qp, Qi, Qx, h_max, err);
        if (rc != SIGSDP_OK) return rc;
    }
    tm.lap("validate");
    P = HostPlan();
    P.n = n;
    P.order = order;

    int device = 0;
    DK(cudaGetDevice(&device));
    Tmp tmp(st, device);
    // ---- inputs to the device (one staged copy; a few elements of padding behind every array)
    int *dSp, *dSi, *dQp, *dQi;
    double *dSx, *dQx, *dh;
    {
        const size_t pad = 64;
        size_t off = 0;
        auto place = [&](size_t bytes) { const size_t o = off; off += (bytes + pad + 255) & ~(size_t)255; return o; };
        const size_t oSp = place((n + 1) * 4), oSi = place(nS * 4), oSx = place(nS * 8), oQp = place((n + 1) * 4), oQi = place(nQ * 4),
                     oQx = place(nQ * 8), oh = place((size_t)n * 8);
        char* dbase;
        DK(tmp.get(&dbase, off));
        char* stage = static_cast<char*>(alloc.pinned(alloc.ctx, off));
        if (!stage) { err = "pinned staging buffer"; return SIGSDP_ENOMEM; }
        parallel_copy({{stage + oSp, Sp, (size_t)(n + 1) * 4}, {stage + oSi, Si, (size_t)nS * 4}, {stage + oSx, Sx, (size_t)nS * 8},
                       {stage + oQp, Qp, (size_t)(n + 1) * 4}, {stage + oQi, Qi, (size_t)nQ * 4}, {stage + oQx, Qx, (size_t)nQ * 8},
                       {stage + oh, h_max, (size_t)n * 8}});
        DK(cudaMemcpyAsync(dbase, stage, off, cudaMemcpyHostToDevice, st));
        dSp = (int*)(dbase + oSp); dSi = (int*)(dbase + oSi); dSx = (double*)(dbase + oSx);
        dQp = (int*)(dbase + oQp); dQi = (int*)(dbase + oQi); dQx = (double*)(dbase + oQx); dh = (double*)(dbase + oh);
    }
    const int blocks = alloc.num_sms * 8, T256 = 256;
    tm.lap("inputs to device");

    // ---- T = S^T filtered: keys (row of T, column of T) of the kept entries, sorted
    unsigned char* keep;
    unsigned long long *key0, *key1, *counters;
    double* Tx;
    DK(tmp.get(&keep, nS));
    DK(tmp.get(&key0, nS));
    DK(tmp.get(&key1, nS));
    DK(tmp.get(&Tx, nS));
    DK(tmp.get(&counters, 4));
    DK(cudaMemsetAsync(counters, 0, 4 * sizeof(unsigned long long), st));
    k_keep<<<blocks, T256, 0, st>>>(n, dSp, dSi, dSx, dQp, dQi, dQx, keep, key0, counters);
    {
        size_t tb = 0;
        DK(cub::DeviceRadixSort::SortPairs(nullptr, tb, key0, key1, dSx, Tx, (int)nS, 0, 64, st));
        char* t;
        DK(tmp.get(&t, tb));
        DK(cub::DeviceRadixSort::SortPairs(t, tb, key0, key1, dSx, Tx, (int)nS, 0, 64, st));
    }
    unsigned long long hc[4];
    if (alloc.overlap) alloc.overlap();
    DK(cudaMemcpyAsync(hc, counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
    DK(cudaStreamSynchronize(st));
    const long long nT = (long long)hc[0];
    tm.lap("keep + sort (+ host copies)");
    P.nnzT = nT;
    int *Tp, *Ti;
    DK(tmp.get(&Tp, n + 1));
    DK(tmp.get(&Ti, nT + 1));
    k_t_rows<<<blocks, T256, 0, st>>>(n, nT, key1, Tp, Ti);
    double *S0, *t0;
    DK(tmp.get(&S0, n));
    DK(tmp.get(&t0, n));
    k_t_sums<<<blocks, T256, 0, st>>>(n, Tp, Tx, S0, t0);

    // ---- union pattern: counts, scans, fill, mirror ids (caller numbering)
    RowSrc R{Tp, Ti, dSp, dSi, dQp, dQi, Tx, dSx, dQx, keep};
    int *len, *gcnt, *acnt, *rowptr0, *g0, *a0, *flags;
    DK(tmp.get(&len, n + 1));
    DK(tmp.get(&gcnt, n + 1));
    DK(tmp.get(&acnt, n + 1));
    DK(tmp.get(&rowptr0, n + 1));
    DK(tmp.get(&g0, n + 1));
    DK(tmp.get(&a0, n + 1));
    DK(tmp.get(&flags, 4));
    DK(cudaMemsetAsync(flags, 0, 4 * sizeof(int), st));
    DK(cudaMemsetAsync(len + n, 0, sizeof(int), st));
    DK(cudaMemsetAsync(gcnt + n, 0, sizeof(int), st));
    DK(cudaMemsetAsync(acnt + n, 0, sizeof(int), st));
    k_union_count<<<blocks, T256, 0, st>>>(n, R, len, gcnt, acnt, flags);
    char* scan_tmp;
    size_t scan_bytes = 0;
    DK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, len, rowptr0, n + 1, st));
    DK(tmp.get(&scan_tmp, scan_bytes));
    DK(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, len, rowptr0, n + 1, st));
    DK(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, gcnt, g0, n + 1, st));
    DK(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, acnt, a0, n + 1, st));
    int totals[3], hflags[4];
    DK(cudaMemcpyAsync(&totals[0], rowptr0 + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    DK(cudaMemcpyAsync(&totals[1], g0 + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    DK(cudaMemcpyAsync(&totals[2], a0 + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    DK(cudaMemcpyAsync(hflags, flags, sizeof(hflags), cudaMemcpyDeviceToHost, st));
    DK(cudaStreamSynchronize(st));
    if (hflags[0]) {
        err = "a node pair is both a gain edge and an association edge";
        return SIGSDP_EINVAL;
    }
    tm.lap("T rows, sums, union count");
    P.nnz = totals[0];
    P.E_g = totals[1];
    P.E_a = totals[2];
    P.max_row = hflags[1];
    const long long nnz = P.nnz;
    const int E_g = (int)P.E_g, E_a = (int)P.E_a;

    // persistent arrays: one slab owned by the plan
    size_t off = 0;
    auto place = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_rowptr = place((size_t)(n + 1) * 4), o_col = place(nnz * 4), o_eid = place(nnz * 4), o_tfwd = place(nnz * 8),
                 o_tbwd = place(nnz * 8), o_ssum = place((size_t)n * 8), o_tnorm = place((size_t)n * 8), o_hm = place((size_t)n * 8),
                 o_perm = place((size_t)n * 4), o_dpos = place((size_t)n * 4), o_apos = place((size_t)E_a * 4), o_gi = place((size_t)E_g * 4),
                 o_gj = place((size_t)E_g * 4), o_tij = place((size_t)E_g * 8), o_tji = place((size_t)E_g * 8), o_ai = place((size_t)E_a * 4),
                 o_aj = place((size_t)E_a * 4);
    char* slab = static_cast<char*>(alloc.device(alloc.ctx, off));
    if (!slab) { err = "plan slab"; return SIGSDP_ENOMEM; }
    D.rowptr = (int*)(slab + o_rowptr); D.col = (int*)(slab + o_col); D.eid = (int*)(slab + o_eid);
    D.tfwd = (double*)(slab + o_tfwd); D.tbwd = (double*)(slab + o_tbwd); D.S_sum = (double*)(slab + o_ssum);
    D.tnorm = (double*)(slab + o_tnorm); D.h_max = (double*)(slab + o_hm); D.perm = order != 0 ? (int*)(slab + o_perm) : nullptr;
    D.dpos = (int*)(slab + o_dpos); D.apos = (int*)(slab + o_apos); D.gi = (int*)(slab + o_gi); D.gj = (int*)(slab + o_gj);
    D.tij = (double*)(slab + o_tij); D.tji = (double*)(slab + o_tji); D.ai = (int*)(slab + o_ai); D.aj = (int*)(slab + o_aj);

    // caller-numbering rows: straight into the final arrays when no renumbering follows
    int *col0 = D.col, *eid0 = D.eid;
    double *tf0 = D.tfwd, *tb0 = D.tbwd;
    if (order != 0) {
        DK(tmp.get(&col0, nnz));
        DK(tmp.get(&eid0, nnz));
        DK(tmp.get(&tf0, nnz));
        DK(tmp.get(&tb0, nnz));
    }
    tm.lap("slab + temporaries");
    k_union_fill<<<blocks, T256, 0, st>>>(n, R, rowptr0, g0, a0, E_g, col0, eid0, tf0, tb0, D.gi, D.gj, D.tij, D.tji, D.ai, D.aj);
    tm.lap("union fill");
    k_mirror<<<blocks, T256, 0, st>>>(n, rowptr0, col0, eid0, flags);

    // the host copies of the pattern are allocated and their pages touched now, on the idle cores, while the device works
    // and the locality ordering is still running: the copy at the end then lands in mapped memory
    P.rowptr.resize(n + 1);
    P.col.resize(nnz);
    P.dpos.resize(n);
    P.S_sum.resize(n);
    P.tnorm.resize(n);
    P.h_max.resize(n);
    parallel_for((nnz + 1023) / 1024, [&](int64_t b, int64_t e) {
        for (int64_t i = b; i < e; ++i) P.col[(size_t)i * 1024] = 0;
    }, 64);
    tm.lap("mirror ids");
    // ---- the kernels' numbering
    P.perm.resize(n);
    P.iperm.resize(n);
    if (order != 0) {
        bfs.join();
        tm.lap("clustered BFS (join)");
        P.perm.swap(perm_h);
        for (int k = 0; k < n; ++k) P.iperm[P.perm[k]] = k;
        int* iperm_d;
        DK(tmp.get(&iperm_d, n));
        DK(cudaMemcpyAsync(D.perm, P.perm.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
        DK(cudaMemcpyAsync(iperm_d, P.iperm.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
        int *newlen, *keys, *src, *src_sorted;
        DK(tmp.get(&newlen, n + 1));
        DK(tmp.get(&keys, nnz));
        DK(tmp.get(&src, nnz));
        DK(tmp.get(&src_sorted, nnz));
        DK(cudaMemsetAsync(newlen + n, 0, sizeof(int), st));
        k_new_len<<<blocks, T256, 0, st>>>(n, D.perm, rowptr0, newlen);
        DK(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, newlen, D.rowptr, n + 1, st));
        k_new_keys<<<blocks, T256, 0, st>>>(n, D.perm, iperm_d, rowptr0, col0, D.rowptr, keys, src);
        size_t sb = 0;
        DK(cub::DeviceSegmentedSort::SortPairs(nullptr, sb, keys, D.col, src, src_sorted, nnz, n, D.rowptr, D.rowptr + 1, st));
        char* t;
        DK(tmp.get(&t, sb));
        DK(cub::DeviceSegmentedSort::SortPairs(t, sb, keys, D.col, src, src_sorted, nnz, n, D.rowptr, D.rowptr + 1, st));
        k_gather<<<blocks, T256, 0, st>>>(nnz, src_sorted, eid0, tf0, tb0, D.eid, D.tfwd, D.tbwd);
    } else {
        for (int k = 0; k < n; ++k) P.perm[k] = P.iperm[k] = k;
        DK(cudaMemcpyAsync(D.rowptr, rowptr0, (size_t)(n + 1) * 4, cudaMemcpyDeviceToDevice, st));
    }
    k_node_vectors<<<blocks, T256, 0, st>>>(n, D.perm, S0, t0, dh, D.S_sum, D.tnorm, D.h_max);
    DK(cudaMemsetAsync(D.apos, 0xff, (size_t)E_a * 4, st));
    k_positions<<<blocks, T256, 0, st>>>(n, E_g, D.rowptr, D.col, D.eid, D.dpos, D.apos);

    tm.lap("renumber + positions");
    // ---- what the host needs right away (tiles, solver set-up): pattern, diagonal positions, node vectors
    {
        size_t ho = 0;
        auto hplace = [&](size_t bytes) { const size_t o = ho; ho += (bytes + 255) & ~(size_t)255; return o; };
        const size_t h_rowptr = hplace((size_t)(n + 1) * 4), h_col = hplace(nnz * 4), h_dpos = hplace((size_t)n * 4),
                     h_ssum = hplace((size_t)n * 8), h_tnorm = hplace((size_t)n * 8), h_hm = hplace((size_t)n * 8);
        char* stage = static_cast<char*>(alloc.pinned(alloc.ctx, ho));
        if (!stage) { err = "pinned staging buffer"; return SIGSDP_ENOMEM; }
        DK(cudaMemcpyAsync(stage + h_rowptr, D.rowptr, (size_t)(n + 1) * 4, cudaMemcpyDeviceToHost, st));
        DK(cudaMemcpyAsync(stage + h_col, D.col, nnz * 4, cudaMemcpyDeviceToHost, st));
        DK(cudaMemcpyAsync(stage + h_dpos, D.dpos, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        DK(cudaMemcpyAsync(stage + h_ssum, D.S_sum, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        DK(cudaMemcpyAsync(stage + h_tnorm, D.tnorm, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        DK(cudaMemcpyAsync(stage + h_hm, D.h_max, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        DK(cudaMemcpyAsync(hflags, flags, sizeof(hflags), cudaMemcpyDeviceToHost, st));
        DK(cudaStreamSynchronize(st));
        if (hflags[0]) {
            err = "internal: asymmetric union pattern";
            return SIGSDP_EINVAL;
        }
        parallel_copy({{P.rowptr.data(), stage + h_rowptr, (size_t)(n + 1) * 4}, {P.col.data(), stage + h_col, (size_t)nnz * 4},
                       {P.dpos.data(), stage + h_dpos, (size_t)n * 4}, {P.S_sum.data(), stage + h_ssum, (size_t)n * 8},
                       {P.tnorm.data(), stage + h_tnorm, (size_t)n * 8}, {P.h_max.data(), stage + h_hm, (size_t)n * 8}});
    }
    tm.lap("pattern to host");
    return SIGSDP_OK;
}

// the per-non-zero and per-edge arrays the host only needs for fetches, shards and images: brought over on first use
int fetch_device_plan_rest(const DevicePlanArrays& D, const DevicePlanAlloc& alloc, HostPlan& P, std::string& err) {
    cudaStream_t st = (cudaStream_t)0;
    const size_t nnz = (size_t)P.nnz, E_g = (size_t)P.E_g, E_a = (size_t)P.E_a;
    size_t ho = 0;
    auto hplace = [&](size_t bytes) { const size_t o = ho; ho += (bytes + 255) & ~(size_t)255; return o; };
    const size_t h_eid = hplace(nnz * 4), h_tf = hplace(nnz * 8), h_tb = hplace(nnz * 8), h_apos = hplace(E_a * 4), h_gi = hplace(E_g * 4),
                 h_gj = hplace(E_g * 4), h_tij = hplace(E_g * 8), h_tji = hplace(E_g * 8), h_ai = hplace(E_a * 4), h_aj = hplace(E_a * 4);
    char* stage = static_cast<char*>(alloc.pinned(alloc.ctx, ho));
    if (!stage) { err = "pinned staging buffer"; return SIGSDP_ENOMEM; }
    DK(cudaMemcpyAsync(stage + h_eid, D.eid, nnz * 4, cudaMemcpyDeviceToHost, st));
    DK(cudaMemcpyAsync(stage + h_tf, D.tfwd, nnz * 8, cudaMemcpyDeviceToHost, st));
    DK(cudaMemcpyAsync(stage + h_tb, D.tbwd, nnz * 8, cudaMemcpyDeviceToHost, st));
    DK(cudaMemcpyAsync(stage + h_apos, D.apos, E_a * 4, cudaMemcpyDeviceToHost, st));
    DK(cudaMemcpyAsync(stage + h_gi, D.gi, E_g * 4, cudaMemcpyDeviceToHost, st));
    DK(cudaMemcpyAsync(stage + h_gj, D.gj, E_g * 4, cudaMemcpyDeviceToHost, st));
    DK(cudaMemcpyAsync(stage + h_tij, D.tij, E_g * 8, cudaMemcpyDeviceToHost, st));
    DK(cudaMemcpyAsync(stage + h_tji, D.tji, E_g * 8, cudaMemcpyDeviceToHost, st));
    DK(cudaMemcpyAsync(stage + h_ai, D.ai, E_a * 4, cudaMemcpyDeviceToHost, st));
    DK(cudaMemcpyAsync(stage + h_aj, D.aj, E_a * 4, cudaMemcpyDeviceToHost, st));
    DK(cudaStreamSynchronize(st));
    P.eid.resize(nnz); P.tfwd.resize(nnz); P.tbwd.resize(nnz); P.apos.resize(E_a);
    P.gi.resize(E_g); P.gj.resize(E_g); P.tij.resize(E_g); P.tji.resize(E_g); P.ai.resize(E_a); P.aj.resize(E_a);
    parallel_copy({{P.eid.data(), stage + h_eid, nnz * 4}, {P.tfwd.data(), stage + h_tf, nnz * 8}, {P.tbwd.data(), stage + h_tb, nnz * 8},
                   {P.apos.data(), stage + h_apos, E_a * 4}, {P.gi.data(), stage + h_gi, E_g * 4}, {P.gj.data(), stage + h_gj, E_g * 4},
                   {P.tij.data(), stage + h_tij, E_g * 8}, {P.tji.data(), stage + h_tji, E_g * 8}, {P.ai.data(), stage + h_ai, E_a * 4},
                   {P.aj.data(), stage + h_aj, E_a * 4}});
    return SIGSDP_OK;
}

}  // namespace sigsdp
