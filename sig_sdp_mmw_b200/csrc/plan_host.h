// Host-side graph plan (native C++): what mmw._process_state (mmw.py:26-41) and the
// edge-list set-up (mmw.py:52-57) produce, in the flat layout the kernels walk.
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace sigsdp {

// std::vector that leaves trivially constructible elements uninitialised on resize: the
// big per-non-zero arrays are written once, in parallel, by the threads that build them,
// so the pages are first touched there instead of being zero-filled by one thread.
template <class T>
struct DefaultInitAllocator : std::allocator<T> {
    template <class U> struct rebind { using other = DefaultInitAllocator<U>; };
    using std::allocator<T>::allocator;
    template <class U> void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) {
        ::new (static_cast<void*>(p)) U;
    }
    template <class U, class... A> void construct(U* p, A&&... a) {
        ::new (static_cast<void*>(p)) U(std::forward<A>(a)...);
    }
};
template <class T> using hvec = std::vector<T, DefaultInitAllocator<T>>;

struct HostPlan {
    int64_t n = 0;
    int64_t E_g = 0, E_a = 0;   // undirected gain / association edges
    int64_t nnz = 0;            // union pattern incl. diagonal = n + 2 (E_g + E_a)
    int64_t nnzT = 0;           // stored entries of T
    int max_row = 0;
    // union pattern in INTERNAL numbering, columns ascending inside a row
    std::vector<int32_t> rowptr;
    hvec<int32_t> col, eid;                  // eid: -1 diag, [0,E_g) gain, E_g + a asso
    hvec<double> tfwd, tbwd;                 // T[row,col], T[col,row] (0 off the gain pattern)
    // edge lists in the CALLER's numbering and the reference's order
    hvec<int32_t> gi, gj, ai, aj;
    hvec<double> tij, tji;
    // node vectors, INTERNAL numbering
    std::vector<double> S_sum, tnorm, h_max;
    // perm[new] = old, iperm[old] = new
    std::vector<int32_t> perm, iperm;
    int order = 0;
    // position of each row's diagonal entry; position of the (row < col, internal
    // numbering) entry of each asso edge
    std::vector<int32_t> dpos, apos;
};

// Row tiles: consecutive rows grouped greedily so that a tile has at most `max_rows` rows,
// `ucap` distinct columns and `nnzcap` non-zeros (see TileDev in mmw_device.cuh).
// ok = false when a single row exceeds the caps (the caller falls back to gather kernels).
struct HostTiles {
    int max_rows = 0, ucap = 0, nnzcap = 0;
    int ntiles = 0, umax = 0, nnzmax = 0;
    bool ok = false;
    std::vector<int32_t> trow;   // ntiles + 1: first row of each tile
    std::vector<int32_t> ucnt;   // ntiles: distinct columns
    std::vector<int32_t> rptr;   // ntiles + 1: runs of consecutive distinct columns
    std::vector<int32_t> runs;   // 4 ints per run: first column, first slot, length, 0
    hvec<uint16_t> lcol;         // nnz (+ 16 zeros: the device copy is read in 16-byte supersets)
    std::vector<int32_t> trec;   // 8 ints per tile: r0, r1, p0, p1, run0, run1, distinct columns, 0
};
void build_tiles(const HostPlan& P, int max_rows, int ucap, int nnzcap, HostTiles& out);

// Flat image of a built plan (header + arrays), so that ONE process of a multi-GPU launch builds the plan on all
// host cores and the others receive it instead of building the same plan side by side on a share of the cores.
size_t host_plan_image_bytes(const HostPlan& P);
void host_plan_to_image(const HostPlan& P, unsigned char* buf);
// returns false (err set) when the image is malformed or does not describe an n-node plan
bool host_plan_from_image(const unsigned char* buf, size_t bytes, int64_t n, HostPlan& P, std::string& err);

// Row sharding of one graph across `nranks` GPUs (include/sigsdp_mmw.h, "row sharding"): rank r owns the rows
// [row0[r], row0[r+1]) -- whole tiles [tile0[r], tile0[r+1]) when the solver is tiled (ht != nullptr) -- cut where
// the cumulative non-zeros reach r / nranks of the total.
void shard_cut_points(const HostPlan& P, const HostTiles* ht, int nranks, std::vector<int32_t>& row0,
                      std::vector<int32_t>& tile0);
// What rank `rank` exchanges: pmask[k] (own rows: bit p set = rank p reads row k of the sketch block / r / q), the
// association edges with an entry in an own row (the first n_inc_owned are owned: their row < col entry is in an
// own row) with the position of that entry, rows pushed per Taylor term (row x destination pairs) and distinct
// foreign rows read.
struct ShardHalo {
    std::vector<uint8_t> pmask;
    std::vector<int32_t> inc_e, inc_p;
    int n_inc_owned = 0;
    long long send = 0, recv = 0;
};
void shard_halo(const HostPlan& P, const std::vector<int32_t>& row0, int rank, ShardHalo& out);

// Copies a list of (dst, src, bytes) segments on the host's cores (1 MB chunks dealt to the
// builder's threads): staging the plan arrays into pinned memory takes ~1.5 ms instead of ~7.
struct CopySeg {
    void* dst;
    const void* src;
    size_t bytes;
};
void parallel_copy(const std::vector<CopySeg>& segs);
// a sequential helper thread starts / ends next to the parallel stages (they leave it a core)
void side_thread_begin();
void side_thread_end();
// fn(begin, end) over contiguous chunks of [0, n) on the builder's threads (serial below min_parallel items)
void parallel_for(int64_t n, const std::function<void(int64_t, int64_t)>& fn, int64_t min_parallel = 4096);

// numpy's legacy normal stream continued natively (numpy_stream.cpp): `count` numbers np.random.standard_normal would return
// next, bit for bit; key (624 words) / pos / has_gauss / gauss are RandomState.get_state() on entry, the state after the draw on exit
void numpy_legacy_normals(uint32_t* key, int32_t* pos, int32_t* has_gauss, double* gauss, int64_t count, double* out);

// checksum of a buffer on the builder's threads (the plan cache's key, sdp_solver._plan_for)
uint64_t checksum_bytes(const void* data, size_t bytes);

// structure checks of the inputs (monotone row pointers, sorted duplicate-free in-range indices, Q_asso symmetric with an
// empty diagonal); 0 or a negative SIGSDP_E* code with the message in err
// the part of validate_state that must hold before a row may be dereferenced at all (non-null arrays, monotone row
// pointers): the locality ordering starts after it, next to the remaining checks
int validate_pointers(int64_t n, const int32_t* Sp, const int32_t* Si, const double* Sx, const int32_t* Qp, const int32_t* Qi,
                      const double* Qx, const double* h_max, std::string& err);
int validate_state(int64_t n, const int32_t* Sp, const int32_t* Si, const double* Sx, const int32_t* Qp, const int32_t* Qi,
                   const double* Qx, const double* h_max, std::string& err);
// the locality ordering (clustered BFS over the rows of S and Q), perm[new] = old
void locality_order_of_inputs(int64_t n, const int32_t* Sp, const int32_t* Si, const int32_t* Qp, const int32_t* Qi, int cluster,
                              std::vector<int32_t>& perm);
// returns 0 or a negative SIGSDP_E* code; err gets the message
int build_host_plan(int64_t n, const int32_t* Sp, const int32_t* Si, const double* Sx,
                    const int32_t* Qp, const int32_t* Qi, const double* Qx, const double* h_max,
                    int order, HostPlan& out, std::string& err);

// sequential greedy pass of sdp_solver.rounding_one_attempt (sdp_solver.py:70-101)
int round_greedy_host(int64_t n, int Z, const int32_t* Sp, const int32_t* Si, const double* Sx,
                      const int32_t* Qp, const int32_t* Qi, const double* Qx, const double* h_max,
                      const int32_t* rank, const int32_t* pref, int32_t* z_vec, int64_t* remainder,
                      std::string& err);

}  // namespace sigsdp
