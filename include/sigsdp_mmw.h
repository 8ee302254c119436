/*
 * sigsdp_mmw.h -- C ABI of the B200-native MMW SDP hot path (libsigsdp_mmw.so).
 *
 * The reference (zhouyou-gu/sig-sdp-mmw) is pure Python and has no FFI layer; the
 * drop-in boundary is the duck-typed solver object handed to
 * sim_src/alg/binary_search_relaxation.py:10,50,53 (class mmw,
 * sim_src/alg/mmw.py:12-229, and sdp_solver.rounding, sim_src/alg/sdp_solver.py:18-107).
 * This header is what a ctypes binding inside that class calls instead of
 * numpy/scipy; every entry point cites the reference lines it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, a negative SIGSDP_E* code on failure;
 *     sigsdp_last_error() returns a thread-local message for the last failure.
 *   - `_host` pointers are host memory, `_dev` pointers are device memory on the
 *     plan's device.  Handles own their device workspace (cudaMalloc).
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).
 *     No entry point synchronises the device unless its comment says so.
 *   - indices are int32, the dual / loss / averaging state is always fp64; only the
 *     sketch block (Omega, Taylor terms, Y_h) follows the solver's dtype.
 *   - no torch types, no C++ types, no exceptions cross this boundary.
 */
#ifndef SIGSDP_MMW_H
#define SIGSDP_MMW_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIGSDP_OK 0
#define SIGSDP_EINVAL (-1)   /* bad argument / malformed CSR / asymmetric Q_asso */
#define SIGSDP_ECUDA (-2)    /* CUDA runtime error (message has the cudaError string) */
#define SIGSDP_ENOMEM (-3)
#define SIGSDP_ESTATE (-4)   /* call order violated (e.g. fetch before iterate) */

#define SIGSDP_F64 0
#define SIGSDP_F32 1

/* launch modes of sigsdp_solver_iterate */
#define SIGSDP_MODE_FUSED 0     /* one persistent cooperative kernel per call (default) */
#define SIGSDP_MODE_STEPWISE 1  /* one kernel per phase / Taylor term (profiling, debugging) */

typedef struct sigsdp_plan sigsdp_plan;     /* graph plan: Z-independent (reused across the binary search) */
typedef struct sigsdp_solver sigsdp_solver; /* MMW state for one (plan, Z, D, eta, dtype) */
typedef struct sigsdp_batch sigsdp_batch;   /* many independent small instances, one thread block each */

const char* sigsdp_last_error(void);
int sigsdp_version(void);
/* number of CUDA devices visible, or a negative error */
int sigsdp_device_count(void);
/* The next `count` numbers of numpy's legacy normal stream (np.random.randn / RandomState.standard_normal: MT19937 ->
 * 53-bit doubles -> polar method), bit for bit, produced on all host cores (the Mersenne Twister runs sequentially, the
 * candidate pairs -- four output words each, accepted or not -- are evaluated in parallel and compacted in order).
 * The reference draws np.random.randn(K, D) once per iteration (mmw.py:226); an unchanged driver sees the same numbers
 * 3-4x sooner.  key624 / pos / has_gauss / cached_gauss: np.random.get_state() on entry, the state numpy would be left
 * in on exit (hand it to np.random.set_state). */
int sigsdp_numpy_standard_normal(uint32_t* key624, int32_t* pos, int32_t* has_gauss, double* cached_gauss, int64_t count,
                                 double* out_host);
/* Position-weighted 64-bit checksum of a host buffer, computed on the builder's host threads.  The host side keys its
 * plan cache with it (the reference re-runs _process_state on every call, mmw.py:26-41; a caller of this library keeps
 * the plan while the state's buffers are unchanged).  bytes may be 0. */
int sigsdp_checksum(const void* data_host, int64_t bytes, uint64_t* out);

/* ------------------------------------------------------------------ plan ----
 * Replaces mmw._process_state (mmw.py:26-41) and the edge-list set-up
 * (mmw.py:52-57): T = S_gain^T with association pairs and the diagonal zeroed,
 * S_sum = T 1, sqrt((T o T) 1), gain-UT / asso-UT edge lists in CSR row-major
 * order, and the symmetric union pattern (diag + gain + asso) every kernel walks.
 * Inputs are the reference's `state` tuple (env.py:168-196): two n x n CSR
 * matrices with sorted, duplicate-free int32 indices and h_max (n).  Built on the
 * host (native C++), uploaded to `device` (device < 0: host-only plan, for inspecting
 * the edge lists without a GPU; no solver can be created on it).  Synchronous.
 * `order`: 0 = keep the caller's node numbering inside the kernels,
 *          1 = renumber nodes internally for locality (clustered BFS); all inputs
 *              and outputs of this API stay in the caller's numbering.
 */
int sigsdp_plan_create(int64_t n,
                       const int32_t* S_indptr_host, const int32_t* S_indices_host, const double* S_data_host,
                       const int32_t* Q_indptr_host, const int32_t* Q_indices_host, const double* Q_data_host,
                       const double* h_max_host, int device, int order, sigsdp_plan** out);
void sigsdp_plan_destroy(sigsdp_plan* plan);
/* A built plan as a flat byte image, and a plan created from such an image plus the SAME state: in a multi-GPU launch
 * (one process per GPU on one box) ONE process builds the plan on all host cores and ships the image (a broadcast by
 * the host language) instead of every process building the same plan side by side on a share of the cores.  The
 * image only skips the host build; the inputs are still needed (host copies for the rounding entry points).  An
 * image that does not describe an n-node plan is rejected (SIGSDP_EINVAL). */
/* 1 when sigsdp_plan_create would build an n-node plan on the device (large graph, device >= 0, not overridden by
 * SIGSDP_PLAN_BUILDER=host), 0 when on the host cores: the host side of a multi-GPU launch lets every rank build its
 * own plan in the first case and broadcasts rank 0's image in the second. */
int sigsdp_plan_builds_on_device(int64_t n, int device);
int sigsdp_plan_image_size(const sigsdp_plan* plan, int64_t* bytes);
int sigsdp_plan_image(const sigsdp_plan* plan, void* image_host);
int sigsdp_plan_create_from_image(int64_t n,
                                  const int32_t* S_indptr_host, const int32_t* S_indices_host, const double* S_data_host,
                                  const int32_t* Q_indptr_host, const int32_t* Q_indices_host, const double* Q_data_host,
                                  const double* h_max_host, const void* image_host, int64_t image_bytes, int device,
                                  sigsdp_plan** out);

/* info[0..7] = n, E_gain, E_asso, nnzL (= n + 2 E), nnz(T), device, order, max row length */
int sigsdp_plan_info(const sigsdp_plan* plan, int64_t info[8]);
/* Edge lists in the reference's order (mmw.py:56-57): gain-UT (E_gain) then asso-UT
 * (E_asso); any pointer may be NULL.  t_ij = T[i,j], t_ji = T[j,i] on gain edges. */
int sigsdp_plan_edges(const sigsdp_plan* plan, int32_t* gain_i_host, int32_t* gain_j_host,
                      double* t_ij_host, double* t_ji_host, int32_t* asso_i_host, int32_t* asso_j_host);
/* S_sum (n) and sqrt of the row sums of T o T (n), mmw.py:34-39 */
int sigsdp_plan_vectors(const sigsdp_plan* plan, double* S_sum_host, double* t_rownorm_host);
/* the internal node numbering: perm[new] = old (identity when order = 0) */
int sigsdp_plan_perm(const sigsdp_plan* plan, int32_t* perm_host);

/* ---------------------------------------------------------------- solver ----
 * Replaces the state set-up of mmw._run (mmw.py:59-73): C = E_asso + 2K,
 * Y = 1/C, e_accu = 0, L_accu = 0, X = I, X_avgd = 0, Y_avgd = 0, and norm_H for
 * this Z (mmw.py:39).  D = Z * rank_radio (mmw.py:180).
 */
int sigsdp_solver_create(const sigsdp_plan* plan, int Z, int D, double eta, int dtype, sigsdp_solver** out);
/* Same, choosing the kernels' row tiling: -1 automatic (the largest tile of consecutive
 * rows whose distinct neighbour rows of the sketch block fit in shared memory; the Taylor
 * SpMM and the Gram then stage that slice with TMA bulk copies), 0 = direct-gather
 * kernels, > 0 = that many rows per tile. */
int sigsdp_solver_create_tiled(const sigsdp_plan* plan, int Z, int D, double eta, int dtype, int tiling,
                               sigsdp_solver** out);

/* ------------------------------------------------- sketch-column sharding ----
 * One graph across several GPUs: the columns of the D_total-wide sketch are split, rank r
 * owning [col0, col0 + D).  The columns of exp(L/2) Omega are independent, so the Taylor
 * terms need NO exchange; the dual / loss state is replicated (every rank computes the same
 * bits); the only exchange is one all-reduce (sum) per iteration of the un-normalised Gram
 * partials and the partial ||y_k||^2 -- sigsdp_solver_exchange_buffer(), nnzL + n doubles.
 * Protocol per iteration on every rank:
 *     sigsdp_solver_split_step(s, 1, ...)   finishes the previous iteration's Gram from the
 *                                           reduced buffer (if any), runs dual..Taylor terms,
 *                                           writes this rank's partials to the buffer
 *     all-reduce(sum) of the buffer         (NCCL, by the host language)
 * and, before reading any state, sigsdp_solver_split_step(s, 0, ...) to finish the last Gram.
 * col0 and D must be multiples of 2 (fp64) / 4 (fp32).  Omega (injected or Philox) is the
 * same D_total-wide matrix on every rank; a rank reads / generates all of a row only to
 * normalise it. */
int sigsdp_solver_create_sharded(const sigsdp_plan* plan, int Z, int D_total, int col0, int D, double eta, int dtype,
                                 int tiling, sigsdp_solver** out);
int sigsdp_solver_split_step(sigsdp_solver* s, int do_iter, const double* omega_dev, uint64_t seed, void* stream);
int sigsdp_solver_exchange_buffer(sigsdp_solver* s, void** dev_ptr, int64_t* count);

/* ------------------------------------------------------------ row sharding ----
 * ONE graph across the GPUs of a box, one process (rank) per GPU -- BASELINE configs[3].
 * Rank r of `nranks` (<= 8) owns a contiguous range of rows of the locality-ordered pattern
 * (whole row tiles, balanced by non-zeros) and with them their part of everything:
 * L_accu / X / X_avgd entries, dual weights (own rows' D and H constraints, the association
 * edges whose smaller endpoint it owns), sketch rows.  Per Taylor term (scipy
 * _expm_multiply.py:291-303) the only rows of the sketch block another rank reads are those
 * of its boundary ("halo"): the SpMM epilogue stores them straight into the neighbour's copy
 * of the block through peer-mapped memory (NVLink), the row sums r (mmw.py:133) and the
 * loss weights q (mmw.py:160-163) travel the same way once per iteration, and the grid
 * barrier that separates the phases is extended across the GPUs: it carries the packed
 * scalars every rank must agree on (max e_accu and the soft-max sums mmw.py:139, ||A||_1
 * and the ||.||_inf of the Taylor terms _expm_multiply.py:259-303, the trace mmw.py:183),
 * reduced in rank order so all ranks continue with identical bits.  No host involvement and
 * no NCCL call inside sigsdp_solver_iterate.
 *
 * Set-up on every rank (same plan arguments, same Z / D / eta / dtype / tiling):
 *   sigsdp_solver_create_rows(plan, ..., rank, nranks, max_blocks, &s)
 *   sigsdp_solver_shard_ipc_handle(s, h64)       64-byte CUDA IPC handle of the exchange arena
 *   (exchange the handles between the processes: e.g. torch.distributed.all_gather)
 *   sigsdp_solver_shard_attach_ipc(s, handles)   nranks x 64 bytes, in rank order
 * or, for several shards driven by ONE process (tests; shards on one GPU or on P2P-capable
 * GPUs): sigsdp_solver_shard_attach_local(all the shards in rank order).
 * Then every rank calls sigsdp_solver_iterate with the same arguments; the kernels of the
 * ranks synchronise among themselves (a barrier that does not complete within
 * SIGSDP_SHARD_TIMEOUT_S seconds, default 20, traps -- e.g. a rank that never launched).
 * sigsdp_solver_reset on a shard must be followed by a barrier between the ranks (host
 * side) before any of them iterates again.  max_blocks > 0 caps the grid (shards sharing one
 * GPU must be co-resident: use blocks-per-GPU / nranks); 0 = fill the GPU.
 * The sigsdp_solver_get_* fetches of a shard return the entries the rank owns and zeros
 * elsewhere, so the sum over the ranks is the whole vector.
 * info[0..11] = rank, nranks, first row, end row, first tile, end tile, halo rows pushed per
 * Taylor term (x 2 blocks x Dp x word bytes = bytes sent), halo rows read, arena bytes,
 * incident association edges, owned association edges, attached. */
int sigsdp_solver_create_rows(const sigsdp_plan* plan, int Z, int D, double eta, int dtype, int tiling, int rank,
                              int nranks, int max_blocks, sigsdp_solver** out);
int sigsdp_solver_shard_info(const sigsdp_solver* s, int64_t info[12]);
/* The row partition a row-sharded solver of `nranks` ranks would use, without a device (host-only plans work): rows
 * cut at tile boundaries for the given tile caps (max_rows > 0, as in sigsdp_plan_tile_stats) or anywhere
 * (max_rows = 0).  row0_out[nranks + 1] = first row of every rank (internal numbering); optional per rank: rows pushed
 * per Taylor term (row x destination pairs), distinct foreign rows read, association edges owned. */
int sigsdp_plan_row_partition(sigsdp_plan* plan, int nranks, int max_rows, int ucap, int nnzcap, int64_t* row0_out,
                              int64_t* halo_send_out, int64_t* halo_recv_out, int64_t* owned_asso_out);
int sigsdp_solver_shard_arena(sigsdp_solver* s, void** dev_ptr, int64_t* bytes);
int sigsdp_solver_shard_ipc_handle(sigsdp_solver* s, void* handle64_host);
int sigsdp_solver_shard_attach_ipc(sigsdp_solver* s, const void* handles64_host);
int sigsdp_solver_shard_attach_local(sigsdp_solver* const* ranks, int count);
void sigsdp_solver_destroy(sigsdp_solver* s);
int sigsdp_solver_reset(sigsdp_solver* s, void* stream);
int sigsdp_solver_set_mode(sigsdp_solver* s, int mode);
/* info[0..11] = n, Z, D, Dp (padded row length), C, iterations done, dtype, grid blocks,
 * threads per block, lanes per row, rows per tile (0 = gather kernels), dynamic shared
 * memory bytes per block */
int sigsdp_solver_info(const sigsdp_solver* s, int64_t info[12]);

/* n_iters MMW iterations, each = mmw.py:77-78 (averaging) + 124-142 (dual) +
 * 144-170 (loss) + 172-197 (sketch: expm_half_randsk 224-229 with scipy's
 * expm_multiply restated on device, and the edge-only Gram).
 *   omega_dev != NULL : device pointer to n_iters blocks of raw standard normals,
 *                       fp64, row-major n x D in the caller's node numbering, block i
 *                       at omega_dev + i * n * D (what np.random.randn returned in
 *                       the reference: parity mode);
 *   omega_dev == NULL : normals are generated on device (Philox4x32-10 keyed by
 *                       `seed`, counter = (iteration, row, column)).
 * Asynchronous on `stream` in fused mode; stepwise mode synchronises the stream
 * after every kernel (it reads the Taylor controller back).
 */
int sigsdp_solver_iterate(sigsdp_solver* s, int n_iters, const double* omega_dev, uint64_t seed, void* stream);

/* Synchronising host fetches (caller's node numbering, fp64):
 *   Y        (C)   dual weights [D | F | H]                    mmw.py:139
 *   e_accu   (C)                                               mmw.py:137
 *   Y_avgd   (C)   running sum of Y                            mmw.py:78
 *   X, X_avgd      diag (n), gain edges (E_gain), asso edges (E_asso); X_avgd is the
 *                  running SUM (the reference divides by nit at mmw.py:203)
 *   L_accu         diag (n), gain edges, asso edges            mmw.py:167
 *   Y_h      (n x D) exp(L_accu/2) Omega of the last iteration mmw.py:180
 * Any pointer may be NULL. */
int sigsdp_solver_get_dual(sigsdp_solver* s, double* Y_host, double* e_accu_host, double* Y_avgd_host);
int sigsdp_solver_get_X(sigsdp_solver* s, int averaged, double* diag_host, double* gain_host, double* asso_host);
int sigsdp_solver_get_L(sigsdp_solver* s, double* diag_host, double* gain_host, double* asso_host);
int sigsdp_solver_get_sketch(sigsdp_solver* s, double* Yh_host);
/* Inverse of sigsdp_solver_get_X: load X (or the running sum X_avgd) from edge-list form,
 * every entry -- e.g. the sum of the row shards' fetches, before the final factor. */
int sigsdp_solver_set_X(sigsdp_solver* s, int averaged, const double* diag_host, const double* gain_host,
                        const double* asso_host);
/* Warm start across the probes of the binary search (binary_search_relaxation.py:44-71 solves the same state for a
 * sequence of Z): the constraint vector [D | F | H] has the same shape for every Z, so the accumulated constraint
 * losses e_accu, the dual weights Y and the soft-max shift of `src` (a solver of the SAME plan, any Z / D / dtype,
 * row shards excluded) become the initial dual state of `dst` instead of e_accu = 0, Y = 1/C (mmw.py:60-62).  `dst`
 * must not have iterated since its creation / reset; L_accu, X and the running sums start as usual.  Device-to-device,
 * asynchronous on `stream`.  Not something the reference does: off by default in the drop-in object. */
int sigsdp_solver_warm_start(sigsdp_solver* dst, const sigsdp_solver* src, void* stream);
/* Per-iteration Taylor controller history of the last `count` iterations
 * (scipy _expm_multiply.py:259-303, _fragment_3_1 :503-558): m_star, s, executed
 * terms (int32 each) and ||A - mu I||_1, mu (fp64 each).  Any pointer may be NULL. */
int sigsdp_solver_get_history(sigsdp_solver* s, int count, int32_t* m_star_host, int32_t* s_host,
                              int32_t* nterms_host, double* a1norm_host, double* mu_host);
/* Device-timed microseconds of the last `count` iterations, count x 4 row-major:
 * dual (mmw.py:124-142) | loss (:144-170) | Taylor terms of the sketch (:180, :224-229) |
 * edge Gram (:182-194) -- the reference logs the first two as mmw_dual / mmw_loss and the sum
 * of the last two as mmw_expm.  Fused mode only (zeros otherwise). */
int sigsdp_solver_get_phase_times(sigsdp_solver* s, int count, double* us_host);
/* Diagnostics since create/reset: out8[4] = nanoseconds the fused kernel's leader thread spent
 * in team barriers (grid barriers; for a row shard this includes the cross-GPU wait).  Row shards
 * split that: out8[5] = waiting for this GPU's other blocks, out8[6] = reducing and sending the
 * packed scalars and epoch flags to the peers, out8[7] = waiting for the peers' flags.  The other
 * entries are reserved (0). */
int sigsdp_solver_debug_cycles(sigsdp_solver* s, int64_t out8[8]);
/* total Taylor terms (SpMM passes) executed since create/reset */
int sigsdp_solver_total_terms(sigsdp_solver* s, int64_t* out);
/* Device pointer and length (doubles) of one of the solver's fp64 state arrays in its INTERNAL
 * layout, so that the host language can run a collective on it in place (NCCL through
 * torch.distributed) without a trip through host memory:
 *   SIGSDP_ARR_X_AVGD / SIGSDP_ARR_X : nnzL values in the pattern's CSR order (sigsdp_plan_pattern)
 *   SIGSDP_ARR_Y_AVGD / SIGSDP_ARR_Y : C values [D | F | H], D and H in the internal node numbering
 * A row shard only ever writes the X_AVGD / Y_AVGD entries it owns (the rest stay zero), so an
 * all-reduce (sum) over the ranks leaves the complete running sums on every rank -- what the
 * final factor (mmw.py:202-216) and the gap log need. */
#define SIGSDP_ARR_X_AVGD 1
#define SIGSDP_ARR_X 2
#define SIGSDP_ARR_Y_AVGD 3
#define SIGSDP_ARR_Y 4
int sigsdp_solver_device_array(sigsdp_solver* s, int which, void** dev_ptr, int64_t* count);

/* Row-tile statistics of a plan for given caps (host only, no device needed): tiles, bulk-copy
 * runs, staged rows (incl. gap rows), max staged rows / non-zeros of a tile, nnz. */
int sigsdp_plan_tile_stats(sigsdp_plan* p, int max_rows, int ucap, int nnzcap, int64_t out6[6]);

/* Standard normals of the throughput-mode generator (Philox4x32-10 + Box-Muller),
 * n x D row-major, for testing its moments.  Synchronising. */
int sigsdp_debug_normals(uint64_t seed, int64_t iter, int n, int D, int dtype, double* out_host);

/* ------------------------------------------------------- eigen building blocks ---
 * The Lanczos solvers that replace ARPACK (eigsh mmw.py:115, svds mmw.py:215) run in the
 * host language over these: a symmetric matrix M on the plan's pattern is materialised in
 * the solver's scratch, then y = M x is applied to device vectors in the INTERNAL node
 * numbering (sigsdp_plan_perm), nvec columns, column-major, leading dimension n.
 *   xavg_matrix : M = scale * X_avgd (running sum)                       mmw.py:203
 *   gap_prepare : the running means at the start of the next iteration,
 *                 X~ = (X_avgd + X)/N, Y~ = (Y_avgd + Y)/N, N = iterations done + 1;
 *                 *e_max_host = max_c e_c(X~) (mmw.py:80-96) and M = L(Y~) (mmw.py:98-113).
 *                 Synchronising.
 *   get_matrix  : M's values in the pattern's CSR order (sigsdp_plan_pattern), for tests. */
int sigsdp_solver_xavg_matrix(sigsdp_solver* s, double scale, void* stream);
int sigsdp_solver_gap_prepare(sigsdp_solver* s, double* e_max_host, void* stream);
int sigsdp_solver_symv(sigsdp_solver* s, const double* x_dev, double* y_dev, int nvec, void* stream);
int sigsdp_solver_get_matrix(sigsdp_solver* s, double* vals_host);
/* Lanczos steps j0 .. j1-1 on the prepared matrix M with full re-orthogonalisation (classical
 * Gram-Schmidt twice, fixed-order reductions): w = M q_j; w -= Q_j^T (Q_j w) twice;
 * alpha[j] = <q_j, M q_j>; beta[j] = ||w||; q_{j+1} = w / beta[j].  Q_dev is (m+1) x n
 * row-major in the internal numbering, rows 0..j0 filled by the caller; alpha_dev / beta_dev
 * are device arrays of length m.  Asynchronous on `stream`. */
int sigsdp_solver_lanczos_steps(sigsdp_solver* s, double* Q_dev, int m, int j0, int j1, double* alpha_dev,
                                double* beta_dev, void* stream);
/* Chebyshev filter for the steps above: with degree >= 2 the operator of sigsdp_solver_lanczos_steps becomes
 * p(M) = T_degree((M - c) / e), c = (lo + cut) / 2, e = (cut - lo) / 2 (three-term recurrence, `degree` mat-vecs per
 * step): eigenvalues of M inside [lo, cut] are mapped into [-1, 1], those above cut grow like cosh(degree acosh(.)),
 * monotonically, so the top eigenvectors of M are the top eigenvectors of p(M) with a far better relative gap.  The
 * host side (lanczos.chebyshev_filtered_lanczos) picks lo / cut from a short unfiltered run and checks the returned
 * pairs against M itself.  degree < 2 switches the filter off.  Takes effect from the next lanczos_steps call. */
int sigsdp_solver_lanczos_filter(sigsdp_solver* s, int degree, double lo, double cut);
/* the symmetric union pattern in the internal numbering: rowptr (n+1), col (nnzL) */
int sigsdp_plan_pattern(const sigsdp_plan* plan, int32_t* rowptr_host, int32_t* col_host);

/* ---------------------------------------------------------------- rounding ---
 * sdp_solver.rounding_one_attempt (sdp_solver.py:27-107) split at its data
 * dependence:
 *  (1) sigsdp_round_project (device): inprod = randv gX^T (:56), the per-user slot
 *      preference order argsort(-inprod, axis=0) (:57) and the visit key ||gX_k|| (:52);
 *      gX_dev (n x r, row-major fp64), randv_dev (Z x r, rows already normalised :49),
 *      pref_dev (n x Z int32, user-major), norm_dev (n).
 *  (2) sigsdp_round_greedy (host, native, O(nnz Z)): the sequential feasibility pass
 *      (:70-101) on sparse rows instead of toarray(); `rank` is the visit order
 *      argsort(-norm).  z_vec[k] = slot or -1 when unassigned; returns the number of
 *      unassigned users in *remainder (the caller draws their random slots, :104-105).
 *  (3) sigsdp_round_conflicts (device): rounding.py:56-66 in sparse form: per-user
 *      same-slot interference I_k, #users with I_k > h_max_k, #association pairs
 *      sharing a slot.  counts[0] = #violations, counts[1] = #asso conflicts.
 */
int sigsdp_round_project(const sigsdp_plan* plan, const double* gX_dev, int r, const double* randv_dev, int Z,
                         int32_t* pref_dev, double* norm_dev, void* stream);
int sigsdp_round_greedy(int64_t n, int Z,
                        const int32_t* S_indptr_host, const int32_t* S_indices_host, const double* S_data_host,
                        const int32_t* Q_indptr_host, const int32_t* Q_indices_host, const double* Q_data_host,
                        const double* h_max_host, const int32_t* rank_host, const int32_t* pref_host,
                        int32_t* z_vec_host, int64_t* remainder);
int sigsdp_round_conflicts(const sigsdp_plan* plan, const int32_t* z_dev, double* I_dev_or_null,
                           int64_t counts_host[2], void* stream);
/* (2') The greedy pass of (2) on the device, with the SAME result as the sequential pass (sdp_solver.py:70-101): two
 * users interact only if they are neighbours or share a neighbour in S_gain / Q_asso, so rounds of mutually
 * non-interacting users -- each the lowest-ranked undecided user of its neighbourhood -- decide in parallel with the
 * host pass's arithmetic, in the same order per accumulator.  rank_dev (n): visit order, rank_dev[i] = i-th user
 * (argsort(-norm)); pref_dev: n x Z from sigsdp_round_project; z_dev (n): slot or -1.  Synchronises the stream. */
int sigsdp_round_greedy_device(const sigsdp_plan* plan, int Z, const int32_t* rank_dev, const int32_t* pref_dev,
                               int32_t* z_dev, int64_t* remainder_host, int64_t* rounds_host_or_null, void* stream);

/* ------------------------------------------------------------------ batch ----
 * Monte-Carlo sweeps (sim_script/journal_version/sim_all_bler.py:30-40 run `for seed in
 * range(REPEAT)` sequentially): many independent instances advanced by ONE launch, one
 * thread block per instance running the same phases with block-level barriers, no
 * communication.  The solvers stay owned by the caller (fetch results through the
 * sigsdp_solver_get_* calls); they must share the device, the dtype and the lane width.
 * Omega is always generated on device (instance i uses a stream derived from `seed` and i). */
int sigsdp_batch_create(sigsdp_solver* const* solvers, int count, sigsdp_batch** out);
/* Same with explicit instance ids (default: the position in `solvers`): instance id i draws its
 * Omega from the stream keyed by seed + 0x9E3779B97F4A7C15 * (i + 1), whichever launch it is in. */
int sigsdp_batch_create_ids(sigsdp_solver* const* solvers, const int64_t* ids, int count, sigsdp_batch** out);
void sigsdp_batch_destroy(sigsdp_batch* b);
/* When the batch has fewer instances than the device has resident block slots, every instance is
 * given several co-resident blocks (a team with its own barrier, cooperative launch); the count
 * used by the last sigsdp_batch_iterate is returned by sigsdp_batch_blocks_per_instance (>= 1). */
int sigsdp_batch_iterate(sigsdp_batch* b, int n_iters, uint64_t seed, void* stream);
int sigsdp_batch_blocks_per_instance(const sigsdp_batch* b);

#ifdef __cplusplus
}
#endif
#endif /* SIGSDP_MMW_H */
