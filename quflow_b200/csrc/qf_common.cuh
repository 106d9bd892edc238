// Shared declarations for the quflow_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/quflow_b200.h"

// ---------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------
void qf_set_error(const char *fmt, ...);

#define QF_CUDA(call)                                                                             \
    do {                                                                                          \
        cudaError_t _e = (call);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            qf_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return QF_ERR_CUDA;                                                                   \
        }                                                                                         \
    } while (0)

// Entry points run on the handle's device and leave the caller's current device as they found it.
struct QfDeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit QfDeviceGuard(int dev)
    {
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess) cur = -1;
        if (cur != dev) {
            ok = (cudaSetDevice(dev) == cudaSuccess);
            prev = cur;
        }
    }
    ~QfDeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
#define QF_ON_DEVICE(dev)                                                                         \
    QfDeviceGuard _qf_guard(dev);                                                                 \
    if (!_qf_guard.ok) {                                                                          \
        qf_set_error("cudaSetDevice(%d) failed at %s:%d", (int)(dev), __FILE__, __LINE__);        \
        return QF_ERR_CUDA;                                                                       \
    }

#define QF_CHECK(expr)                  \
    do {                                \
        int _s = (expr);                \
        if (_s != QF_OK) return _s;     \
    } while (0)

// ---------------------------------------------------------------------------------------
// device-resident control block of one ensemble member (one per batch entry)
// ---------------------------------------------------------------------------------------
struct QfCtrl {
    double tol;           // tolerance in use
    double resnorm;       // residual of the last evaluated iteration (inf at step start)
    double resnorm_old;
    double norm0;         // ||W||_inf at call start (tol = factor * norm0)
    long long total_it;
    long long n_maxit;
    unsigned long long gseq;          // executed iterations over the life of the handle (never reset)
    unsigned long long xseq;          // exchange barriers passed over the life of the handle (tile-exchange path; never reset)
    unsigned long long resmax_bits;   // running max of residual row sums (bit pattern of a non-negative double)
    unsigned int ticket;              // blocks of k_control that have finished
    int active;           // 1 while the fixed-point loop of the current step runs
    int it;               // iterations executed in the current step
    int nonfinite;        // sticky: 1 = residual was NaN/Inf, 2 = a peer did not answer in time -> everything becomes a no-op
    int skew_exact;       // tile exchange: W is bit-for-bit skew-Hermitian (checked at call start), so the lower triangle of
                          // W~ can be rebuilt locally from the upper one and only upper tiles cross NVLink
    int steps_done;
};

// double2 helpers: (x, y) = (re, im)
__device__ __forceinline__ double2 zadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 zsub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 zconj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double2 zscale(double s, double2 a) { return make_double2(s * a.x, s * a.y); }
__device__ __forceinline__ double zabs(double2 a) { return hypot(a.x, a.y); }

// ---------------------------------------------------------------------------------------
// the handle
// ---------------------------------------------------------------------------------------
struct QfGemmPlan;   // zgemm.cu
#define QF_MAX_RANKS 16

struct qf_handle_s {
    int N = 0;
    int batch = 1;
    int device = 0;
    int sm_count = 148;
    size_t mat_elems = 0;      // N*N
    // unit-packed LDL^T factors of the Hoppe-Yau tridiagonal systems: w_k = o_k / u_{k-1}, 1/u_k (poisson.cu)
    double *ptab_w = nullptr, *ptab_iu = nullptr;
    int *ptab_units = nullptr;      // [nunits][8] = bL, posbase, bS, PS, nlink, 0, 0, 0 (poisson.cu: qf_build_tables)
    int p_L = 0;                    // positions per thread (0: band kernel not available for this N)
    int p_M = 8;                    // diagonals per band
    int p_CL = 1;                   // CTAs per cluster
    int p_NT = 0;                   // threads per CTA
    int p_NTMAX = 256;              // launch bound of the instantiation in use
    int p_nunits = 0;               // CTAs in the launch
    int p_pf = 1;                   // L2 prefetch of the following unit (QF_POISSON_PF, read at handle creation)
    // work matrices, batch * N * N complex128 each
    double2 *dW = nullptr, *Wh = nullptr, *P = nullptr, *A = nullptr, *S = nullptr, *scratch = nullptr;
    double2 *Wst = nullptr;       // tile-exchange path: the state W during a call (IPC-visible; W_dev is copied in and out)
    void *arena = nullptr;        // tile-exchange path: ONE allocation [W~ | A | Wst | rowpart2 | flags] exported to the peers
    double2 *kahan_c = nullptr;   // compensation term (compsum), lazily allocated
    double2 *io = nullptr;        // staging for the *_host entry points, lazily allocated
    double2 *io2 = nullptr;
    // residual partial row sums: [batch][2][nslots][N]
    double *rowpart = nullptr;
    // the same for the fused GEMM-2 epilogue: direct [batch][N][nsd] then mirrored [batch][N][nsm] (QfEpiPost)
    double *rowpart2 = nullptr;
    int nsd = 0, nsm = 0;
    int fuse_post = 1;            // QF_FUSE_POST=0: separate k_post launch after a plain second GEMM
    double *inner_part = nullptr; // qf_inner partial sums, lazily allocated
    int nslots = 0;
    QfCtrl *ctrl = nullptr;       // [batch] device
    int32_t *iters_dev = nullptr; // [batch * steps_cap]
    int steps_cap = 0;
    QfCtrl *ctrl_host = nullptr;  // pinned
    long long launches = 0;
    // multi-GPU
    void *nccl_comm = nullptr;
    void *p2p = nullptr;          // QfP2P (comm.cu)
    int comm_mode = 0;            // 0: none / emulated, 1: NCCL all-gather of A and S (eager only), 2: peer-memory pull
                                  // kernels for A and S, 5: tile exchange (default: sharded fused tail, W~ pushed by owners)
    int xchg_ce = 0;              // tile exchange: W~ tiles travel by copy engines (memcpy nodes) instead of a copy kernel
    int skew_host = -1;           // host copy of QfCtrl.skew_exact of the current call (copy-engine mode only)
    int rank = 0, nranks = 1;     // nranks > 1 with nccl_comm == nullptr: all ranks emulated on this GPU (tests)
    QfGemmPlan *gemm = nullptr;
    // CUDA-graph execution of a step (isomp.cu)
    void *step_graph = nullptr;
    cudaStream_t cap_stream = nullptr;
    cudaStream_t side_stream = nullptr;          // forked branch of the iteration (tile exchange: local mirror of W~)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int use_graph = 1;
    unsigned long long cap_cond = 0;   // conditional handle of the step graph being captured
    int cap_use_cond = 0;              // k_control arms the WHILE node (set only while capturing)
    // host-stepped driver (qf_step_*): parameters fixed by qf_step_open
    double step_eps = 0.0;
    unsigned step_flags = 0;
    int step_open = 0;
    int multistate = 0;           // the members are one (k, N, N) run: member 0's stream function drives everybody
    int graph_warned = 0;
};

// ---------------------------------------------------------------------------------------
// kernels' host launchers (each returns a qf_status)
// ---------------------------------------------------------------------------------------
// poisson.cu
int qf_build_tables(qf_handle_s *h);
// Wh = W (+ dW);  P = eps * Delta^{-1} Wh.  If ctrl != null the launch is skipped on device when !ctrl->active.
int qf_launch_poisson(qf_handle_s *h, const double2 *W, const double2 *dW, double2 *Wh, double2 *P, double eps,
                      bool gated, cudaStream_t st, int members = -1 /* solve only the first `members` (default: all) */);
int qf_launch_laplace(qf_handle_s *h, const double2 *P, double2 *W, cudaStream_t st);
int qf_launch_whalf(qf_handle_s *h, const double2 *W, const double2 *dW, double2 *Wh, cudaStream_t st);

// zgemm.cu
int qf_gemm_create(qf_handle_s *h);
void qf_gemm_destroy(qf_handle_s *h);
// ---------------------------------------------------------------------------------------
// Tile-exchange data path of the row-sharded multi-GPU run (comm_mode 5, comm.cu).
// The N rows are cut into 2G blocks of hb rows; rank r owns blocks r and 2G-1-r (balances the upper-triangular second
// GEMM).  A tile pair {(I,J), (J,I)}, I <= J, belongs to the owner of row block I.  Device-visible description, passed
// by value to the kernels that store into peer memory; every table has QF_MAX_RANKS entries, the own rank included.
// ---------------------------------------------------------------------------------------
#define QF_XF_G1 0      // flag kind: "my GEMM-1 tiles of iteration seq have landed at their consumers"
#define QF_XF_X 1       // flag kind: "everything I pushed for exchange #seq has landed" (W~ tiles, residual partials, state)
struct QfXchg {
    int nranks = 1, rank = 0, hb = 0;
    double2 *const *peerWh = nullptr;                 // W~ (replicated: every owner pushes its tiles to everybody)
    double2 *const *peerA = nullptr;                  // A = P~ W~ (own rows + the transposed tiles the fused tail needs)
    double2 *const *peerWst = nullptr;                // state W (sharded by tile pairs during a call, completed at its end)
    double *const *peerPart = nullptr;                // residual partial sums of every rank: [rowpart (k_post) | rowpart2 (fused tail)]
    long long part2_off = 0;                          // offset (doubles) of rowpart2 inside that allocation
    unsigned long long *const *peerFlags = nullptr;   // [2][QF_MAX_RANKS] per rank, written by the peers
    unsigned long long *myFlags = nullptr;
    long long timeout_cycles = 60000000000ll;         // give up on a silent peer after this many SM cycles (QF_COMM_TIMEOUT_S)
    int upper_only = 0;                               // W~ exchange: send upper tiles only when W is exactly skew-Hermitian
    int push_inline = 0;                              // W~ tiles are stored to the peers by the tail / update kernels themselves
                                                      // (overlaps the NVLink stores with their memory latency) instead of k_xchg_push_wh
};
__host__ __device__ __forceinline__ int qf_owner_of_row(int row, int hb, int G)
{
    const int blk = row / hb;
    return blk < G ? blk : 2 * G - 1 - blk;
}

#ifdef __CUDACC__
// Spin (one thread) until every peer's flag of `kind` in this rank's flag array has reached seq.  Bounded: a peer that
// never answers (a failed rank, mismatched launch sequences) must not hang the GPU inside a kernel; after
// x.timeout_cycles (30 s by default, QF_COMM_TIMEOUT_S) the wait gives up and returns false, and the caller marks the
// run as failed (QfCtrl.nonfinite = 2); once a run is marked failed the later waits return at once.
__device__ __forceinline__ bool xchg_wait_flags(const QfXchg &x, int kind, unsigned long long seq)
{
    const volatile unsigned long long *f = x.myFlags + kind * QF_MAX_RANKS;
    const long long t0 = clock64();
    bool ok = true;
    for (int p = 0; p < x.nranks && ok; ++p) {
        if (p == x.rank) continue;
        while (f[p] < seq) {
            __nanosleep(100);
            if (clock64() - t0 > x.timeout_cycles) { ok = false; break; }
        }
    }
    __threadfence_system();      // acquire: what the peers stored before raising their flags is visible from here on
    return ok;
}
#endif

// C = A * B.  upper_only: compute only the 64-wide column blocks that intersect the upper triangle
// (used for S = A P~ which is skew-Hermitian).  rank/nranks: row-block sharding (rank < 0: all blocks).
// a_permuted: the A operand is itself stored in the rank-permuted row layout (an earlier GEMM's output; legacy
// all-gather data paths, comm_mode 1 and 2).  natural: the output keeps its natural row positions (tile-exchange path);
// with xg (nranks > 1) every finished tile strictly below the diagonal is also stored into the A buffer of the rank
// that owns the tile's column block: that rank needs it, transposed, in the fused tail of its second GEMM.
int qf_launch_zgemm(qf_handle_s *h, const double2 *A, const double2 *B, double2 *C, bool upper_only, bool gated,
                    int rank, int nranks, bool a_permuted, cudaStream_t st, bool natural = false, const QfXchg *xg = nullptr);

// Fused tail of the fixed-point iteration (isospectral.py:499-536), executed in the epilogue of the second GEMM by the
// CTA that finishes an upper tile (I <= J) of S = A P~, straight from the accumulators (S never goes to memory):
//     c = A_ij - conj(A_ji)           conj_subtract_, isospectral.py:66-81
//     d = S_ij + c                    the new iterate dW_ij (:499,:509);  dW_ji = -conj(d)
//     res_ij = |dW_old_ij - d|        residual (:526,:534): partial row sums for the infinity norm
//     W~_ij = W_ij + d,  W~_ji = W_ji - conj(d)     the next midpoint state (:481-482)
// Residual partial sums are written to fixed slots, one per (row, 16-column group) for the direct elements and one per
// (column, 32-row group) for the mirrored ones, so the row sums k_control forms do not depend on scheduling.
struct QfEpiPost {
    const double2 *A;        // A = P~ W~ of this iteration, natural row layout
    double2 *dW;             // in: previous iterate; out: new iterate (both triangles)
    const double2 *W;        // state at the start of the step
    double2 *Wh;             // out: W + dW (both triangles)
    double *part_direct;     // [batch][N][nsd]
    double *part_mirror;     // [batch][N][nsm]
    int nsd, nsm;
};
// S = A P~ restricted to the upper tiles with the fused tail above; A is in natural row layout.
bool qf_gemm_can_fuse_post(qf_handle_s *h);
int qf_launch_zgemm_post(qf_handle_s *h, const double2 *A, const double2 *B, const QfEpiPost &epi, bool gated,
                         int rank, int nranks, cudaStream_t st, const QfXchg *xg = nullptr);

// ---------------------------------------------------------------------------------------
// Row-block sharding across G ranks (multi-GPU).
// The N rows are cut into 2G blocks of hb rows; rank r owns logical blocks r and 2G-1-r, which balances the
// upper-triangular second GEMM.  GEMM outputs (A, S) are stored with rows PERMUTED so that each rank's two
// blocks are contiguous: permuted row = (2r + slot) * hb + (i mod hb).  One in-place ncclAllGather then
// completes the matrix on every rank.  G = 1 is the identity.
// ---------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int qf_block_rows(int N, int G) { return G == 1 ? N : (N + 2 * G - 1) / (2 * G); }
__host__ __device__ __forceinline__ int qf_prow(int i, int hb, int G)
{
    if (G == 1) return i;
    const int blk = i / hb;
    const int r = blk < G ? blk : 2 * G - 1 - blk;
    const int slot = blk < G ? 0 : 1;
    return (2 * r + slot) * hb + (i - blk * hb);
}
int qf_comm_allgather_rows(qf_handle_s *h, double2 *M, cudaStream_t st);   // comm.cu
int qf_comm_p2p_allgather(qf_handle_s *h, int kind, bool gated, cudaStream_t st);   // comm.cu
void qf_p2p_destroy(qf_handle_s *h);
// tile-exchange path (comm.cu)
const QfXchg *qf_xchg_desc(qf_handle_s *h);                                          // nullptr unless comm_mode == 5
int qf_xchg_signal(qf_handle_s *h, int kind, bool gated, cudaStream_t st);           // raise my flag of `kind` at every peer
int qf_xchg_wait(qf_handle_s *h, int kind, bool gated, cudaStream_t st);             // until every peer raised its flag here
int qf_xchg_barrier(qf_handle_s *h, int kind, bool gated, cudaStream_t st);          // both in one launch (not for the lock-step emulation)
int qf_xchg_push_wh(qf_handle_s *h, bool gated, cudaStream_t st);                    // my tiles of W~ -> every peer
int qf_xchg_mirror_wh(qf_handle_s *h, bool gated, cudaStream_t st);                  // lower triangle of W~ from the upper one
int qf_xchg_skew_check(qf_handle_s *h, const double2 *W, cudaStream_t st);           // -> ctrl[0].skew_exact
int qf_xchg_push_state(qf_handle_s *h, cudaStream_t st);                             // my tile pairs of Wst -> every peer
int qf_xchg_push_rows(qf_handle_s *h, cudaStream_t st);                              // my row blocks of Wst -> every peer
// isomp.cu: qf_isomp with W_dev == NULL allowed on the tile-exchange path (state already staged in h->Wst)
int qf_isomp_impl(qf_handle_s *h, void *W_dev, double dt, int steps, double tol, int maxit, int minit, unsigned flags,
                  qf_stats *stats, int32_t *iters_per_step, cudaStream_t st);
int qf_gemm_prepare(qf_handle_s *h, int rank, int nranks);
int qf_gemm_prepare_gather(qf_handle_s *h);                                 // zgemm.cu: tile lists of the all-gather path                  // zgemm.cu: build tile lists (allocates)
void qf_graph_destroy(qf_handle_s *h);                                      // isomp.cu

// isomp.cu
int qf_launch_norm_inf(qf_handle_s *h, const double2 *W, cudaStream_t st);   // -> ctrl[b].norm0
