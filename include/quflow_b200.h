/*
 * quflow_b200 — C ABI of the B200-native isomp hot path.
 *
 * This header is the drop-in boundary for the time-stepping hot path of
 * klasmodin/quflow (Python).  Every entry point names the reference interface
 * it replaces (paths relative to the upstream tree).  The reference is a pure
 * Python package, so its "FFI" for this path is a ctypes binding; the stub a
 * maintainer would add is shown in INTEGRATION.md and implemented in
 * quflow_b200/_cuda/binding.py.
 *
 * Conventions
 *   - plain C, no C++/torch types; all matrices are N x N complex128,
 *     row-major, interleaved (re, im) — numpy's C-contiguous complex128;
 *   - pointers suffixed _dev are device pointers borrowed from the caller
 *     (the Python side owns them through torch tensors); _host are host
 *     pointers (pageable or pinned);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - every function returns 0 on success or a negative qf_status; nothing
 *     throws across the boundary; qf_last_error() returns the message of the
 *     last failure on the calling thread;
 *   - a handle owns all scratch memory, the precomputed Laplacian factors and
 *     the CUDA graphs; one handle per (N, batch, device); not re-entrant;
 *   - there is no CPU fallback: without a CUDA device qf_create fails.
 */
#ifndef QUFLOW_B200_H
#define QUFLOW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qf_handle_s *qf_handle_t;

typedef enum {
    QF_OK = 0,
    QF_ERR_INVALID = -1,    /* bad argument (maps to AssertionError / ValueError in Python) */
    QF_ERR_CUDA = -2,       /* CUDA runtime / driver failure */
    QF_ERR_NONFINITE = -3,  /* residual became NaN/Inf: scipy.linalg.norm's ValueError, isospectral.py:534 */
    QF_ERR_NCCL = -4,       /* NCCL failure (multi-GPU path) */
    QF_ERR_UNSUPPORTED = -5,
    QF_ERR_COMM = -6        /* multi-GPU peer exchange: a peer rank did not answer within the time limit */
} qf_status;

/* Statistics of one qf_isomp call.
 * Replaces the `stats` dict of isomp_fixedpoint (quflow/integrators/isospectral.py:426-427,
 * 451-452, 538-540, 607-611): tol_auto, iterations = total/steps, number_of_maxit = count/steps. */
typedef struct {
    double tol_used;            /* tolerance actually used (auto or given) */
    double last_resnorm;        /* residual of the last executed iteration */
    int64_t total_iterations;   /* fixed-point iterations over all steps */
    int64_t number_of_maxit;    /* steps that ran maxit iterations without meeting the stopping rule */
    int32_t nonfinite;          /* 1 if a NaN/Inf residual stopped the run */
    int32_t steps_done;         /* steps completed (== steps unless nonfinite) */
} qf_stats;

/* Flags for qf_isomp */
#define QF_FLAG_COMPSUM 1u       /* compensated (Kahan) update, isospectral.py:553-589 */
#define QF_FLAG_REINITIALIZE 2u  /* zero the iterate dW at every step, isospectral.py:471-472 */
#define QF_FLAG_MULTISTATE 4u    /* the `batch` members are ONE multi-state (k, N, N) run of the reference: members
                                    1.. are advected by member 0's stream function (select_first, cpu.py:672-674),
                                    tolerance and residual come from member 0 (isospectral.py:444-446, 528-531) */

#define QF_FLAG_HOST_ROWS_OWN 8u  /* qf_isomp_host on a row-sharded handle (tile-exchange path): the host array of this
                                    rank is read and written on the rank's OWN two row blocks only (a row-distributed
                                    host state): 1/nranks of the matrix crosses each GPU's PCIe link per direction */

/* Library / device info. Returns the number of CUDA devices (>=0) or a negative qf_status. */
int qf_device_count(void);
const char *qf_version(void);
const char *qf_last_error(void);

/* Create / destroy a solver context for N x N matrices and `batch` independent members
 * (batch = 1 for a single simulation; >1 for ensembles where every member has its own
 * convergence — BASELINE config 5).  Builds the Hoppe-Yau coefficient table and its LU
 * factors (quflow/laplacian/cpu.py:55-95 `_compute_cpu_laplacian`, cached by `laplacian()`
 * :604-625; the reference recomputes the LU on every solve, :320-328). */
int qf_create(int N, int batch, int device, qf_handle_t *out);
int qf_destroy(qf_handle_t h);

/* P = Delta_N^{-1} W   — replaces quflow.laplacian.cpu.solve_poisson (cpu.py:681-734,
 * kernel `_solve_cpu_skewh` :281-362), dense skew-Hermitian branch: only the upper triangle
 * of W is read, tr(W)/N is removed from the m=0 right-hand side, P is trace-free and exactly
 * skew-Hermitian.  W_dev and P_dev hold `batch` matrices; they may not alias. */
int qf_solve_poisson(qf_handle_t h, const void *W_dev, void *P_dev, void *stream);

/* Work plan of the Poisson kernel for size N (pure host code, no CUDA call; exported for tests and tooling):
 * params_out[6] = positions per thread, diagonals per band, threads per CTA, CTAs per cluster, positions per CTA,
 * number of units; units_out (may be NULL) receives 8 ints per unit: band of the long piece, its first position, band
 * of the short piece (-1: none), local position where it starts, cluster ranks the long band spans, 0, 0, 0.
 * Returns the number of units (0: N needs more than 8 CTAs per band, which qf_create rejects) or a negative qf_status. */
int qf_poisson_plan(int N, int *params_out, int *units_out, int cap);

/* W = Delta_N P        — replaces quflow.laplacian.cpu.laplace (cpu.py:628-669, kernel
 * `_dot_cpu_generic` :98-108), dense branch, general (not necessarily skew-Hermitian) P. */
int qf_laplace(qf_handle_t h, const void *P_dev, void *W_dev, void *stream);

/* max row sum of |z| per member — np.linalg.norm(W, inf) (isospectral.py:446-448) /
 * scipy.linalg.norm(., ord=inf) (:534).  out_host receives `batch` doubles. Synchronises. */
int qf_norm_inf(qf_handle_t h, const void *W_dev, double *out_host, void *stream);

/* out_host[b] = sum_ij Re(P_ij conj(W_ij)) per member: N * inner_L2(P, W) (quflow/geometry.py:72-76), the building block
 * of norm_L2 (:53-68) and of the loggers energy_euler / enstrophy / inner_Hm1 / inner_H1 (quflow/physics.py:9-38).
 * Deterministic (fixed grid and summation tree).  Synchronises. */
int qf_inner(qf_handle_t h, const void *P_dev, const void *W_dev, double *out_host, void *stream);

/* C = A @ B for N x N complex128 (batch members) with the hand-written DMMA kernel —
 * the np.matmul / zgemm calls at isospectral.py:496,499.  Exposed for tests and micro-benchmarks. */
int qf_zgemm(qf_handle_t h, const void *A_dev, const void *B_dev, void *C_dev, void *stream);

/* Advance W by `steps` isospectral-midpoint steps — replaces
 * quflow.integrators.isospectral.isomp_fixedpoint (isospectral.py:338-613) for the default
 * autonomous Hamiltonian solve_poisson, 2-D state (or `batch` independent states).
 *   W_dev         in/out, overwritten like the reference's W
 *   dt            time step;  epsilon = dt / (2 hbar(N))            (:436-437)
 *   tol           < 0 => 'auto': sqrt(eps) * dt/hbar * ||W||_inf (eps un-rooted with compsum) (:440-452)
 *   maxit, minit  iteration cap / floor (:400-401 asserted by the caller and re-checked here)
 *   flags         QF_FLAG_*
 *   stats         [batch] out, may be NULL
 *   iters_per_step[batch*steps] out (member-major), may be NULL — per-step iteration counts
 * The fixed-point loop, the stopping rule (:523-536) and the update (:547-596) run on the
 * device without host synchronisation; the call synchronises the stream once at the end. */
int qf_isomp(qf_handle_t h, void *W_dev, double dt, int steps, double tol, int maxit, int minit,
             unsigned flags, qf_stats *stats, int32_t *iters_per_step, void *stream);

/* Same as qf_isomp / qf_solve_poisson with HOST buffers: the library copies in, runs, copies out.
 * This is what the ctypes shim calls for numpy inputs (the reference's calling convention). */
int qf_isomp_host(qf_handle_t h, void *W_host, double dt, int steps, double tol, int maxit, int minit,
                  unsigned flags, qf_stats *stats, int32_t *iters_per_step);
int qf_solve_poisson_host(qf_handle_t h, const void *W_host, void *P_host);
int qf_laplace_host(qf_handle_t h, const void *P_host, void *W_host);

/* ---- host-stepped driver -----------------------------------------------------------------
 * The callers' hooks of isomp_fixedpoint run HOST code inside the step: `callback(W, dW)`
 * (isospectral.py:550-551), `forcing(P, W[, time])` (:403-414, :511-520, :594-596),
 * `strang_splitting(dt/2, W)` (:466-467, :602-603) and custom or time-dependent Hamiltonians
 * (:416-424, :488-491).  For them the same device kernels are driven one fixed-point iteration at a
 * time; the host may read or replace the intermediate matrices between the calls.  One member
 * (batch = 1).  On a row-sharded handle (several GPUs) every rank makes the same calls: qf_step_products shards the
 * two GEMMs and completes A and S on every rank (all-gather path), everything else runs replicated on identical
 * bytes, so the host code sees complete matrices on every rank.  Call order per qf_isomp-equivalent run:
 *   qf_step_open                                   dW = 0, tolerance from ||W||_inf      (:430, :440-452)
 *   per step:   [W = strang(dt/2, W)]  qf_step_begin        W~ = W + dW, resnorm = inf   (:470-472, :481-482)
 *     per iteration:  qf_step_hamiltonian  P~ = eps Delta^-1 W~                          (:489, :492)
 *                     (or: write P into QF_BUF_P, then qf_step_scale_p(h, 0))
 *                     qf_step_products     A = P~ W~,  S = A P~                          (:496, :499)
 *                     [qf_step_scale_p(h, 1); FW = forcing(QF_BUF_P, QF_BUF_WHALF)]      (:513-517)
 *                     qf_step_close_iteration   dW = S + (A - A^H) [+ fscale FW], residual and stopping
 *                                               rule; *active = 0 ends the loop          (:503-536)
 *     [qf_step_increment -> callback(W, dW)]  qf_step_update   W += 2 (A - A^H) [+ 2 fscale FW]  (:547-596)
 *   qf_step_stats
 * Every matrix on this path is assumed skew-Hermitian (as the reference's solve_poisson and
 * conj_subtract_ assume); FW is read in full. */
#define QF_BUF_WHALF 0    /* W~ = W + dW, the midpoint state the Hamiltonian and the forcing receive */
#define QF_BUF_P 1        /* P~ (scaled by eps) / P (after qf_step_scale_p(h, 1)) */
#define QF_BUF_SCRATCH 2  /* free N x N buffer (used by the Python side for the callback increment) */
int qf_step_open(qf_handle_t h, const void *W_dev, double dt, double tol, unsigned flags, double *tol_used, void *stream);
int qf_step_begin(qf_handle_t h, const void *W_dev, void *stream);
void *qf_step_buffer(qf_handle_t h, int which);   /* device pointer owned by the handle, NULL for a bad index */
int qf_step_hamiltonian(qf_handle_t h, void *stream);
int qf_step_scale_p(qf_handle_t h, int divide, void *stream);   /* QF_BUF_P *= eps (0) or /= eps (1) */
int qf_step_products(qf_handle_t h, void *stream);
int qf_step_close_iteration(qf_handle_t h, const void *W_dev, const void *F_dev /* may be NULL */, double fscale,
                            int maxit, int minit, int *active, double *resnorm, void *stream);
int qf_step_increment(qf_handle_t h, void *out_dev, void *stream);   /* out = 2 (A - A^H) */
int qf_step_update(qf_handle_t h, void *W_dev, const void *F_dev /* may be NULL */, double fscale, void *stream);
int qf_step_stats(qf_handle_t h, qf_stats *stats, void *stream);

/* ---- matrix <-> real spherical-harmonic coefficients (the callers' data format either side of the path) --------------
 * Replace quflow/quantization.py `mat2shr_parallel_` (:283-325) and `shr2mat_parallel_` (:172-227) for a quantization
 * basis resident in HBM.  basis_dev: the reference's flat float64 array (quflow.quantization.get_basis(N)): for every
 * m = 0..N-1 a dense (N-m) x (N-m) block, row-major [k][el-m], starting at basis_break_index(m, N) (:25-42);
 * qf_basis_size(N) doubles in all.  omega_dev: float64, nomega = (elmax+1)^2 coefficients ordered el^2 + el + m
 * (quflow/utils.py:91-105), elmax <= N-1.  Deterministic (fixed summation order); enqueued on `stream`, no sync. */
long long qf_basis_size(int N);
int qf_mat2shr(qf_handle_t h, const void *W_dev, const void *basis_dev, void *omega_dev, long long nomega, void *stream);
int qf_shr2mat(qf_handle_t h, const void *omega_dev, long long nomega, const void *basis_dev, void *W_dev, void *stream);

/* The tail of the fixed-point iteration (dW = S + A - A^H, W~ = W + dW, residual partial sums; isospectral.py:499-536)
 * can run fused into the epilogue of the second GEMM (1) or as a kernel of its own after it (0, default: measured faster,
 * DESIGN.md).  Also QF_FUSE_POST=1 in the environment at handle creation. */
int qf_set_fuse_post(qf_handle_t h, int enable);

/* Introspection used by bench.py: number of kernels this library launched since creation
 * of the handle, and per-phase device time of the last qf_profile_iteration call. */
int64_t qf_launch_count(qf_handle_t h);
/* Real FP64 flops one GEMM launch EXECUTES on this handle (tiles * tile area * K * 6 or 8), full or upper-only
 * variant, and whether the 3-multiplication (Karatsuba) complex arithmetic is active. For roofline accounting. */
double qf_gemm_executed_flops(qf_handle_t h, int upper_only);
int qf_gemm_is_3m(qf_handle_t h);

/* Roofline denominator of the two GEMMs, measured on `device` at the clocks of the calling run: issue peak of
 * mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) in TFLOP/s, best of `reps` launches timed with CUDA events on `stream`.
 * No reference counterpart (the reference reports no roofline); used by bench.py because MEASURED_PEAKS.json has no
 * FP64 entry. */
int qf_measure_fp64_tensor_peak(int device, int reps, double *tflops_out, void *stream);

typedef struct {
    float poisson_ms;   /* W~ = W + dW, P~ = eps * Delta^{-1} W~ */
    float gemm1_ms;     /* A = P~ W~ */
    float gemm2_ms;     /* S = A P~ (+ fused epilogue when enabled) */
    float post_ms;      /* dW = S + A - A^H, residual row sums, control */
    float update_ms;    /* W += 2 (A - A^H) */
    /* tile-exchange path only (0 otherwise): post_ms split into its parts */
    float x_tail_ms;    /* sharded tail kernel (dW, W~, residual partials of the owned tile pairs) */
    float x_push_ms;    /* copy kernel: owned W~ tiles -> every peer */
    float x_wait_ms;    /* signal + wait for every peer's signal (absorbs the skew between ranks) */
    float x_mirror_ms;  /* lower triangle of W~ rebuilt locally (upper-only exchange) */
    float x_control_ms; /* stopping rule */
} qf_phase_times;

/* Runs `reps` fixed-point iterations on the current state of W_dev (which is left unchanged)
 * with CUDA events around every phase, on `stream`; returns the per-phase averages. */
int qf_profile_iteration(qf_handle_t h, const void *W_dev, double dt, int reps, qf_phase_times *out, void *stream);

/* ---- multi-GPU (one process per GPU) -------------------------------------------------
 * New functionality: the reference is single-process.  ONE large-N simulation is sharded by row
 * blocks of the two GEMMs: the N rows are cut into 2*nranks blocks and rank r owns blocks r and 2*nranks-1-r (which
 * balances the upper-triangular second GEMM).  Ensembles shard per member on the host side and need none of this.
 *
 * Peer-memory data paths (default).  Every rank exports a blob of CUDA IPC handles, the host side all-gathers the blobs
 * (torch.distributed) and every rank imports all of them; the per-iteration communication then runs as plain kernels over
 * the NVLink peer mappings, inside the step graph:
 *   tile exchange (qf_comm_mode 5, default when N is divisible by 128*nranks): the tail of the iteration (dW, W~,
 *     residual) and the update are sharded by tile pairs too.  Per iteration a rank stores the lower tiles of A its peers
 *     need straight from the GEMM epilogue into their memory, and its new W~ tiles and residual partial sums into every
 *     peer's copy; two flag exchanges per iteration are all that is left of a collective.  Only the Poisson solve runs
 *     replicated.
 *   pull all-gather (qf_comm_mode 2): A and S are completed on every rank by kernels that pull the peers' rows; the tail
 *     and the update run replicated.
 * NCCL path (qf_comm_mode 1): the same two gathers as one in-place ncclAllGather each on the compute stream, eager
 * launches; the 128-byte unique id is created on rank 0 and distributed by the host side. */
#define QF_UNIQUE_ID_BYTES 128
int qf_comm_get_unique_id(void *id_out);
int qf_comm_init(qf_handle_t h, const void *unique_id, int rank, int nranks);
#define QF_P2P_BLOB_BYTES 256
int qf_comm_p2p_export(qf_handle_t h, void *blob_out /* QF_P2P_BLOB_BYTES */);
int qf_comm_p2p_import(qf_handle_t h, const void *blobs /* nranks * QF_P2P_BLOB_BYTES */, int rank, int nranks);
/* After the import: 1 selects the tile exchange, 0 the pull all-gather (also: QF_COMM=tile|pull in the environment). */
int qf_comm_set_tile(qf_handle_t h, int enable);
/* Data path in use: 0 none (single GPU / emulated ranks), 1 NCCL all-gather, 2 pull all-gather, 5 tile exchange. */
int qf_comm_mode(qf_handle_t h);
/* Test hook: run the row-sharded GEMM schedule of `nranks` ranks on ONE GPU (all ranks' tiles, no communication). */
int qf_set_emulated_ranks(qf_handle_t h, int nranks);
/* Test hooks for the tile-exchange path on ONE GPU: attach G handles of this process (same device, same N) to each other
 * through plain device pointers, then advance all of them in lock step — every phase of the iteration is enqueued for
 * all ranks before the next phase of any rank, so no kernel waits for a kernel queued behind it.  Arguments as qf_isomp,
 * with one W_dev, one stats entry and `steps` iteration counts per rank. */
int qf_comm_attach_local(qf_handle_t *handles, int G);
int qf_isomp_lockstep(qf_handle_t *handles, int G, void **W_devs, double dt, int steps, double tol, int maxit, int minit,
                      unsigned flags, qf_stats *stats, int32_t *iters_per_step, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* QUFLOW_B200_H */
