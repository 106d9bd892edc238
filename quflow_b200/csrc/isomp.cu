// The isospectral-midpoint fixed-point loop on the device.
//
// Replaces quflow/integrators/isospectral.py:338-613 (`isomp_fixedpoint`) for the default autonomous
// Hamiltonian.  One fixed-point iteration (isospectral.py:475-536) is
//     W~ = W + dW;  P~ = eps * Delta^{-1} W~          (poisson.cu)
//     A  = P~ W~                                       (zgemm.cu)
//     S  = A P~          (skew-Hermitian: only the column blocks touching the upper triangle are computed)
//     dW_new = S + (A - A^H);  res = || dW_old - dW_new ||_inf       (k_post + k_control, this file)
// and the step ends with  W += 2 (A - A^H)  (k_update; isospectral.py:547-596).
// The stopping rule lives in a device-resident control block (QfCtrl): all kernels of the iterations that
// follow convergence return immediately, so a whole step is enqueued without any host synchronisation.
#include <math.h>
#include <stdlib.h>
#include <algorithm>

#include "qf_common.cuh"

namespace {

constexpr int TS = 32;   // tile edge of the transpose-based kernels

// ------------------------------------------------------------------------------- ||W||_inf
__global__ void k_rowsum_abs(const double2 *__restrict__ W, int N, QfCtrl *ctrl)
{
    // one warp per row; max over rows through an order-independent atomicMax on the bit pattern (values >= 0)
    const int b = blockIdx.y;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= N) return;
    const double2 *R = W + (size_t)b * N * N + (size_t)row * N;
    double s = 0.0;
    for (int j = threadIdx.x & 31; j < N; j += 32) s += zabs(R[j]);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) {
        if (!(s == s)) s = INFINITY;   // NaN -> propagate as inf
        atomicMax(reinterpret_cast<unsigned long long *>(&ctrl[b].norm0), (unsigned long long)__double_as_longlong(s));
    }
}

__global__ void k_call_begin(QfCtrl *ctrl, double tol, double tol_factor, int multistate)
{
    QfCtrl &c = ctrl[blockIdx.x];
    const double n0 = multistate ? ctrl[0].norm0 : c.norm0;      // multi-state: ||W[0]||_inf (isospectral.py:444-446)
    c.tol = (tol < 0.0) ? tol_factor * n0 : tol;
    c.resnorm = INFINITY;
    c.resnorm_old = INFINITY;
    c.total_it = 0;
    c.n_maxit = 0;
    c.active = 0;
    c.it = 0;
    c.nonfinite = 0;
    c.steps_done = 0;
    c.resmax_bits = 0ull;
    c.ticket = 0u;
}

// Start of a step for every member (single thread; batch is small).  In graph mode it also arms the WHILE node.
__global__ void k_step_begin(QfCtrl *ctrl, int batch, cudaGraphConditionalHandle cond, int use_cond)
{
    unsigned any = 0;
    for (int b = 0; b < batch; ++b) {
        QfCtrl &c = ctrl[b];
        c.it = 0;
        c.resnorm = INFINITY;          // isospectral.py:470
        c.active = c.nonfinite ? 0 : 1;
        any |= (unsigned)c.active;
    }
    if (use_cond) cudaGraphSetConditional(cond, any);
}

// End of the loop body in graph mode: keep iterating while any member is still active.
__global__ void k_loop_cond(const QfCtrl *ctrl, int batch, cudaGraphConditionalHandle cond)
{
    unsigned any = 0;
    for (int b = 0; b < batch; ++b) any |= (unsigned)ctrl[b].active;
    cudaGraphSetConditional(cond, any);
}

// zero dW of the members that are still alive (reinitialize=True, isospectral.py:471-472)
__global__ void k_zero(double2 *X, size_t n2, const QfCtrl *__restrict__ ctrl)
{
    const int b = blockIdx.y;
    if (ctrl[b].nonfinite) return;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x)
        X[(size_t)b * n2 + i] = make_double2(0.0, 0.0);
}

// ------------------------------------------------------------------------------- post-GEMM
// For the tile pair (bi <= bj):  c = A_ij - conj(A_ji),  d = S_ij + c,  r = |dW_ij - d|,
// dW_ij = d, dW_ji = -conj(d);  deterministic partial row sums of r for the infinity norm:
//   direct[i][bj] = sum_{j in tile, j >= i} r_ij      mirr[j][bi] = sum_{i in tile, i < j} r_ij
// It also writes the next iterate W~ = W + dW (both triangles), so no separate pass is needed before the Poisson solve.
// Tile-exchange path (xg.nranks > 1): a tile pair is processed by the rank that owns row block bi only — the rows of A and
// S it needs are its own, the transposed tile of A was pushed to it by the owner of row block bj during the first GEMM
// (a wait kernel for every peer's "GEMM 1 complete" flag precedes this one) — and the residual partials are stored into
// every peer's copy as well, through the NVLink peer mappings; the new W~ tiles follow in a copy kernel of their own
// (comm.cu: k_xchg_push_wh).  Peer stores need no fence here: the kernel boundary orders them before the signal kernel,
// whose system-scope fence precedes the flag.
template <bool FORCING>
__global__ void __launch_bounds__(256)
k_post(const double2 *__restrict__ Ag, const double2 *__restrict__ Sg, double2 *__restrict__ dWg, double *__restrict__ rowpart,
       int N, int nslots, QfCtrl *__restrict__ ctrl, int hb, int G, const double2 *__restrict__ Wg, double2 *__restrict__ Whg,
       const double2 *__restrict__ Fg, double fscale, const QfXchg xg)
{
    const int b = blockIdx.z;
    if (!ctrl[b].active) return;
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bi > bj) return;
    const bool xpush = xg.nranks > 1;
    if (xpush && qf_owner_of_row(bi * TS, xg.hb, xg.nranks) != xg.rank) return;
    const bool wpush = xpush && xg.push_inline;                                   // W~ tiles go to the peers from here
    const bool wpush_lower = wpush && !(xg.upper_only && ctrl[0].skew_exact);     // ... the mirrored half only if it cannot be rebuilt
    __shared__ double2 T[TS][TS + 1];
    __shared__ double2 D[TS][TS + 1];
    __shared__ double R[TS][TS + 1];
    const size_t off = (size_t)b * N * N;
    const double2 *A = Ag + off;
    const double2 *S = Sg + off;
    double2 *dW = dWg + off;
    const double2 *W = Wg + off;
    double2 *Wh = Whg + off;
    const double2 *F = FORCING ? Fg + off : nullptr;
    double *direct = rowpart + (size_t)b * nslots * N;                              // [member][N][nslots]
    double *mirr = rowpart + ((size_t)gridDim.z + b) * nslots * N;                  // second array, same shape
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;

    // A_ji tile, coalesced along i
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int jj = ty + 8 * q;
        const int j = bj * TS + jj, i = bi * TS + tx;
        T[jj][tx] = (j < N && i < N) ? __ldcg(A + (size_t)qf_prow(j, hb, G) * N + i) : make_double2(0.0, 0.0);
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int ii = ty + 8 * q;
        const int i = bi * TS + ii, j = bj * TS + tx;
        double r = 0.0;
        double2 d = make_double2(0.0, 0.0);
        if (i < N && j < N && i <= j) {
            const size_t ij = (size_t)i * N + j;
            const size_t pij = (size_t)qf_prow(i, hb, G) * N + j;     // A and S use the rank-permuted row layout
            const double2 c = zsub(A[pij], zconj(T[tx][ii]));         // isospectral.py:66-81
            double2 s = S[pij];
            if (i == j) s.x = 0.0;                                    // P W P is skew-Hermitian
            d = zadd(s, c);                                           // :499,:509
            double2 dn = d;
            if (FORCING) dn = zadd(d, zscale(fscale, F[ij]));         // dW += FW * dt/2  (:518-520)
            const double2 old = dW[ij];
            r = zabs(zsub(old, dn));                                  // :526,:534
            dW[ij] = dn;
            const double2 wh = zadd(W[ij], dn);                       // next iterate W~ = W + dW (:481-482)
            Wh[ij] = wh;
            if (wpush)
                for (int p = 0; p < xg.nranks; ++p)
                    if (p != xg.rank) xg.peerWh[p][off + ij] = wh;
        }
        D[ii][tx] = d;                                                // without the forcing term: mirrored below
        R[ii][tx] = r;
        // direct row sum over the 32 columns of this tile (fixed shuffle tree => deterministic)
        double s = r;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (tx == 0 && i < N) {
            direct[(size_t)i * nslots + bj] = s;
            if (xpush)
                for (int p = 0; p < xg.nranks; ++p)
                    if (p != xg.rank) xg.peerPart[p][(size_t)(direct - rowpart) + (size_t)i * nslots + bj] = s;
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int jj = ty + 8 * q;
        const int j = bj * TS + jj, i = bi * TS + tx;
        double rl = 0.0;
        if (j < N && i < N && i < j) {
            const size_t ji = (size_t)j * N + i;
            const double2 d = D[tx][jj];
            double2 dm = make_double2(-d.x, d.y);
            if (FORCING) {                                            // the forcing term of the lower triangle is its own
                dm = zadd(dm, zscale(fscale, F[ji]));
                rl = zabs(zsub(dW[ji], dm));
            } else {
                rl = R[tx][jj];
            }
            dW[ji] = dm;
            const double2 wh = zadd(W[ji], dm);
            Wh[ji] = wh;
            if (wpush_lower)
                for (int p = 0; p < xg.nranks; ++p)
                    if (p != xg.rank) xg.peerWh[p][off + ji] = wh;
        }
        double s = rl;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (tx == 0 && j < N) {
            mirr[(size_t)j * nslots + bi] = s;
            if (xpush)
                for (int p = 0; p < xg.nranks; ++p)
                    if (p != xg.rank) xg.peerPart[p][(size_t)(mirr - rowpart) + (size_t)j * nslots + bi] = s;
        }
    }
}

// ------------------------------------------------------------------------------- control
// Row sums of the residual (one warp per row, slots contiguous), max over rows through an order-independent
// atomicMax, and the stopping rule evaluated by the last block to finish (ticket counter).
// The two partial-sum arrays are [member][N][na] and [member][N][nb] (k_post: na = nb = nslots; fused GEMM-2 tail: nsd, nsm).
__global__ void __launch_bounds__(256)
k_control(const double *__restrict__ parta, int na, const double *__restrict__ partb, int nb, int N, QfCtrl *ctrl, int maxit,
          int minit, int nfollow, cudaGraphConditionalHandle cond, int use_cond)
{
    const int b = blockIdx.y;
    QfCtrl &c = ctrl[b];
    if (!c.active) return;
    const double *direct = parta + (size_t)b * na * N;
    const double *mirr = partb + (size_t)b * nb * N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x * 8 + warp;
    double s = 0.0;
    if (row < N) {
        for (int k = lane; k < na; k += 32) s += direct[(size_t)row * na + k];
        for (int k = lane; k < nb; k += 32) s += mirr[(size_t)row * nb + k];
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ int last;
    if (lane == 0 && row < N) {
        if (!(s == s)) s = INFINITY;
        atomicMax(&c.resmax_bits, (unsigned long long)__double_as_longlong(s));
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(&c.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!last || threadIdx.x != 0) return;
    __threadfence();
    const double best = __longlong_as_double((long long)atomicExch(&c.resmax_bits, 0ull));
    c.ticket = 0;
    const int nan_seen = isinf(best) ? 1 : 0;
    c.it += 1;                                   // isospectral.py:478
    c.total_it += 1;
    c.gseq += 1;
    int active = 1;
    if (c.it >= minit) {                         // :523
        c.resnorm_old = c.resnorm;               // :525
        c.resnorm = best;
        if (nan_seen) {                          // scipy.linalg.norm -> ValueError (:534)
            c.nonfinite = 1;
            active = 0;
        } else if (best <= c.tol || best >= c.resnorm_old) {   // :535
            active = 0;
        }
    }
    if (active && c.it >= maxit) {               // for/else, :538-540
        active = 0;
        c.n_maxit += 1;
    }
    // multi-state run: members 1..nfollow take every decision from member 0 (isospectral.py:529-531)
    for (int f = 1; f <= nfollow; ++f) {
        QfCtrl &d = ctrl[f];
        d.it = c.it;
        d.total_it = c.total_it;
        d.gseq = c.gseq;
        d.resnorm = c.resnorm;
        d.resnorm_old = c.resnorm_old;
        d.n_maxit = c.n_maxit;
        d.nonfinite = c.nonfinite;
    }
    __threadfence();
    c.active = active;
    for (int f = 1; f <= nfollow; ++f) ctrl[f].active = active;
    // one deciding member (single run or multi-state): arm the WHILE node of the step graph right here instead of in a
    // separate k_loop_cond launch
    if (use_cond) cudaGraphSetConditional(cond, (unsigned)active);
}

// multi-state run: P~ of member 0 is everybody's stream function (select_first, cpu.py:672-674)
__global__ void k_bcast_p(double2 *P, size_t n2, int batch, const QfCtrl *__restrict__ ctrl)
{
    if (!ctrl[0].active) return;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
        const double2 v = P[i];
        for (int b = 1; b < batch; ++b) P[(size_t)b * n2 + i] = v;
    }
}

// ------------------------------------------------------------------------------- update
// W += 2 (A - A^H)   (isospectral.py:547, :592) or its Kahan-compensated form (:553-586).  Both triangles of W are
// updated element by element (like the reference's full-matrix `W += PWcomm`), with the increment of the lower
// triangle being the exact mirror -conj(.) of the upper one (isospectral.py:74).  The kernel also prepares the first
// iterate of the next step, W~ = W_new + dW (or W_new when the iterate is re-initialised every step).
__device__ __forceinline__ double2 kahan_add(double2 w, double2 inc, double2 &kc)
{
    const double2 y = zsub(inc, kc);          // :570-571
    const double2 tt = zadd(w, y);            // :575-576
    kc = zsub(zsub(tt, w), y);                // :580-583
    return tt;                                // :586
}

// Tile-exchange path (xg.nranks > 1): a tile pair is updated by the rank that owns row block bi only (A, dW and the state
// are valid there); the first iterate of the next step, W~, goes to the peers in a copy kernel of its own (k_xchg_push_wh).
template <bool COMPSUM, bool FORCING>
__global__ void __launch_bounds__(256)
k_update(const double2 *__restrict__ Ag, double2 *__restrict__ Wg, double2 *__restrict__ Kg, int N, QfCtrl *ctrl,
         int32_t *iters, int steps_cap, int hb, int G, const double2 *__restrict__ dWg, double2 *__restrict__ Whg, int reinit,
         const double2 *__restrict__ Fg, double fscale, const QfXchg xg)
{
    const int b = blockIdx.z;
    QfCtrl &c = ctrl[b];
    if (c.nonfinite) return;
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bi == 0 && bj == 0 && threadIdx.x == 0) {
        if (iters && c.steps_done < steps_cap) iters[(size_t)b * steps_cap + c.steps_done] = c.it;
        c.steps_done += 1;
    }
    if (bi > bj) return;
    const bool xpush = xg.nranks > 1;
    if (xpush && qf_owner_of_row(bi * TS, xg.hb, xg.nranks) != xg.rank) return;
    const bool wpush = xpush && xg.push_inline;
    const bool wpush_lower = wpush && !(xg.upper_only && c.skew_exact);
    __shared__ double2 T[TS][TS + 1];
    const size_t off = (size_t)b * N * N;
    const double2 *A = Ag + off;
    double2 *W = Wg + off;
    double2 *K = COMPSUM ? Kg + off : nullptr;
    const double2 *dW = dWg + off;
    double2 *Wh = Whg + off;
    const double2 *F = FORCING ? Fg + off : nullptr;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int jj = ty + 8 * q;
        const int j = bj * TS + jj, i = bi * TS + tx;
        T[jj][tx] = (j < N && i < N) ? __ldcg(A + (size_t)qf_prow(j, hb, G) * N + i) : make_double2(0.0, 0.0);
    }
    __syncthreads();
    double2 cv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int ii = ty + 8 * q;
        const int i = bi * TS + ii, j = bj * TS + tx;
        double2 cm = make_double2(0.0, 0.0);
        if (i < N && j < N && i <= j) {
            const size_t ij = (size_t)i * N + j;
            cm = zsub(A[(size_t)qf_prow(i, hb, G) * N + j], zconj(T[tx][ii]));
            cm = make_double2(2.0 * cm.x, 2.0 * cm.y);                 // :547
            double2 w = W[ij];
            if (COMPSUM) {
                double2 kc = K[ij];
                w = kahan_add(w, cm, kc);
                K[ij] = kc;
            } else {
                w = zadd(w, cm);                                       // :592
                if (FORCING) w = zadd(w, zscale(2.0, zscale(fscale, F[ij])));   // FW *= 2; W += FW (:594-596)
            }
            W[ij] = w;
            const double2 wh = reinit ? w : zadd(w, dW[ij]);
            Wh[ij] = wh;
            if (wpush)
                for (int p = 0; p < xg.nranks; ++p)
                    if (p != xg.rank) xg.peerWh[p][off + ij] = wh;
        }
        cv[q] = cm;
    }
    __syncthreads();   // everyone is done reading T
#pragma unroll
    for (int q = 0; q < 4; ++q) T[ty + 8 * q][tx] = cv[q];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int jj = ty + 8 * q;
        const int j = bj * TS + jj, i = bi * TS + tx;
        if (j < N && i < N && i < j) {
            const size_t ji = (size_t)j * N + i;
            const double2 cu = T[tx][jj];
            const double2 cm = make_double2(-cu.x, cu.y);              // PWcomm[j,i] = -conj(PWcomm[i,j])
            double2 w = W[ji];
            if (COMPSUM) {
                double2 kc = K[ji];
                w = kahan_add(w, cm, kc);
                K[ji] = kc;
            } else {
                w = zadd(w, cm);
                if (FORCING) w = zadd(w, zscale(2.0, zscale(fscale, F[ji])));
            }
            W[ji] = w;
            const double2 wh = reinit ? w : zadd(w, dW[ji]);
            Wh[ji] = wh;
            if (wpush_lower)
                for (int p = 0; p < xg.nranks; ++p)
                    if (p != xg.rank) xg.peerWh[p][off + ji] = wh;
        }
    }
}

// out = 2 (A - A^H): the increment handed to `callback(W, dW)` just before the update (isospectral.py:547-551)
__global__ void __launch_bounds__(256)
k_increment(const double2 *__restrict__ Ag, double2 *__restrict__ Og, int N, int hb, int G)
{
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bi > bj) return;
    __shared__ double2 T[TS][TS + 1];
    const size_t off = (size_t)blockIdx.z * N * N;
    const double2 *A = Ag + off;
    double2 *O = Og + off;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int jj = ty + 8 * q;
        const int j = bj * TS + jj, i = bi * TS + tx;
        T[jj][tx] = (j < N && i < N) ? A[(size_t)qf_prow(j, hb, G) * N + i] : make_double2(0.0, 0.0);
    }
    __syncthreads();
    double2 cv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int ii = ty + 8 * q;
        const int i = bi * TS + ii, j = bj * TS + tx;
        double2 cm = make_double2(0.0, 0.0);
        if (i < N && j < N && i <= j) {
            cm = zsub(A[(size_t)qf_prow(i, hb, G) * N + j], zconj(T[tx][ii]));
            cm = make_double2(2.0 * cm.x, 2.0 * cm.y);
            O[(size_t)i * N + j] = cm;
        }
        cv[q] = cm;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) T[ty + 8 * q][tx] = cv[q];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int jj = ty + 8 * q;
        const int j = bj * TS + jj, i = bi * TS + tx;
        if (j < N && i < N && i < j) {
            const double2 cu = T[tx][jj];
            O[(size_t)j * N + i] = make_double2(-cu.x, cu.y);
        }
    }
}

// X *= s  or  X /= s  (Phalf *= vareps, isospectral.py:492;  Phalf /= vareps, :513)
__global__ void k_scale(double2 *X, size_t n, double s, int divide)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double2 v = X[i];
        X[i] = divide ? make_double2(v.x / s, v.y / s) : make_double2(v.x * s, v.y * s);
    }
}

// ------------------------------------------------------------------------------- <P, W>_L2
// sum_ij Re(P_ij conj(W_ij)) per member (quflow/geometry.py:72-76 without the 1/N): fixed grid, fixed tree => the
// result does not depend on scheduling.  Stage 1: per-block partial sums; stage 2: one block adds them in order.
constexpr int INNER_BLOCKS = 256;
__global__ void __launch_bounds__(256)
k_inner_partial(const double2 *__restrict__ P, const double2 *__restrict__ W, size_t n2, double *__restrict__ part)
{
    const int b = blockIdx.y;
    const double2 *p = P + (size_t)b * n2, *w = W + (size_t)b * n2;
    double s = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
        const double2 a = p[i], c = w[i];
        s += a.x * c.x + a.y * c.y;
    }
    __shared__ double sh[8];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < 8; ++q) t += sh[q];
        part[(size_t)b * gridDim.x + blockIdx.x] = t;
    }
}
__global__ void k_inner_final(const double *__restrict__ part, int nblocks, double *__restrict__ out)
{
    const int b = blockIdx.x;
    double s = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 32) s += part[(size_t)b * nblocks + i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) out[b] = s;
}

}   // namespace

// out_host[b] = sum_ij Re(P_ij conj(W_ij)) of member b; the caller divides by N (inner_L2) or takes the root (norm_L2).
extern "C" int qf_inner(qf_handle_t h, const void *P_dev, const void *W_dev, double *out_host, void *stream)
{
    if (!h || !P_dev || !W_dev || !out_host) { qf_set_error("qf_inner: null argument"); return QF_ERR_INVALID; }
    QF_ON_DEVICE(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int B = h->batch;
    if (!h->inner_part) QF_CUDA(cudaMalloc(&h->inner_part, sizeof(double) * (size_t)B * (INNER_BLOCKS + 1)));
    double *part = h->inner_part, *res = h->inner_part + (size_t)B * INNER_BLOCKS;
    k_inner_partial<<<dim3(INNER_BLOCKS, B), 256, 0, st>>>((const double2 *)P_dev, (const double2 *)W_dev, h->mat_elems, part);
    k_inner_final<<<B, 32, 0, st>>>(part, INNER_BLOCKS, res);
    h->launches += 2;
    QF_CUDA(cudaGetLastError());
    QF_CUDA(cudaMemcpyAsync(out_host, res, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
    QF_CUDA(cudaStreamSynchronize(st));
    return QF_OK;
}

// ------------------------------------------------------------------------------- launchers
int qf_launch_norm_inf(qf_handle_s *h, const double2 *W, cudaStream_t st)
{
    const int N = h->N;
    for (int b = 0; b < h->batch; ++b)
        QF_CUDA(cudaMemsetAsync(&h->ctrl[b].norm0, 0, sizeof(double), st));
    dim3 g((N + 7) / 8, h->batch);
    k_rowsum_abs<<<g, 256, 0, st>>>(W, N, h->ctrl);
    h->launches++;
    QF_CUDA(cudaGetLastError());
    return QF_OK;
}

static inline double qf_hbar(int N) { return 2.0 / sqrt((double)N * (double)N - 1.0); }   // quflow/geometry.py:7-9

// One fixed-point iteration, in three phases.  A single rank runs them back to back; the lock-step emulation of the
// tile-exchange path on one GPU (qf_isomp_lockstep, tests) runs phase A for every rank, then B, then C, so that no kernel
// ever waits for a kernel that is queued behind it.
//   A  P~ = eps Delta^-1 W~;  A = P~ W~ (own row blocks)  [tile exchange: lower tiles -> the owner of their column block; signal]
//   B  S = A P~ with the fused tail: dW, W~, residual partials   [tile exchange: W~ tiles + partials -> every peer; signal]
//      (legacy all-gather paths and QF_FUSE_POST=0: plain second GEMM, gathers, k_post)
//   C  [tile exchange: wait for every peer's signal]  stopping rule (k_control)
enum { QF_PH_A = 1, QF_PH_B = 2, QF_PH_C = 4, QF_PH_ALL = 7 };

int qf_enqueue_iteration(qf_handle_s *h, const double2 *W, double eps, int maxit, int minit, cudaStream_t st,
                         cudaEvent_t *ev /* 12 events or null: 0-4 phases, 7-11 the exchange of the tile path */, int phases = QF_PH_ALL)
{
    const int N = h->N;
    const int G = h->nranks;
    const bool xmode = (G > 1 && h->comm_mode == 5);                    // tile exchange
    const bool legacy_comm = (G > 1 && (h->comm_mode == 1 || h->comm_mode == 2));
    const int my = (xmode || legacy_comm) ? h->rank : -1;             // -1: compute every rank's blocks here (single GPU / emulation)
    const bool permuted = G > 1 && !xmode;                            // legacy paths store A and S with rank-permuted rows
    const int hb = permuted ? qf_block_rows(N, G) : N;                // qf_prow arguments of k_post (identity unless permuted)
    const int Gp = permuted ? G : 1;
    const bool fused = h->fuse_post && (G == 1 || xmode) && qf_gemm_can_fuse_post(h);
    const QfXchg *xg = xmode ? qf_xchg_desc(h) : nullptr;
    const bool multistate = h->multistate && h->batch > 1;
    if (phases & QF_PH_A) {
        if (ev) QF_CUDA(cudaEventRecord(ev[0], st));
        // Tile exchange, upper-only W~ exchange: the lower triangle of W~ is rebuilt locally from the upper tiles the peers
        // pushed.  Only the first GEMM reads it — the Poisson solve reads the upper triangle — so inside the step graph the
        // mirror runs on a forked branch next to the solve and joins before the GEMM.
        bool forked = false;
        if (xmode && xg->upper_only) {
            if (ev) QF_CUDA(cudaEventRecord(ev[9], st));
            cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
            QF_CUDA(cudaStreamIsCapturing(st, &cs));
            if (cs == cudaStreamCaptureStatusActive) {
                if (!h->side_stream) {
                    QF_CUDA(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
                    QF_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
                    QF_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
                }
                QF_CUDA(cudaEventRecord(h->ev_fork, st));
                QF_CUDA(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
                QF_CHECK(qf_xchg_mirror_wh(h, true, h->side_stream));
                QF_CUDA(cudaEventRecord(h->ev_join, h->side_stream));
                forked = true;
            } else {
                QF_CHECK(qf_xchg_mirror_wh(h, true, st));
            }
            if (ev) QF_CUDA(cudaEventRecord(ev[10], st));
        }
        // W~ = W + dW was written by the previous iteration / update (or copied at call start): solve straight from it
        QF_CHECK(qf_launch_poisson(h, h->Wh, nullptr, h->Wh, h->P, eps, true, st, multistate ? 1 : h->batch));
        if (forked) QF_CUDA(cudaStreamWaitEvent(st, h->ev_join, 0));
        if (multistate) {
            const size_t n2 = h->mat_elems;
            k_bcast_p<<<(unsigned)std::min<size_t>((n2 + 255) / 256, (size_t)h->sm_count * 8), 256, 0, st>>>(h->P, n2, h->batch, h->ctrl);
            h->launches++;
        }
        if (ev) QF_CUDA(cudaEventRecord(ev[1], st));
        QF_CHECK(qf_launch_zgemm(h, h->P, h->Wh, h->A, false, true, my, G, false, st, !permuted, xg));   // rows of A = P~ W~
        if (xmode) QF_CHECK(qf_xchg_signal(h, QF_XF_G1, true, st));
        // (Running the A gather on a forked branch next to the second GEMM was measured twice: no gain — the cooperative
        // GEMM launch waits for the gather's CTAs — so the gathers stay in stream order.)
        if (legacy_comm) QF_CHECK(h->comm_mode == 2 ? qf_comm_p2p_allgather(h, 0, true, st) : qf_comm_allgather_rows(h, h->A, st));
        if (ev) QF_CUDA(cudaEventRecord(ev[2], st));
    }
    const dim3 gc((N + 7) / 8, multistate ? 1 : h->batch);
    const int nfollow = multistate ? h->batch - 1 : 0;
    double *part_direct = h->rowpart2, *part_mirror = h->rowpart2 + (size_t)h->batch * N * h->nsd;
    if (fused) {
        if (phases & QF_PH_B) {
            // GEMM 2 with the fused tail: dW = S + (A - A^H), W~ = W + dW and the residual partials come straight from the
            // accumulators of the upper tiles of S = A P~ (QfEpiPost); only the stopping rule is left as a launch of its own
            QfEpiPost epi;
            epi.A = h->A;
            epi.dW = h->dW;
            epi.W = W;
            epi.Wh = h->Wh;
            epi.part_direct = part_direct;
            epi.part_mirror = part_mirror;
            epi.nsd = h->nsd;
            epi.nsm = h->nsm;
            QF_CHECK(qf_launch_zgemm_post(h, h->A, h->P, epi, true, my, G, st, xg));
            if (ev) QF_CUDA(cudaEventRecord(ev[3], st));
            if (xmode) {
                QF_CHECK(qf_xchg_push_wh(h, true, st));
                if (phases != QF_PH_ALL) QF_CHECK(qf_xchg_signal(h, QF_XF_X, true, st));
            }
        }
        if (phases & QF_PH_C) {
            if (xmode) {
                // every peer's W~ tiles and partials have landed
                QF_CHECK(phases == QF_PH_ALL ? qf_xchg_barrier(h, QF_XF_X, true, st) : qf_xchg_wait(h, QF_XF_X, true, st));
            }
            k_control<<<gc, 256, 0, st>>>(part_direct, h->nsd, part_mirror, h->nsm, N, h->ctrl, maxit, minit, nfollow, h->cap_cond,
                                          h->cap_use_cond);
            h->launches += 1;
            if (ev) QF_CUDA(cudaEventRecord(ev[4], st));
        }
        QF_CUDA(cudaGetLastError());
        return QF_OK;
    }
    const int nb = (N + TS - 1) / TS;
    const dim3 g(nb, nb, h->batch);
    const QfXchg solo;
    if (phases & QF_PH_B) {
        QF_CHECK(qf_launch_zgemm(h, h->A, h->P, h->S, true, true, my, G, permuted, st, !permuted, nullptr));   // rows of S = A P~ (A rows are local)
        if (legacy_comm) QF_CHECK(h->comm_mode == 2 ? qf_comm_p2p_allgather(h, 1, true, st) : qf_comm_allgather_rows(h, h->S, st));
        if (ev) QF_CUDA(cudaEventRecord(ev[3], st));
        if (xmode) {
            // the owners of the tile pairs form dW, W~ and the residual partials and push W~ / the partials to every peer;
            // the transposed A tiles they read were pushed by the peers during THEIR first GEMM: wait for those first
            QF_CHECK(qf_xchg_wait(h, QF_XF_G1, true, st));
            k_post<false><<<g, 256, 0, st>>>(h->A, h->S, h->dW, h->rowpart, N, h->nslots, h->ctrl, hb, Gp, W, h->Wh, nullptr, 0.0, *xg);
            h->launches++;
            if (ev) QF_CUDA(cudaEventRecord(ev[7], st));
            QF_CHECK(qf_xchg_push_wh(h, true, st));
            if (ev) QF_CUDA(cudaEventRecord(ev[8], st));
            if (phases != QF_PH_ALL) QF_CHECK(qf_xchg_signal(h, QF_XF_X, true, st));
        }
    }
    if (phases & QF_PH_C) {
        if (xmode) {
            QF_CHECK(phases == QF_PH_ALL ? qf_xchg_barrier(h, QF_XF_X, true, st) : qf_xchg_wait(h, QF_XF_X, true, st));
            if (ev) QF_CUDA(cudaEventRecord(ev[11], st));
        } else {
            k_post<false><<<g, 256, 0, st>>>(h->A, h->S, h->dW, h->rowpart, N, h->nslots, h->ctrl, hb, Gp, W, h->Wh, nullptr, 0.0, solo);
            h->launches++;
        }
        k_control<<<gc, 256, 0, st>>>(h->rowpart, h->nslots, h->rowpart + (size_t)h->batch * h->nslots * N, h->nslots, N, h->ctrl, maxit,
                                      minit, nfollow, h->cap_cond, h->cap_use_cond);
        h->launches++;
        if (ev) QF_CUDA(cudaEventRecord(ev[4], st));
    }
    QF_CUDA(cudaGetLastError());
    return QF_OK;
}

// End of a step: W += 2 (A - A^H).  Tile exchange: every rank updates the tile pairs it owns and pushes the next step's
// first iterate W~ to its peers (phase 1), then waits until everybody's tiles have landed (phase 2).
int qf_enqueue_update(qf_handle_s *h, double2 *W, bool compsum, bool reinit, cudaStream_t st, int phases = 3)
{
    const int N = h->N;
    const int nb = (N + TS - 1) / TS;
    const bool xmode = (h->nranks > 1 && h->comm_mode == 5);
    const bool permuted = h->nranks > 1 && !xmode;
    const int hb = permuted ? qf_block_rows(N, h->nranks) : N;
    const int Gp = permuted ? h->nranks : 1;
    dim3 g(nb, nb, h->batch);
    const QfXchg solo;
    const QfXchg xg = xmode ? *qf_xchg_desc(h) : solo;
    if (phases & 1) {
        if (compsum)
            k_update<true, false><<<g, 256, 0, st>>>(h->A, W, h->kahan_c, N, h->ctrl, h->iters_dev, h->steps_cap, hb, Gp, h->dW, h->Wh, reinit ? 1 : 0, nullptr, 0.0, xg);
        else
            k_update<false, false><<<g, 256, 0, st>>>(h->A, W, nullptr, N, h->ctrl, h->iters_dev, h->steps_cap, hb, Gp, h->dW, h->Wh, reinit ? 1 : 0, nullptr, 0.0, xg);
        h->launches++;
        QF_CUDA(cudaGetLastError());
        if (xmode) {
            QF_CHECK(qf_xchg_push_wh(h, false, st));
            if (phases != 3) QF_CHECK(qf_xchg_signal(h, QF_XF_X, false, st));
        }
    }
    if ((phases & 2) && xmode) {
        QF_CHECK(phases == 3 ? qf_xchg_barrier(h, QF_XF_X, false, st) : qf_xchg_wait(h, QF_XF_X, false, st));
        // (the lower triangle of the new W~ is rebuilt at the start of the next iteration, next to its Poisson solve)
    }
    return QF_OK;
}

// ------------------------------------------------------------------------------- step graph
// One CUDA graph per step:  k_step_begin -> [k_zero] -> WHILE(any member active) { one fixed-point iteration } -> k_update.
// The WHILE node is a CUDA conditional node whose handle is set on the device (k_step_begin, then k_control or k_loop_cond), so the
// number of iterations is decided by the GPU and no launch is wasted on iterations after convergence.
struct QfStepGraph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaGraphConditionalHandle cond = 0;
    // key
    const void *W = nullptr;
    double eps = 0.0;
    int maxit = 0, minit = 0, compsum = 0, reinit = 0, nranks = 0, rank = 0;
    const void *iters = nullptr;
    int steps_cap = 0;
    const void *kahan = nullptr;
    int kernels_per_iter = 0, kernels_per_step = 0;
};

static void free_step_graph(QfStepGraph *g)
{
    if (!g) return;
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
}

void qf_graph_destroy(qf_handle_s *h)
{
    free_step_graph(reinterpret_cast<QfStepGraph *>(h->step_graph));
    h->step_graph = nullptr;
    if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
    h->cap_stream = nullptr;
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    h->side_stream = nullptr;
    h->ev_fork = h->ev_join = nullptr;

}

template <typename... Args>
static cudaError_t add_kernel_node(cudaGraphNode_t *node, cudaGraph_t graph, const cudaGraphNode_t *deps, size_t ndeps,
                                   void *func, dim3 grid, dim3 block, Args... args)
{
    void *argv[] = {(void *)&args...};
    cudaKernelNodeParams p = {};
    p.func = func;
    p.gridDim = grid;
    p.blockDim = block;
    p.sharedMemBytes = 0;
    p.kernelParams = argv;
    p.extra = nullptr;
    return cudaGraphAddKernelNode(node, graph, deps, ndeps, &p);
}

static int build_step_graph(qf_handle_s *h, double2 *W, double eps, int maxit, int minit, bool compsum, bool reinit,
                            QfStepGraph **out)
{
    QfStepGraph *old = reinterpret_cast<QfStepGraph *>(h->step_graph);
    if (old && old->W == W && old->eps == eps && old->maxit == maxit && old->minit == minit && old->compsum == (int)compsum &&
        old->reinit == (int)reinit && old->nranks == h->nranks && old->rank == h->rank && old->iters == h->iters_dev &&
        old->steps_cap == h->steps_cap && old->kahan == h->kahan_c) {
        *out = old;
        return QF_OK;
    }
    free_step_graph(old);
    h->step_graph = nullptr;
    if (!h->cap_stream) QF_CUDA(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));

    const int B = h->batch;
    const size_t n2 = h->mat_elems;
    QfStepGraph *g = new QfStepGraph();
    g->W = W; g->eps = eps; g->maxit = maxit; g->minit = minit; g->compsum = compsum; g->reinit = reinit;
    g->nranks = h->nranks; g->rank = h->rank; g->iters = h->iters_dev; g->steps_cap = h->steps_cap; g->kahan = h->kahan_c;
#define QF_G(call)                                                                                \
    do {                                                                                          \
        cudaError_t _e = (call);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            qf_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            free_step_graph(g);                                                                   \
            return QF_ERR_CUDA;                                                                   \
        }                                                                                         \
    } while (0)
    QF_G(cudaGraphCreate(&g->graph, 0));
    QF_G(cudaGraphConditionalHandleCreate(&g->cond, g->graph, 1, cudaGraphCondAssignDefault));

    cudaGraphNode_t n_begin, n_zero, n_while, last;
    QfCtrl *ctrl = h->ctrl;
    int use_cond = 1;
    QF_G(add_kernel_node(&n_begin, g->graph, nullptr, 0, (void *)k_step_begin, dim3(1), dim3(1), ctrl, B, g->cond, use_cond));
    last = n_begin;
    g->kernels_per_step = 2;
    if (reinit) {
        const dim3 gz((unsigned)std::min<size_t>((n2 + 255) / 256, (size_t)h->sm_count * 8), B);
        double2 *dW = h->dW;
        size_t n2v = n2;
        const QfCtrl *cc = h->ctrl;
        QF_G(add_kernel_node(&n_zero, g->graph, &last, 1, (void *)k_zero, gz, dim3(256), dW, n2v, cc));
        last = n_zero;
        g->kernels_per_step = 3;
    }
    cudaGraphNodeParams wp = {};
    wp.type = cudaGraphNodeTypeConditional;
    wp.conditional.handle = g->cond;
    wp.conditional.type = cudaGraphCondTypeWhile;
    wp.conditional.size = 1;
    QF_G(cudaGraphAddNode(&n_while, g->graph, &last, 1, &wp));
    cudaGraph_t body = wp.conditional.phGraph_out[0];

    // loop body by stream capture (the launchers below are the same ones the eager path uses)
    const long long l0 = h->launches;
    QF_G(cudaStreamBeginCaptureToGraph(h->cap_stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    // with ONE deciding member (single run, multi-state) k_control arms the WHILE node itself; ensembles need the OR
    // over their members: k_loop_cond
    const bool single_decider = (B == 1) || h->multistate;
    h->cap_cond = g->cond;
    h->cap_use_cond = single_decider ? 1 : 0;
    int rc = qf_enqueue_iteration(h, W, eps, maxit, minit, h->cap_stream, nullptr);
    h->cap_use_cond = 0;
    if (rc == QF_OK && !single_decider) {
        k_loop_cond<<<1, 1, 0, h->cap_stream>>>(h->ctrl, B, g->cond);
        h->launches++;
    }
    cudaGraph_t captured = nullptr;
    cudaError_t ce = cudaStreamEndCapture(h->cap_stream, &captured);
    g->kernels_per_iter = (int)(h->launches - l0);
    h->launches = l0;
    if (rc != QF_OK) { free_step_graph(g); return rc; }
    QF_G(ce);

    // end of the step, again by stream capture: k_update (+ the exchange of the next W~ on the tile-exchange path)
    {
        const long long l1 = h->launches;
        QF_G(cudaStreamBeginCaptureToGraph(h->cap_stream, g->graph, &n_while, nullptr, 1, cudaStreamCaptureModeRelaxed));
        rc = qf_enqueue_update(h, W, compsum, reinit, h->cap_stream);
        ce = cudaStreamEndCapture(h->cap_stream, &captured);
        g->kernels_per_step += (int)(h->launches - l1) - 1;     // the plain k_update is already counted
        h->launches = l1;
        if (rc != QF_OK) { free_step_graph(g); return rc; }
        QF_G(ce);
    }
    QF_G(cudaGraphInstantiate(&g->exec, g->graph, 0));
#undef QF_G
    h->step_graph = g;
    *out = g;
    return QF_OK;
}

// ------------------------------------------------------------------------------- one qf_isomp call, in pieces
// (shared by qf_isomp, the host-buffer entry point and the lock-step emulation of several ranks on one GPU)
struct IsompCall {
    double2 *W = nullptr;       // the buffer the kernels advance: the caller's W_dev, or the handle's Wst (tile exchange)
    double eps = 0.0;
    int maxit = 0, minit = 0, steps = 0;
    bool compsum = false, reinit = false, xmode = false;
};

// Call start (isospectral.py:426-459): work buffers, tolerance from ||W||_inf, control block.  W_dev == nullptr (tile
// exchange only): the state was already staged in Wst by the caller (row-sharded host upload).
static int isomp_call_begin(qf_handle_s *h, void *W_dev, double dt, int steps, double tol, int maxit, int minit, unsigned flags,
                            cudaStream_t st, IsompCall *c)
{
    if (minit < 1) { qf_set_error("minit must be at least 1."); return QF_ERR_INVALID; }       // isospectral.py:400
    if (maxit < minit) { qf_set_error("maxit must be at minit."); return QF_ERR_INVALID; }     // isospectral.py:401
    if (steps < 0) { qf_set_error("steps must be non-negative"); return QF_ERR_INVALID; }
    const int N = h->N, B = h->batch;
    const size_t n2 = h->mat_elems;
    c->xmode = h->nranks > 1 && h->comm_mode == 5;
    if (!W_dev && !c->xmode) { qf_set_error("qf_isomp: null handle or pointer"); return QF_ERR_INVALID; }
    c->compsum = (flags & QF_FLAG_COMPSUM) != 0;
    c->reinit = (flags & QF_FLAG_REINITIALIZE) != 0;
    c->maxit = maxit;
    c->minit = minit;
    c->steps = steps;
    const bool multistate = (flags & QF_FLAG_MULTISTATE) != 0 && B > 1;
    if (multistate && h->nranks > 1) { qf_set_error("multi-state runs are single-GPU"); return QF_ERR_UNSUPPORTED; }
    if ((int)multistate != h->multistate) {
        h->multistate = multistate ? 1 : 0;
        qf_graph_destroy(h);        // the step graph bakes the member coupling in
    }
    if (steps > h->steps_cap) {
        if (h->iters_dev) QF_CUDA(cudaFree(h->iters_dev));
        h->steps_cap = std::max(steps, 1024);
        QF_CUDA(cudaMalloc(&h->iters_dev, sizeof(int32_t) * (size_t)B * h->steps_cap));
    }
    if (c->compsum && !h->kahan_c) QF_CUDA(cudaMalloc(&h->kahan_c, sizeof(double2) * n2 * B));

    const double hb = qf_hbar(N);
    c->eps = dt / (2.0 * hb);                                               // isospectral.py:436-437
    double mach_eps = 2.220446049250313e-16;                                // np.finfo(complex128).eps
    if (!c->compsum) mach_eps = sqrt(mach_eps);                             // :441-443
    const double tol_factor = mach_eps * dt / hb;                           // :448

    c->W = (double2 *)W_dev;
    if (c->xmode) {
        // tile exchange: the state lives in the handle's peer-visible buffer for the duration of the call
        if (W_dev) QF_CUDA(cudaMemcpyAsync(h->Wst, W_dev, sizeof(double2) * n2, cudaMemcpyDeviceToDevice, st));
        c->W = h->Wst;
    }
    QF_CUDA(cudaMemsetAsync(h->dW, 0, sizeof(double2) * n2 * B, st));       // :430
    QF_CUDA(cudaMemcpyAsync(h->Wh, c->W, sizeof(double2) * n2 * B, cudaMemcpyDeviceToDevice, st));   // W~ = W + 0
    if (c->compsum) QF_CUDA(cudaMemsetAsync(h->kahan_c, 0, sizeof(double2) * n2 * B, st));   // :457
    QF_CHECK(qf_launch_norm_inf(h, c->W, st));
    k_call_begin<<<B, 1, 0, st>>>(h->ctrl, tol, tol_factor, multistate ? 1 : 0);
    h->launches++;
    QF_CUDA(cudaGetLastError());
    if (c->xmode) QF_CHECK(qf_xchg_skew_check(h, c->W, st));
    return QF_OK;
}

// Call end, tile exchange only: every rank holds the tile pairs it owns; complete the state everywhere (phase 1: push +
// signal, phase 2: wait) and hand it back to the caller's buffer.
static int isomp_call_gather(qf_handle_s *h, const IsompCall &c, void *W_dev, cudaStream_t st, int phases)
{
    if (!c.xmode) return QF_OK;
    if (phases & 1) {
        QF_CHECK(qf_xchg_push_state(h, st));
        if (phases != 3) QF_CHECK(qf_xchg_signal(h, QF_XF_X, false, st));
    }
    if (phases & 2) {
        QF_CHECK(phases == 3 ? qf_xchg_barrier(h, QF_XF_X, false, st) : qf_xchg_wait(h, QF_XF_X, false, st));
        if (W_dev) QF_CUDA(cudaMemcpyAsync(W_dev, h->Wst, sizeof(double2) * h->mat_elems, cudaMemcpyDeviceToDevice, st));
    }
    return QF_OK;
}

// Statistics of the call (isospectral.py:607-611); synchronises the stream.
static int isomp_call_collect(qf_handle_s *h, int steps, qf_stats *stats, int32_t *iters_per_step, cudaStream_t st)
{
    const int B = h->batch;
    QF_CUDA(cudaMemcpyAsync(h->ctrl_host, h->ctrl, sizeof(QfCtrl) * B, cudaMemcpyDeviceToHost, st));
    if (iters_per_step && steps > 0) {
        QF_CUDA(cudaMemcpy2DAsync(iters_per_step, sizeof(int32_t) * steps, h->iters_dev, sizeof(int32_t) * h->steps_cap,
                                  sizeof(int32_t) * steps, B, cudaMemcpyDeviceToHost, st));
    }
    QF_CUDA(cudaStreamSynchronize(st));
    int rc = QF_OK;
    for (int b = 0; b < B; ++b) {
        const QfCtrl &c = h->ctrl_host[b];
        if (stats) {
            stats[b].tol_used = c.tol;
            stats[b].last_resnorm = c.resnorm;
            stats[b].total_iterations = c.total_it;
            stats[b].number_of_maxit = c.n_maxit;
            stats[b].nonfinite = c.nonfinite;
            stats[b].steps_done = c.steps_done;
        }
        if (c.nonfinite == 2) {
            qf_set_error("a peer rank did not answer within the time limit of the tile exchange (member %d, step %d)", b, c.steps_done);
            rc = QF_ERR_COMM;
        } else if (c.nonfinite && rc == QF_OK) {
            qf_set_error("array must not contain infs or NaNs (member %d, step %d)", b, c.steps_done);
            rc = QF_ERR_NONFINITE;
        }
    }
    return rc;
}

// W_dev may be NULL on the tile-exchange path only (state already staged in the handle, qf_isomp_host).
int qf_isomp_impl(qf_handle_s *h, void *W_dev, double dt, int steps, double tol, int maxit, int minit, unsigned flags,
                  qf_stats *stats, int32_t *iters_per_step, cudaStream_t st)
{
    IsompCall c;
    QF_CHECK(isomp_call_begin(h, W_dev, dt, steps, tol, maxit, minit, flags, st, &c));
    const int B = h->batch;
    const size_t n2 = h->mat_elems;
    const dim3 gz((unsigned)std::min<size_t>((n2 + 255) / 256, (size_t)h->sm_count * 8), B);
    QfStepGraph *sg = nullptr;
    if (h->use_graph && steps > 0) {
        // prepare everything that allocates (tile lists) before capturing
        const bool real_comm = (h->nranks > 1 && h->comm_mode != 0);
        QF_CHECK(qf_gemm_prepare(h, real_comm ? h->rank : -1, h->nranks));
        int rc = build_step_graph(h, c.W, c.eps, maxit, minit, c.compsum, c.reinit, &sg);
        if (rc != QF_OK) {
            if (!h->graph_warned) fprintf(stderr, "quflow_b200: step graph unavailable (%s); using eager launches\n", qf_last_error());
            h->graph_warned = 1;
            h->use_graph = 0;
            sg = nullptr;
            cudaGetLastError();   // clear the error state left by the failed instantiation
        }
    }
    if (sg) {
        for (int k = 0; k < steps; ++k) QF_CUDA(cudaGraphLaunch(sg->exec, st));
    } else {
        for (int k = 0; k < steps; ++k) {
            k_step_begin<<<1, 1, 0, st>>>(h->ctrl, B, 0, 0);
            h->launches++;
            if (c.reinit) {
                k_zero<<<gz, 256, 0, st>>>(h->dW, n2, h->ctrl);
                h->launches++;
            }
            for (int i = 0; i < maxit; ++i) QF_CHECK(qf_enqueue_iteration(h, c.W, c.eps, maxit, minit, st, nullptr));
            QF_CHECK(qf_enqueue_update(h, c.W, c.compsum, c.reinit, st));
        }
    }
    QF_CHECK(isomp_call_gather(h, c, W_dev, st, 3));
    const int rc = isomp_call_collect(h, steps, stats, iters_per_step, st);
    if (sg) {
        long long max_it = 0;
        for (int b = 0; b < B; ++b) max_it = std::max(max_it, (long long)h->ctrl_host[b].total_it);
        // kernels actually executed by the graphs: per step begin/update (+zero), per executed loop pass the body
        h->launches += (long long)steps * sg->kernels_per_step + max_it * sg->kernels_per_iter;
    }
    return rc;
}

extern "C" int qf_isomp(qf_handle_t h, void *W_dev, double dt, int steps, double tol, int maxit, int minit, unsigned flags,
                        qf_stats *stats, int32_t *iters_per_step, void *stream)
{
    if (!h || !W_dev) { qf_set_error("qf_isomp: null handle or pointer"); return QF_ERR_INVALID; }
    QF_ON_DEVICE(h->device);
    return qf_isomp_impl(h, W_dev, dt, steps, tol, maxit, minit, flags, stats, iters_per_step, (cudaStream_t)stream);
}

// Test driver: G handles on ONE device attached to each other by qf_comm_attach_local run the tile-exchange path in lock
// step — every phase is enqueued for all ranks before the next phase of any rank, on one stream, with eager launches — so
// the ownership, push and flag logic of the multi-GPU path can be checked on a single GPU without kernels that wait for
// kernels queued behind them (B200_PROFILING.md).  Same arguments as qf_isomp, one W_dev / stats entry per rank.
extern "C" int qf_isomp_lockstep(qf_handle_t *hs, int G, void **W_devs, double dt, int steps, double tol, int maxit, int minit,
                                 unsigned flags, qf_stats *stats, int32_t *iters_per_step, void *stream)
{
    if (!hs || !W_devs || G < 1 || G > QF_MAX_RANKS) { qf_set_error("qf_isomp_lockstep: bad arguments"); return QF_ERR_INVALID; }
    for (int r = 0; r < G; ++r)
        if (!hs[r] || !W_devs[r] || hs[r]->device != hs[0]->device || hs[r]->N != hs[0]->N || hs[r]->batch != 1 ||
            (G > 1 && (hs[r]->comm_mode != 5 || hs[r]->nranks != G || hs[r]->rank != r))) {
            qf_set_error("qf_isomp_lockstep: handle %d is not rank %d of a local tile-exchange group of %d", r, r, G);
            return QF_ERR_INVALID;
        }
    QF_ON_DEVICE(hs[0]->device);
    cudaStream_t st = (cudaStream_t)stream;
    IsompCall c[QF_MAX_RANKS];
    const size_t n2 = hs[0]->mat_elems;
    const dim3 gz((unsigned)std::min<size_t>((n2 + 255) / 256, (size_t)hs[0]->sm_count * 8), 1);
    for (int r = 0; r < G; ++r) {
        QF_CHECK(isomp_call_begin(hs[r], W_devs[r], dt, steps, tol, maxit, minit, flags, st, &c[r]));
        QF_CHECK(qf_gemm_prepare(hs[r], G > 1 ? r : -1, G));
    }
    for (int k = 0; k < steps; ++k) {
        for (int r = 0; r < G; ++r) {
            k_step_begin<<<1, 1, 0, st>>>(hs[r]->ctrl, 1, 0, 0);
            if (c[r].reinit) k_zero<<<gz, 256, 0, st>>>(hs[r]->dW, n2, hs[r]->ctrl);
        }
        for (int i = 0; i < maxit; ++i)
            for (int ph : {QF_PH_A, QF_PH_B, QF_PH_C})
                for (int r = 0; r < G; ++r) QF_CHECK(qf_enqueue_iteration(hs[r], c[r].W, c[r].eps, maxit, minit, st, nullptr, ph));
        for (int ph : {1, 2})
            for (int r = 0; r < G; ++r) QF_CHECK(qf_enqueue_update(hs[r], c[r].W, c[r].compsum, c[r].reinit, st, ph));
    }
    for (int ph : {1, 2})
        for (int r = 0; r < G; ++r) QF_CHECK(isomp_call_gather(hs[r], c[r], W_devs[r], st, ph));
    int rc = QF_OK;
    for (int r = 0; r < G; ++r) {
        const int rr = isomp_call_collect(hs[r], steps, stats ? stats + r : nullptr, iters_per_step ? iters_per_step + (size_t)r * steps : nullptr, st);
        if (rr != QF_OK) rc = rr;
    }
    return rc;
}

// ------------------------------------------------------------------------------- host-stepped driver
// The same kernels as qf_isomp, driven one fixed-point iteration at a time by the host, for callers whose step
// contains host code: callback (isospectral.py:550-551), forcing (:403-414, :511-520, :594-596), strang_splitting
// (:466-467, :602-603), custom or time-dependent Hamiltonians (:416-424, :488-491).
static int step_check(qf_handle_s *h, const char *who)
{
    if (!h) { qf_set_error("%s: null handle", who); return QF_ERR_INVALID; }
    if (h->batch != 1) { qf_set_error("%s: the host-stepped driver advances one member", who); return QF_ERR_UNSUPPORTED; }
    if (h->nranks > 1 && !(h->comm_mode == 1 || h->comm_mode == 2 || h->comm_mode == 5)) {
        qf_set_error("%s: the handle is set up for emulated ranks", who);
        return QF_ERR_UNSUPPORTED;
    }
    if (!h->step_open) { qf_set_error("%s: qf_step_open was not called", who); return QF_ERR_INVALID; }
    return QF_OK;
}

// Several GPUs in the host-stepped mode: the hooks run host code on EVERY rank and must see complete matrices there, so
// this mode uses the all-gather data path whatever the handle's default is — the two GEMMs are sharded by row blocks
// (rank-permuted output rows, qf_prow), A and S are completed on every rank by the pull kernels (or NCCL), and the tail,
// the stopping rule and the update run replicated on identical bytes.
struct StepLayout { int rank, G, hb; bool nccl; };
static StepLayout step_layout(const qf_handle_s *h)
{
    StepLayout l;
    l.G = h->nranks;
    l.rank = l.G > 1 ? h->rank : -1;
    l.hb = l.G > 1 ? qf_block_rows(h->N, l.G) : h->N;
    l.nccl = l.G > 1 && h->comm_mode == 1;
    return l;
}

extern "C" int qf_step_open(qf_handle_t h, const void *W_dev, double dt, double tol, unsigned flags, double *tol_used, void *stream)
{
    if (!h || !W_dev) { qf_set_error("qf_step_open: null argument"); return QF_ERR_INVALID; }
    if (h->batch != 1) { qf_set_error("qf_step_open: the host-stepped driver advances one member"); return QF_ERR_UNSUPPORTED; }
    if (h->nranks > 1 && !(h->comm_mode == 1 || h->comm_mode == 2 || h->comm_mode == 5)) { qf_set_error("qf_step_open: the handle is set up for emulated ranks"); return QF_ERR_UNSUPPORTED; }
    if (h->nranks > 1) QF_CHECK(qf_gemm_prepare_gather(h));
    QF_ON_DEVICE(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n2 = h->mat_elems;
    const bool compsum = (flags & QF_FLAG_COMPSUM) != 0;
    if (compsum && !h->kahan_c) QF_CUDA(cudaMalloc(&h->kahan_c, sizeof(double2) * n2));
    const double hb = qf_hbar(h->N);
    h->step_eps = dt / (2.0 * hb);                                          // isospectral.py:436-437
    h->step_flags = flags;
    double mach_eps = 2.220446049250313e-16;
    if (!compsum) mach_eps = sqrt(mach_eps);                                // :441-443
    QF_CUDA(cudaMemsetAsync(h->dW, 0, sizeof(double2) * n2, st));           // :430
    if (compsum) QF_CUDA(cudaMemsetAsync(h->kahan_c, 0, sizeof(double2) * n2, st));
    QF_CHECK(qf_launch_norm_inf(h, (const double2 *)W_dev, st));
    k_call_begin<<<1, 1, 0, st>>>(h->ctrl, tol, mach_eps * dt / hb, 0);     // :440-448
    h->launches++;
    QF_CUDA(cudaMemcpyAsync(h->ctrl_host, h->ctrl, sizeof(QfCtrl), cudaMemcpyDeviceToHost, st));
    QF_CUDA(cudaStreamSynchronize(st));
    if (tol_used) *tol_used = h->ctrl_host[0].tol;
    h->step_open = 1;
    return QF_OK;
}

extern "C" int qf_step_begin(qf_handle_t h, const void *W_dev, void *stream)
{
    QF_CHECK(step_check(h, "qf_step_begin"));
    if (!W_dev) { qf_set_error("qf_step_begin: null W"); return QF_ERR_INVALID; }
    QF_ON_DEVICE(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n2 = h->mat_elems;
    k_step_begin<<<1, 1, 0, st>>>(h->ctrl, 1, 0, 0);                        // resnorm = inf (:470)
    h->launches++;
    if (h->step_flags & QF_FLAG_REINITIALIZE) QF_CUDA(cudaMemsetAsync(h->dW, 0, sizeof(double2) * n2, st));   // :471-472
    QF_CHECK(qf_launch_whalf(h, (const double2 *)W_dev, h->dW, h->Wh, st));   // W~ = W + dW (:481-482); W may have been changed by the host
    return QF_OK;
}

extern "C" void *qf_step_buffer(qf_handle_t h, int which)
{
    if (!h) return nullptr;
    switch (which) {
        case QF_BUF_WHALF: return h->Wh;
        case QF_BUF_P: return h->P;
        case QF_BUF_SCRATCH: return h->scratch;
        default: return nullptr;
    }
}

extern "C" int qf_step_hamiltonian(qf_handle_t h, void *stream)
{
    QF_CHECK(step_check(h, "qf_step_hamiltonian"));
    QF_ON_DEVICE(h->device);
    return qf_launch_poisson(h, h->Wh, nullptr, h->Wh, h->P, h->step_eps, true, (cudaStream_t)stream);   // :489,:492
}

extern "C" int qf_step_scale_p(qf_handle_t h, int divide, void *stream)
{
    QF_CHECK(step_check(h, "qf_step_scale_p"));
    QF_ON_DEVICE(h->device);
    const size_t n2 = h->mat_elems;
    k_scale<<<(unsigned)std::min<size_t>((n2 + 255) / 256, (size_t)h->sm_count * 8), 256, 0, (cudaStream_t)stream>>>(h->P, n2, h->step_eps, divide);
    h->launches++;
    QF_CUDA(cudaGetLastError());
    return QF_OK;
}

extern "C" int qf_step_products(qf_handle_t h, void *stream)
{
    QF_CHECK(step_check(h, "qf_step_products"));
    QF_ON_DEVICE(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const StepLayout l = step_layout(h);
    QF_CHECK(qf_launch_zgemm(h, h->P, h->Wh, h->A, false, true, l.rank, l.G, false, st));       // :496 (own row blocks)
    if (l.G > 1) QF_CHECK(l.nccl ? qf_comm_allgather_rows(h, h->A, st) : qf_comm_p2p_allgather(h, 0, true, st));
    QF_CHECK(qf_launch_zgemm(h, h->A, h->P, h->S, true, true, l.rank, l.G, l.G > 1, st));        // :499
    if (l.G > 1) QF_CHECK(l.nccl ? qf_comm_allgather_rows(h, h->S, st) : qf_comm_p2p_allgather(h, 1, true, st));
    return QF_OK;
}

extern "C" int qf_step_close_iteration(qf_handle_t h, const void *W_dev, const void *F_dev, double fscale, int maxit, int minit,
                                       int *active, double *resnorm, void *stream)
{
    QF_CHECK(step_check(h, "qf_step_close_iteration"));
    if (!W_dev) { qf_set_error("qf_step_close_iteration: null W"); return QF_ERR_INVALID; }
    QF_ON_DEVICE(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int N = h->N;
    const int nb = (N + TS - 1) / TS;
    dim3 g(nb, nb, 1);
    const StepLayout l = step_layout(h);
    if (F_dev)
        k_post<true><<<g, 256, 0, st>>>(h->A, h->S, h->dW, h->rowpart, N, h->nslots, h->ctrl, l.hb, l.G, (const double2 *)W_dev, h->Wh,
                                        (const double2 *)F_dev, fscale, QfXchg());
    else
        k_post<false><<<g, 256, 0, st>>>(h->A, h->S, h->dW, h->rowpart, N, h->nslots, h->ctrl, l.hb, l.G, (const double2 *)W_dev, h->Wh, nullptr, 0.0,
                                         QfXchg());
    k_control<<<dim3((N + 7) / 8, 1), 256, 0, st>>>(h->rowpart, h->nslots, h->rowpart + (size_t)h->nslots * N, h->nslots, N, h->ctrl,
                                                    maxit, minit, 0, 0, 0);
    h->launches += 2;
    QF_CUDA(cudaGetLastError());
    QF_CUDA(cudaMemcpyAsync(h->ctrl_host, h->ctrl, sizeof(QfCtrl), cudaMemcpyDeviceToHost, st));
    QF_CUDA(cudaStreamSynchronize(st));
    const QfCtrl &c = h->ctrl_host[0];
    if (active) *active = c.active;
    if (resnorm) *resnorm = c.resnorm;
    if (c.nonfinite) {
        qf_set_error("array must not contain infs or NaNs");
        return QF_ERR_NONFINITE;
    }
    return QF_OK;
}

extern "C" int qf_step_increment(qf_handle_t h, void *out_dev, void *stream)
{
    QF_CHECK(step_check(h, "qf_step_increment"));
    if (!out_dev) { qf_set_error("qf_step_increment: null output"); return QF_ERR_INVALID; }
    QF_ON_DEVICE(h->device);
    const int N = h->N;
    const int nb = (N + TS - 1) / TS;
    const StepLayout l = step_layout(h);
    k_increment<<<dim3(nb, nb, 1), 256, 0, (cudaStream_t)stream>>>(h->A, (double2 *)out_dev, N, l.hb, l.G);
    h->launches++;
    QF_CUDA(cudaGetLastError());
    return QF_OK;
}

extern "C" int qf_step_update(qf_handle_t h, void *W_dev, const void *F_dev, double fscale, void *stream)
{
    QF_CHECK(step_check(h, "qf_step_update"));
    if (!W_dev) { qf_set_error("qf_step_update: null W"); return QF_ERR_INVALID; }
    const bool compsum = (h->step_flags & QF_FLAG_COMPSUM) != 0;
    if (compsum && F_dev) { qf_set_error("Compensated sum with forcing is not yet implemented."); return QF_ERR_UNSUPPORTED; }   // :588-589
    QF_ON_DEVICE(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int N = h->N;
    const int nb = (N + TS - 1) / TS;
    dim3 g(nb, nb, 1);
    const int reinit = (h->step_flags & QF_FLAG_REINITIALIZE) ? 1 : 0;
    double2 *W = (double2 *)W_dev;
    const StepLayout l = step_layout(h);
    if (compsum)
        k_update<true, false><<<g, 256, 0, st>>>(h->A, W, h->kahan_c, N, h->ctrl, nullptr, 0, l.hb, l.G, h->dW, h->Wh, reinit, nullptr, 0.0, QfXchg());
    else if (F_dev)
        k_update<false, true><<<g, 256, 0, st>>>(h->A, W, nullptr, N, h->ctrl, nullptr, 0, l.hb, l.G, h->dW, h->Wh, reinit, (const double2 *)F_dev, fscale, QfXchg());
    else
        k_update<false, false><<<g, 256, 0, st>>>(h->A, W, nullptr, N, h->ctrl, nullptr, 0, l.hb, l.G, h->dW, h->Wh, reinit, nullptr, 0.0, QfXchg());
    h->launches++;
    QF_CUDA(cudaGetLastError());
    return QF_OK;
}

extern "C" int qf_step_stats(qf_handle_t h, qf_stats *stats, void *stream)
{
    QF_CHECK(step_check(h, "qf_step_stats"));
    if (!stats) { qf_set_error("qf_step_stats: null output"); return QF_ERR_INVALID; }
    QF_ON_DEVICE(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    QF_CUDA(cudaMemcpyAsync(h->ctrl_host, h->ctrl, sizeof(QfCtrl), cudaMemcpyDeviceToHost, st));
    QF_CUDA(cudaStreamSynchronize(st));
    const QfCtrl &c = h->ctrl_host[0];
    stats->tol_used = c.tol;
    stats->last_resnorm = c.resnorm;
    stats->total_iterations = c.total_it;
    stats->number_of_maxit = c.n_maxit;
    stats->nonfinite = c.nonfinite;
    stats->steps_done = c.steps_done;
    return QF_OK;
}

extern "C" int qf_profile_iteration(qf_handle_t h, const void *W_dev, double dt, int reps, qf_phase_times *out, void *stream)
{
    if (!h || !W_dev || !out || reps < 1) { qf_set_error("qf_profile_iteration: bad arguments"); return QF_ERR_INVALID; }
    QF_ON_DEVICE(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int N = h->N, B = h->batch;
    const size_t n2 = h->mat_elems;
    const double eps = dt / (2.0 * qf_hbar(N));
    cudaEvent_t ev[12];
    for (auto &e : ev) QF_CUDA(cudaEventCreate(&e));
    const bool xsplit = h->nranks > 1 && h->comm_mode == 5 && !(h->fuse_post && qf_gemm_can_fuse_post(h));
    float xacc[5] = {0, 0, 0, 0, 0};
    if (!h->io) QF_CUDA(cudaMalloc(&h->io, sizeof(double2) * n2 * B));
    QF_CUDA(cudaMemcpyAsync(h->io, W_dev, sizeof(double2) * n2 * B, cudaMemcpyDeviceToDevice, st));
    QF_CUDA(cudaMemsetAsync(h->dW, 0, sizeof(double2) * n2 * B, st));
    QF_CUDA(cudaMemcpyAsync(h->Wh, h->io, sizeof(double2) * n2 * B, cudaMemcpyDeviceToDevice, st));
    QF_CHECK(qf_launch_norm_inf(h, h->io, st));
    k_call_begin<<<B, 1, 0, st>>>(h->ctrl, 0.0, 0.0, 0);   // tol = 0: never converges by tolerance
    float acc[5] = {0, 0, 0, 0, 0};
    for (int r = -1; r < reps; ++r) {   // r = -1: warm-up
        k_step_begin<<<1, 1, 0, st>>>(h->ctrl, B, 0, 0);
        QF_CHECK(qf_enqueue_iteration(h, h->io, eps, 1 << 30, 1 << 30, st, ev));
        QF_CUDA(cudaEventRecord(ev[5], st));
        QF_CHECK(qf_enqueue_update(h, h->io, false, false, st));
        QF_CUDA(cudaEventRecord(ev[6], st));
        QF_CUDA(cudaStreamSynchronize(st));
        if (r < 0) continue;
        float ms;
        for (int p = 0; p < 4; ++p) {
            QF_CUDA(cudaEventElapsedTime(&ms, ev[p], ev[p + 1]));
            acc[p] += ms;
        }
        QF_CUDA(cudaEventElapsedTime(&ms, ev[5], ev[6]));
        acc[4] += ms;
        if (xsplit) {
            const int order[5] = {3, 7, 8, 11, 4};         // GEMM 2 done | tail kernel | W~ push | signal + wait | control
            for (int p = 0; p < 4; ++p) {
                QF_CUDA(cudaEventElapsedTime(&ms, ev[order[p]], ev[order[p + 1]]));
                xacc[p == 3 ? 4 : p] += ms;
            }
            if (qf_xchg_desc(h)->upper_only) {
                QF_CUDA(cudaEventElapsedTime(&ms, ev[9], ev[10]));   // the local mirror at the start of the iteration (serial here, forked in the graph)
                xacc[3] += ms;
                acc[0] -= ms;                                        // ... which the event pair of the Poisson phase spans as well
            }
        }
    }
    out->poisson_ms = acc[0] / reps;
    out->gemm1_ms = acc[1] / reps;
    out->gemm2_ms = acc[2] / reps;
    out->post_ms = acc[3] / reps;
    out->update_ms = acc[4] / reps;
    out->x_tail_ms = xacc[0] / reps;
    out->x_push_ms = xacc[1] / reps;
    out->x_wait_ms = xacc[2] / reps;
    out->x_mirror_ms = xacc[3] / reps;
    out->x_control_ms = xacc[4] / reps;
    for (auto &e : ev) cudaEventDestroy(e);
    return QF_OK;
}
