// Hoppe–Yau Laplacian on the device: P = eps * Delta_N^{-1} (W + dW), and W = Delta_N P.
//
// Replaces quflow/laplacian/cpu.py: `_compute_cpu_laplacian` (:55-95), `_solve_cpu_skewh`
// (:281-362), `solve_poisson` (:681-734), `laplace`/`_dot_cpu_generic` (:628-669, :98-108).
//
// Math.  Element (k, k+m) of the upper triangle is position k of the tridiagonal system of
// diagonal m (length N-m):  o_k x_{k-1} + d_k x_k + o_{k+1} x_{k+1} = r_k  with
//   d_k = -((N-1)(2k+1+m) - 2k(k+m)),   o_k = sqrt((k+m)(N-k-m) k (N-k)),   d_0 -= 1/2 for m = 0.
// The matrices do not depend on W, so their LU factors  w_k = o_k/u_{k-1}, u_k = d_k - w_k o_k
// are built once per N on the host (the reference recomputes them every call) and kept in HBM in
// the same [k][k+m] layout as the matrix, so a warp that walks "one system per lane" reads whole
// rows: fully coalesced.
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#include "qf_common.cuh"

// ---------------------------------------------------------------------------------------
// host: coefficient / factor tables
// ---------------------------------------------------------------------------------------
int qf_build_tables(qf_handle_s *h)
{
    const int N = h->N;
    const size_t n2 = (size_t)N * N;
    std::vector<double> tw(n2, 0.0), tiu(n2, 0.0), to(n2, 0.0);
    const double dN = (double)N;
    for (int m = 0; m < N; ++m) {
        double u_prev = 0.0;
        for (int k = 0; k < N - m; ++k) {
            const size_t idx = (size_t)k * N + (k + m);
            const double dk = (double)k, dm = (double)m;
            double d = -((dN - 1.0) * (2.0 * dk + 1.0 + dm) - 2.0 * dk * (dk + dm));   // cpu.py:82
            const double o = sqrt(((dk + dm) * (dN - dk - dm)) * (dk * (dN - dk)));    // cpu.py:83
            if (m == 0 && k == 0) d -= 0.5;                                             // cpu.py:90
            double w = 0.0, u = d;
            if (k > 0) {
                w = o / u_prev;                                                         // cpu.py:324
                u = d - w * o;                                                          // cpu.py:325
            }
            tw[idx] = w;
            tiu[idx] = 1.0 / u;
            to[idx] = o;
            u_prev = u;
        }
    }
    QF_CUDA(cudaMalloc(&h->tab_w, n2 * sizeof(double)));
    QF_CUDA(cudaMalloc(&h->tab_iu, n2 * sizeof(double)));
    QF_CUDA(cudaMalloc(&h->tab_o, n2 * sizeof(double)));
    QF_CUDA(cudaMemcpy(h->tab_w, tw.data(), n2 * sizeof(double), cudaMemcpyHostToDevice));
    QF_CUDA(cudaMemcpy(h->tab_iu, tiu.data(), n2 * sizeof(double), cudaMemcpyHostToDevice));
    QF_CUDA(cudaMemcpy(h->tab_o, to.data(), n2 * sizeof(double), cudaMemcpyHostToDevice));
    return QF_OK;
}

// ---------------------------------------------------------------------------------------
// kernels (v1: one thread per diagonal, row-coalesced; see DESIGN.md for the roadmap)
// ---------------------------------------------------------------------------------------
// Wh = W + dW over the full matrix (GEMM 1 needs all of W~), fused with the trace of W~.
__global__ void k_whalf(const double2 *__restrict__ W, const double2 *__restrict__ dW, double2 *__restrict__ Wh,
                        size_t n2, const QfCtrl *__restrict__ ctrl, int gated)
{
    const int b = blockIdx.y;
    if (gated && !ctrl[b].active) return;
    const size_t off = (size_t)b * n2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
        double2 w = W[off + i];
        if (dW) w = zadd(w, dW[off + i]);
        Wh[off + i] = w;
    }
}

// mean of the diagonal of Wh -> ctrl[b].trW (complex kept in trW / trW_im)
__global__ void k_trace(const double2 *__restrict__ Wh, int N, QfCtrl *ctrl, double2 *trbuf, int gated)
{
    const int b = blockIdx.x;
    if (gated && !ctrl[b].active) return;
    __shared__ double sre[32], sim[32];
    const double2 *M = Wh + (size_t)b * N * N;
    double re = 0.0, im = 0.0;
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        double2 v = M[(size_t)k * N + k];
        re += v.x;
        im += v.y;
    }
    for (int o = 16; o > 0; o >>= 1) {
        re += __shfl_xor_sync(0xffffffffu, re, o);
        im += __shfl_xor_sync(0xffffffffu, im, o);
    }
    if ((threadIdx.x & 31) == 0) { sre[threadIdx.x >> 5] = re; sim[threadIdx.x >> 5] = im; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = 0.0, i = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { r += sre[w]; i += sim[w]; }
        trbuf[b] = make_double2(r / N, i / N);
    }
}

__global__ void k_thomas(const double2 *__restrict__ Wh, double2 *__restrict__ P, double2 *__restrict__ scratch,
                         const double *__restrict__ tw, const double *__restrict__ tiu, const double *__restrict__ to,
                         int N, double eps, const QfCtrl *__restrict__ ctrl, const double2 *__restrict__ trbuf, int gated)
{
    const int b = blockIdx.y;
    if (gated && !ctrl[b].active) return;
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= N) return;
    const size_t off = (size_t)b * N * N;
    const double2 *R = Wh + off;
    double2 *X = P + off;
    double2 *C = scratch + off;
    const int n = N - m;
    double2 tr = make_double2(0.0, 0.0);
    if (m == 0) tr = trbuf[b];
    // forward sweep  c_k = r_k - w_k c_{k-1}
    double2 c = make_double2(0.0, 0.0);
#pragma unroll 4
    for (int k = 0; k < n; ++k) {
        const size_t idx = (size_t)k * N + (k + m);
        double2 r = R[idx];
        const double w = tw[idx];
        r.x -= tr.x;
        r.y -= tr.y;
        c.x = r.x - w * c.x;
        c.y = r.y - w * c.y;
        C[idx] = c;
    }
    // backward sweep  x_k = (c_k - o_{k+1} x_{k+1}) / u_k
    double2 x = make_double2(0.0, 0.0);
    double o_next = 0.0;
#pragma unroll 4
    for (int k = n - 1; k >= 0; --k) {
        const size_t idx = (size_t)k * N + (k + m);
        const double2 ck = C[idx];
        const double iu = tiu[idx];
        x.x = (ck.x - o_next * x.x) * iu;
        x.y = (ck.y - o_next * x.y) * iu;
        o_next = to[idx];
        if (m == 0) {
            X[idx] = x;   // unscaled: the trace of P is removed (and eps applied) by k_fix_trace
        } else {
            X[idx] = make_double2(eps * x.x, eps * x.y);
            X[(size_t)(k + m) * N + k] = make_double2(-eps * x.x, eps * x.y);   // P[j,i] = -conj(P[i,j]), cpu.py:334,340
        }
    }
}

// P_kk <- eps * (P_kk - mean(diag P))   cpu.py:342-352 followed by isospectral.py:492
__global__ void k_fix_trace(double2 *P, int N, double eps, const QfCtrl *__restrict__ ctrl, int gated)
{
    const int b = blockIdx.x;
    if (gated && !ctrl[b].active) return;
    __shared__ double sre[32], sim[32];
    __shared__ double2 mean;
    double2 *M = P + (size_t)b * N * N;
    double re = 0.0, im = 0.0;
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        double2 v = M[(size_t)k * N + k];
        re += v.x;
        im += v.y;
    }
    for (int o = 16; o > 0; o >>= 1) {
        re += __shfl_xor_sync(0xffffffffu, re, o);
        im += __shfl_xor_sync(0xffffffffu, im, o);
    }
    if ((threadIdx.x & 31) == 0) { sre[threadIdx.x >> 5] = re; sim[threadIdx.x >> 5] = im; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = 0.0, i = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { r += sre[w]; i += sim[w]; }
        mean = make_double2(r / N, i / N);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        double2 v = M[(size_t)k * N + k];
        M[(size_t)k * N + k] = make_double2(eps * (v.x - mean.x), eps * (v.y - mean.y));
    }
}

// ---------------------------------------------------------------------------------------
// v2: chunked parallel solve.  One CTA owns GS = 4 adjacent diagonals m0..m0+3; thread (s, c) owns
// positions [cL, (c+1)L) of system m0+s in registers.  With the LDL^T factors (w_k, 1/u_k) both sweeps
// are first-order linear recurrences,
//     forward   c_k = r_k - w_k c_{k-1}                 backward  x_k = c_k/u_k - w_{k+1} x_{k+1},
// so each chunk is an affine map of its carry-in.  Pass 1 evaluates the chunk with carry-in 0 and the
// map's slope (a running product), a warp-shuffle + shared-memory scan composes the maps across the
// chunks of a system, pass 2 re-runs the chunk from its true carry-in.  Lanes s = 0..3 of a quad touch
// 64 contiguous bytes of a matrix row, every thread has L independent 16-byte loads in flight, and the
// per-CTA sequential depth is 4L FMAs + two log-depth scans instead of 2N.
// HBM traffic: upper triangle of W~ (8 N^2 B) + w and 1/u tables (8 N^2 B) + full P (16 N^2 B) = 32 N^2 B.
// ---------------------------------------------------------------------------------------
template <int L, int GS, int MAXT>
__global__ void __launch_bounds__(MAXT, 512 / MAXT)
k_poisson_scan(const double2 *__restrict__ Wh, double2 *__restrict__ P, const double *__restrict__ tw,
               const double *__restrict__ tiu, int N, double eps, const QfCtrl *__restrict__ ctrl, int gated)
{
    const int b = blockIdx.y;
    if (gated && !ctrl[b].active) return;
    __shared__ double totA[16][GS], totBx[16][GS], totBy[16][GS];
    __shared__ double redx[16], redy[16];
    __shared__ double2 bcast;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    constexpr int GSH = (GS == 4) ? 2 : 1;
    const int s = tid & (GS - 1), c = tid >> GSH;
    const int m0 = blockIdx.x * GS;
    const int m = m0 + s;
    const int n = N - m;                 // length of this thread's system (<= 0: none)
    const int k0 = c * L;
    const size_t off = (size_t)b * N * N;
    const double2 *R = Wh + off;
    double2 *X = P + off;

    double2 r[L];
    double w[L + 1];                     // w[L] = w of the first position of the next chunk
    // Element (k, k+m) sits at flat index k (N+1) + m; its mirror (k+m, k) at k (N+1) + m N: both walk with stride N+1.
    // N <= 2048 so flat indices fit 32 bits.
    const unsigned stride = (unsigned)N + 1u;
    const unsigned e0 = (unsigned)k0 * stride + (unsigned)m;
    const int nvalid = min(L, max(0, n - k0));
    extern __shared__ double iu_s[];     // [L][blockDim.x]: 1/u of this thread's chunk, prefetched while the forward sweep runs
    const uint32_t iu_base = (uint32_t)__cvta_generic_to_shared(iu_s) + (uint32_t)tid * 8u;
    const uint32_t iu_pitch = blockDim.x * 8u;
    if (nvalid == L) {
        const double2 *rp = R + e0;
        const double *wp = tw + e0;
        const double *up = tiu + e0;
#pragma unroll
        for (int i = 0; i < L; ++i) {
            r[i] = rp[(size_t)i * stride];
            w[i] = __ldg(wp + (size_t)i * stride);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(iu_base + i * iu_pitch), "l"(up + (size_t)i * stride));
        }
        w[L] = (k0 + L < n) ? __ldg(wp + (size_t)L * stride) : 0.0;
    } else {
#pragma unroll
        for (int i = 0; i < L; ++i) {
            const bool ok = i < nvalid;
            const unsigned idx = ok ? e0 + (unsigned)i * stride : 0u;
            r[i] = ok ? R[idx] : make_double2(0.0, 0.0);
            w[i] = ok ? __ldg(tw + idx) : 0.0;
            const int sz = ok ? 8 : 0;       // src-size 0: zero fill
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(iu_base + i * iu_pitch), "l"(tiu + idx), "r"(sz));
        }
        w[L] = 0.0;
    }
    asm volatile("cp.async.commit_group;");

    // ---- m = 0: remove the mean of the diagonal from the right-hand side (cpu.py:311-317,327-328)
    if (m0 == 0) {
        double sx = 0.0, sy = 0.0;
        if (s == 0) {
#pragma unroll
            for (int i = 0; i < L; ++i) { sx += r[i].x; sy += r[i].y; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, o);
            sy += __shfl_xor_sync(0xffffffffu, sy, o);
        }
        if (lane == 0) { redx[warp] = sx; redy[warp] = sy; }
        __syncthreads();
        if (tid == 0) {
            double ax = 0.0, ay = 0.0;
            for (int q = 0; q < nwarps; ++q) { ax += redx[q]; ay += redy[q]; }
            bcast = make_double2(ax / N, ay / N);
        }
        __syncthreads();
        if (s == 0) {
            const double2 tr = bcast;
#pragma unroll
            for (int i = 0; i < L; ++i)
                if (k0 + i < n) { r[i].x -= tr.x; r[i].y -= tr.y; }
        }
        __syncthreads();
    }

    // ---- forward, pass 1: chunk map  c_out = A c_in + B
    double A = 1.0;
    double2 B = make_double2(0.0, 0.0);
#pragma unroll
    for (int i = 0; i < L; ++i) {
        B.x = r[i].x - w[i] * B.x;
        B.y = r[i].y - w[i] * B.y;
        A = -w[i] * A;
    }
    // inclusive scan over the 8 chunks of this warp (lanes with equal s are 4 apart)
#pragma unroll
    for (int d = GS; d < 32; d <<= 1) {
        const double eA = __shfl_up_sync(0xffffffffu, A, d);
        const double eBx = __shfl_up_sync(0xffffffffu, B.x, d);
        const double eBy = __shfl_up_sync(0xffffffffu, B.y, d);
        if (lane >= d) {
            B.x = A * eBx + B.x;
            B.y = A * eBy + B.y;
            A = A * eA;
        }
    }
    double xA = __shfl_up_sync(0xffffffffu, A, GS);
    double xBx = __shfl_up_sync(0xffffffffu, B.x, GS);
    double xBy = __shfl_up_sync(0xffffffffu, B.y, GS);
    if (lane < GS) { xA = 1.0; xBx = 0.0; xBy = 0.0; }
    if (lane >= 32 - GS) { totA[warp][s] = A; totBx[warp][s] = B.x; totBy[warp][s] = B.y; }
    __syncthreads();
    double2 carry;
    {
        double pBx = 0.0, pBy = 0.0;
        for (int q = 0; q < warp; ++q) {
            const double tA = totA[q][s];
            pBx = tA * pBx + totBx[q][s];
            pBy = tA * pBy + totBy[q][s];
        }
        carry.x = xA * pBx + xBx;
        carry.y = xA * pBy + xBy;
    }
    // ---- forward, pass 2 from the true carry-in; r becomes z = c / u
    asm volatile("cp.async.wait_group 0;" ::: "memory");   // this thread's own 1/u values (no cross-thread sharing)
#pragma unroll
    for (int i = 0; i < L; ++i) {
        carry.x = r[i].x - w[i] * carry.x;
        carry.y = r[i].y - w[i] * carry.y;
        const double iu = iu_s[i * blockDim.x + tid];
        r[i].x = carry.x * iu;
        r[i].y = carry.y * iu;
    }
    __syncthreads();   // tot* reused below

    // ---- backward, pass 1: x_out = A x_in + B  (x_in = value just after the chunk)
    A = 1.0;
    B = make_double2(0.0, 0.0);
#pragma unroll
    for (int i = L - 1; i >= 0; --i) {
        B.x = r[i].x - w[i + 1] * B.x;
        B.y = r[i].y - w[i + 1] * B.y;
        A = -w[i + 1] * A;
    }
#pragma unroll
    for (int d = GS; d < 32; d <<= 1) {
        const double eA = __shfl_down_sync(0xffffffffu, A, d);
        const double eBx = __shfl_down_sync(0xffffffffu, B.x, d);
        const double eBy = __shfl_down_sync(0xffffffffu, B.y, d);
        if (lane + d < 32) {
            B.x = A * eBx + B.x;
            B.y = A * eBy + B.y;
            A = A * eA;
        }
    }
    xA = __shfl_down_sync(0xffffffffu, A, GS);
    xBx = __shfl_down_sync(0xffffffffu, B.x, GS);
    xBy = __shfl_down_sync(0xffffffffu, B.y, GS);
    if (lane >= 32 - GS) { xA = 1.0; xBx = 0.0; xBy = 0.0; }
    if (lane < GS) { totA[warp][s] = A; totBx[warp][s] = B.x; totBy[warp][s] = B.y; }
    __syncthreads();
    {
        double pBx = 0.0, pBy = 0.0;
        for (int q = nwarps - 1; q > warp; --q) {
            const double tA = totA[q][s];
            pBx = tA * pBx + totBx[q][s];
            pBy = tA * pBy + totBy[q][s];
        }
        carry.x = xA * pBx + xBx;
        carry.y = xA * pBy + xBy;
    }
    // ---- backward, pass 2; r becomes x
#pragma unroll
    for (int i = L - 1; i >= 0; --i) {
        carry.x = r[i].x - w[i + 1] * carry.x;
        carry.y = r[i].y - w[i + 1] * carry.y;
        r[i] = carry;
    }

    // ---- m = 0: remove the mean of diag(P) (cpu.py:342-352)
    if (m0 == 0) {
        double sx = 0.0, sy = 0.0;
        if (s == 0) {
#pragma unroll
            for (int i = 0; i < L; ++i)
                if (k0 + i < n) { sx += r[i].x; sy += r[i].y; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, o);
            sy += __shfl_xor_sync(0xffffffffu, sy, o);
        }
        __syncthreads();
        if (lane == 0) { redx[warp] = sx; redy[warp] = sy; }
        __syncthreads();
        if (tid == 0) {
            double ax = 0.0, ay = 0.0;
            for (int q = 0; q < nwarps; ++q) { ax += redx[q]; ay += redy[q]; }
            bcast = make_double2(ax / N, ay / N);
        }
        __syncthreads();
        if (s == 0) {
            const double2 tr = bcast;
#pragma unroll
            for (int i = 0; i < L; ++i) { r[i].x -= tr.x; r[i].y -= tr.y; }
        }
    }

    // ---- store P = eps x and its skew-Hermitian mirror (cpu.py:334,340; isospectral.py:492)
    {
        double2 *xp = X + e0;
        double2 *xm = X + (unsigned)k0 * stride + (unsigned)m * (unsigned)N;
        if (nvalid == L && m != 0) {
#pragma unroll
            for (int i = 0; i < L; ++i) {
                const double2 v = make_double2(eps * r[i].x, eps * r[i].y);
                xp[(size_t)i * stride] = v;
                xm[(size_t)i * stride] = make_double2(-v.x, v.y);
            }
        } else {
#pragma unroll
            for (int i = 0; i < L; ++i) {
                if (i < nvalid) {
                    const double2 v = make_double2(eps * r[i].x, eps * r[i].y);
                    xp[(size_t)i * stride] = v;
                    if (m != 0) xm[(size_t)i * stride] = make_double2(-v.x, v.y);
                }
            }
        }
    }
}

// W = Delta P for a general matrix (cpu.py:98-108); coefficients recomputed on the fly.
__global__ void k_laplace(const double2 *__restrict__ P, double2 *__restrict__ W, int N)
{
    const int b = blockIdx.z;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= N) return;
    const double2 *Pm = P + (size_t)b * N * N;
    double2 *Wm = W + (size_t)b * N * N;
    const double dN = (double)N;
    const double m = fabs((double)(j - i));
    const double k = (double)min(i, j);
    const double d = -((dN - 1.0) * (2.0 * k + 1.0 + m) - 2.0 * k * (k + m));
    const size_t ij = (size_t)i * N + j;
    double2 p = Pm[ij];
    double2 w = make_double2(d * p.x, d * p.y);
    if (i < N - 1 && j < N - 1) {
        const double kp = k + 1.0;
        const double o = sqrt(((kp + m) * (dN - kp - m)) * (kp * (dN - kp)));
        double2 q = Pm[ij + N + 1];
        w.x += o * q.x;
        w.y += o * q.y;
    }
    if (i > 0 && j > 0) {
        const double o = sqrt(((k + m) * (dN - k - m)) * (k * (dN - k)));
        double2 q = Pm[ij - N - 1];
        w.x += o * q.x;
        w.y += o * q.y;
    }
    Wm[ij] = w;
}

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
int qf_launch_poisson(qf_handle_s *h, const double2 *W, const double2 *dW, double2 *Wh, double2 *P, double eps,
                      bool gated, cudaStream_t st)
{
    const int N = h->N;
    const size_t n2 = h->mat_elems;
    const int g = gated ? 1 : 0;
    if (Wh != W) {
        dim3 gw((unsigned)std::min<size_t>((n2 + 255) / 256, (size_t)h->sm_count * 8), h->batch);
        k_whalf<<<gw, 256, 0, st>>>(W, dW, Wh, n2, h->ctrl, g);
        h->launches++;
    }
    if (N <= 2048) {
        // chunked scan: 4 diagonals per CTA, 16 positions per thread (QF_POISSON_GS=2 selects 2 diagonals per CTA)
        const int chunks = (N + 15) / 16;
        static bool attr_done = false;
        if (!attr_done) {
            QF_CUDA(cudaFuncSetAttribute(k_poisson_scan<16, 4, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 512 * 16 * 8));
            QF_CUDA(cudaFuncSetAttribute(k_poisson_scan<16, 2, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 16 * 8));
            attr_done = true;
        }
        static int gs = 0;
        if (!gs) {
            const char *env = getenv("QF_POISSON_GS");
            gs = (env && env[0] == '2') ? 2 : 4;
        }
        if (gs == 4) {
            const int threads = ((4 * chunks + 31) / 32) * 32;
            dim3 grid((N + 3) / 4, h->batch);
            k_poisson_scan<16, 4, 512><<<grid, threads, threads * 16 * sizeof(double), st>>>(Wh, P, h->tab_w, h->tab_iu, N, eps, h->ctrl, g);
        } else {
            const int threads = ((2 * chunks + 31) / 32) * 32;
            dim3 grid((N + 1) / 2, h->batch);
            k_poisson_scan<16, 2, 256><<<grid, threads, threads * 16 * sizeof(double), st>>>(Wh, P, h->tab_w, h->tab_iu, N, eps, h->ctrl, g);
        }
        h->launches++;
    } else {
        // large-N fallback: one thread per diagonal
        k_trace<<<h->batch, 256, 0, st>>>(Wh, N, h->ctrl, h->trbuf, g);
        dim3 gt((N + 63) / 64, h->batch);
        k_thomas<<<gt, 64, 0, st>>>(Wh, P, h->scratch, h->tab_w, h->tab_iu, h->tab_o, N, eps, h->ctrl, h->trbuf, g);
        k_fix_trace<<<h->batch, 256, 0, st>>>(P, N, eps, h->ctrl, g);
        h->launches += 3;
    }
    QF_CUDA(cudaGetLastError());
    return QF_OK;
}

int qf_launch_laplace(qf_handle_s *h, const double2 *P, double2 *W, cudaStream_t st)
{
    const int N = h->N;
    dim3 g((N + 127) / 128, N, h->batch);
    k_laplace<<<g, 128, 0, st>>>(P, W, N);
    h->launches++;
    QF_CUDA(cudaGetLastError());
    return QF_OK;
}
