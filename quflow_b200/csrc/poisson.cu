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
    double2 *trbuf = h->trbuf;
    if (Wh != W) {
        dim3 g((unsigned)std::min<size_t>((n2 + 255) / 256, (size_t)h->sm_count * 8), h->batch);
        k_whalf<<<g, 256, 0, st>>>(W, dW, Wh, n2, h->ctrl, gated ? 1 : 0);
        h->launches++;
    }
    k_trace<<<h->batch, 256, 0, st>>>(Wh, N, h->ctrl, trbuf, gated ? 1 : 0);
    dim3 gt((N + 63) / 64, h->batch);
    k_thomas<<<gt, 64, 0, st>>>(Wh, P, h->scratch, h->tab_w, h->tab_iu, h->tab_o, N, eps, h->ctrl, trbuf, gated ? 1 : 0);
    k_fix_trace<<<h->batch, 256, 0, st>>>(P, N, eps, h->ctrl, gated ? 1 : 0);
    h->launches += 3;
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
