// Matrix <-> real spherical-harmonic coefficients on the device: the callers' data format on either side of the hot
// path.  Replaces quflow/quantization.py `mat2shr_parallel_` (:283-325) and
// `shr2mat_parallel_` (:172-227) for a basis already resident in HBM.
//
// The quantization basis is the reference's own array (quflow.quantization.get_basis(N), cached on disk by the
// reference): for every m = 0 .. N-1 a dense real (N-m) x (N-m) block, row-major [k][el - m], blocks concatenated
// (block m starts at basis_break_index(m, N), quantization.py:25-42).  Computing the basis (an eigenproblem per m) stays
// with the reference; this file only applies it:
//   mat2shr   omega[el^2 + el +- m] from  sum_k W[k+m, k] * B_m[k][el-m]         (one dot product per (m, el))
//   shr2mat   W[k, k+m], W[k+m, k]  from  sum_el B_m[k][el-m] * omega_complex(el, m)
// Both are HBM-bound streams over the basis (N^3/3 doubles: 0.36 GB at N = 512, 2.9 GB at N = 1024).
#include <math.h>

#include "qf_common.cuh"

namespace {

__host__ __device__ __forceinline__ long long basis_break(long long m, long long N)
{
    // quantization.py:38-41:  absm -= 1; ind = absm + 2 absm^2 - 6 absm N + 6 N^2; ind *= 1 + absm; return ind // 6
    const long long a = m - 1;
    return ((a + 2 * a * a - 6 * a * N + 6 * N * N) * (1 + a)) / 6;
}

// one thread per coefficient column el of block m; k runs sequentially (fixed summation order => deterministic)
__global__ void __launch_bounds__(128)
k_mat2shr(const double2 *__restrict__ W, const double *__restrict__ basis, double *__restrict__ omega, int N, int Nmax)
{
    const int m = blockIdx.y;
    const int el = m + blockIdx.x * 128 + threadIdx.x;
    if (blockIdx.x * 128 >= Nmax - m) return;
    const int n = N - m;                                    // length of the diagonal = block dimension
    const double *__restrict__ B = basis + basis_break(m, N) + (el - m);
    const bool ok = el < Nmax;
    __shared__ double2 diag[128];
    double ax = 0.0, ay = 0.0;
    for (int k0 = 0; k0 < n; k0 += 128) {
        const int k = k0 + threadIdx.x;
        __syncthreads();
        diag[threadIdx.x] = (k < n) ? W[(size_t)(k + m) * N + k] : make_double2(0.0, 0.0);     // lower diagonal -m (:301, :312)
        __syncthreads();
        if (ok) {
            const int kend = min(128, n - k0);
#pragma unroll 8
            for (int q = 0; q < kend; ++q) {
                const double b = __ldg(B + (size_t)(k0 + q) * n);
                ax = fma(diag[q].x, b, ax);
                ay = fma(diag[q].y, b, ay);
            }
        }
    }
    if (!ok) return;
    const double inv = 1.0 / (double)N;                     // omega_out /= N (:324)
    if (m == 0) {
        omega[(size_t)el * el + el] = ay * inv;             // Re(sum / 1j) = Im(sum) (:306-307)
    } else {
        const double s = ((m & 1) ? -1.0 : 1.0) * 1.4142135623730951;
        omega[(size_t)el * el + el + m] = s * ay * inv;     // sqrt2 * sgn * Im (:318)
        omega[(size_t)el * el + el - m] = -s * ax * inv;    // -sqrt2 * sgn * Re (:322)
    }
}

// one warp per position k of diagonal m; lanes stride over el, fixed shuffle tree
__global__ void __launch_bounds__(256)
k_shr2mat(const double *__restrict__ omega, const double *__restrict__ basis, double2 *__restrict__ W, int N, int Nmax)
{
    const int m = blockIdx.y;
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int n = N - m;
    if (k >= n) return;
    const int lane = threadIdx.x & 31;
    const double *__restrict__ B = basis + basis_break(m, N) + (size_t)k * n;
    double ax = 0.0, ay = 0.0;
    const double c = 0.7071067811865476;
    for (int el = m + lane; el < Nmax; el += 32) {
        const double b = __ldg(B + (el - m));
        if (m == 0) {
            ax = fma(b, omega[(size_t)el * el + el], ax);                                     // :203-207
        } else {
            ax = fma(b, c * omega[(size_t)el * el + el + m], ax);                             // omega_complex (:215)
            ay = fma(b, -c * omega[(size_t)el * el + el - m], ay);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        ax += __shfl_xor_sync(0xffffffffu, ax, o);
        ay += __shfl_xor_sync(0xffffffffu, ay, o);
    }
    if (lane != 0) return;
    if (m & 1) { ax = -ax; ay = -ay; }                                                        // sgn (:218-219)
    // W_out *= 1j (:227):  upper diagonal gets 1j * d, lower diagonal 1j * conj(d)  (:220-223)
    W[(size_t)k * N + (k + m)] = make_double2(-ay, ax);
    if (m > 0) W[(size_t)(k + m) * N + k] = make_double2(ay, ax);
}

__global__ void k_zero_d2(double2 *X, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) X[i] = make_double2(0.0, 0.0);
}

static bool nmax_of(long long nomega, int N, int *Nmax)
{
    // omega holds (elmax + 1)^2 coefficients, elmax <= N - 1 (quantization.py:283-290, 181-185)
    long long r = (long long)llround(sqrt((double)nomega));
    if (r * r != nomega || r < 1) return false;
    *Nmax = (int)(r < N ? r : N);
    return true;
}

}   // namespace

extern "C" long long qf_basis_size(int N)
{
    return N >= 1 ? basis_break(N, N) : 0;      // sum_{m<N} (N-m)^2
}

extern "C" int qf_mat2shr(qf_handle_t h, const void *W_dev, const void *basis_dev, void *omega_dev, long long nomega, void *stream)
{
    if (!h || !W_dev || !basis_dev || !omega_dev) { qf_set_error("qf_mat2shr: null argument"); return QF_ERR_INVALID; }
    int Nmax = 0;
    if (!nmax_of(nomega, h->N, &Nmax)) { qf_set_error("qf_mat2shr: the number of coefficients must be (elmax+1)^2, got %lld", nomega); return QF_ERR_INVALID; }
    QF_ON_DEVICE(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    QF_CUDA(cudaMemsetAsync(omega_dev, 0, sizeof(double) * (size_t)nomega, st));
    k_mat2shr<<<dim3((Nmax + 127) / 128, Nmax), 128, 0, st>>>((const double2 *)W_dev, (const double *)basis_dev, (double *)omega_dev, h->N, Nmax);
    h->launches++;
    QF_CUDA(cudaGetLastError());
    return QF_OK;
}

extern "C" int qf_shr2mat(qf_handle_t h, const void *omega_dev, long long nomega, const void *basis_dev, void *W_dev, void *stream)
{
    if (!h || !W_dev || !basis_dev || !omega_dev) { qf_set_error("qf_shr2mat: null argument"); return QF_ERR_INVALID; }
    int Nmax = 0;
    if (!nmax_of(nomega, h->N, &Nmax)) { qf_set_error("qf_shr2mat: the number of coefficients must be (elmax+1)^2, got %lld", nomega); return QF_ERR_INVALID; }
    QF_ON_DEVICE(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int N = h->N;
    k_zero_d2<<<h->sm_count * 4, 256, 0, st>>>((double2 *)W_dev, h->mat_elems);
    k_shr2mat<<<dim3((N + 7) / 8, Nmax), 256, 0, st>>>((const double *)omega_dev, (const double *)basis_dev, (double2 *)W_dev, N, Nmax);
    h->launches += 2;
    QF_CUDA(cudaGetLastError());
    return QF_OK;
}
