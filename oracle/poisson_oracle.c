/*
 * TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT PATH.
 *
 * CPU restatement (plain C, OpenMP across diagonals) of the Poisson / Laplace
 * kernels on quflow's isomp hot path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * quflow_b200 package never does.
 *
 * Each function cites the reference lines it follows (paths relative to the
 * upstream quflow tree).  The arithmetic is written in the same per-element
 * operation order as the reference numba kernels; numba compiles those with
 * fastmath=True, so agreement with the reference is to rounding (1e-16 ..
 * 1e-14 relative, see tests/test_oracle.py), not bit-for-bit.
 *
 * Parity status: PINNED — tests/test_oracle.py checks this file against
 * fixtures generated from the reference itself (oracle/gen_golden.py) and
 * against the reference's own known answers (tests/golden/README.md).
 */
#include <complex.h>
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

typedef double complex zc;

/* quflow/laplacian/cpu.py:34-42  mk2ij */
static inline void mk2ij(int m, int k, int *i, int *j)
{
    if (m >= 0) { *i = k; *j = k + m; }
    else        { *i = k - m; *j = k; }
}

/* quflow/laplacian/cpu.py:55-95  _compute_cpu_laplacian
 * lap is (N,N,2) row-major: [..,0] diagonal entry, [..,1] coupling to position k-1.
 * bc_shift is added to lap[0,0,0]: the reference default is -0.5 (cpu.py:90),
 * the legacy backend used +0.5 (laplacian/gpu.py:73). Pass 0.0 for bc=False. */
void qfo_compute_laplacian(int N, double bc_shift, double *lap)
{
    memset(lap, 0, sizeof(double) * (size_t)N * N * 2);
    for (int m = -N + 1; m < N; ++m) {
        int absm = m < 0 ? -m : m;
        for (int k = 0; k < N - absm; ++k) {
            int i, j;
            mk2ij(m, k, &i, &j);
            double dk = (double)k, dm = (double)absm, dN = (double)N;
            lap[((size_t)i * N + j) * 2 + 0] = -((dN - 1.0) * (2.0 * dk + 1.0 + dm) - 2.0 * dk * (dk + dm));
            lap[((size_t)i * N + j) * 2 + 1] = sqrt(((dk + dm) * (dN - dk - dm)) * (dk * (dN - dk)));
        }
    }
    lap[0] += bc_shift;
}

/* quflow/laplacian/cpu.py:281-362  _solve_cpu_skewh
 * Thomas algorithm per upper diagonal m, complex rhs, real coefficients.
 * remove_trace bit 0: subtract mean(diag W) from the m==0 rhs (cpu.py:311-317,327-328);
 * remove_trace bit 1: subtract mean(diag P) afterwards (cpu.py:342-352).
 * The reference default is 3.  The legacy backend laplacian/gpu.py:143-173 only does
 * the second (value 2, with bc +0.5); that mode exists solely to replay the reference's
 * stale N=16 golden vector. */
void qfo_solve_poisson_skewh(int N, const double *lap, const zc *W, zc *P,
                             double *buf_float, zc *buf_complex, int remove_trace)
{
#pragma omp parallel for schedule(dynamic, 4)
    for (int m = 0; m < N; ++m) {
        int i, j;
        zc trW = 0.0;
        mk2ij(m, 0, &i, &j);
        buf_float[(size_t)i * N + j] = lap[((size_t)i * N + j) * 2];
        buf_complex[(size_t)i * N + j] = W[(size_t)i * N + j];
        if (m == 0 && (remove_trace & 1)) {
            trW = W[0];
            for (int k = 1; k < N; ++k) trW += W[(size_t)k * N + k];
            trW /= N;
            buf_complex[(size_t)i * N + j] -= trW;
        }
        /* forward sweep  cpu.py:320-328 */
        for (int k = 1; k < N - m; ++k) {
            mk2ij(m, k, &i, &j);
            size_t ij = (size_t)i * N + j, pm = (size_t)(i - 1) * N + (j - 1);
            double w = lap[ij * 2 + 1] / buf_float[pm];
            buf_float[ij] = lap[ij * 2] - w * lap[ij * 2 + 1];
            buf_complex[ij] = W[ij] - w * buf_complex[pm];
            if (m == 0 && (remove_trace & 1)) buf_complex[ij] -= trW;
        }
        /* backward sweep  cpu.py:331-340 */
        mk2ij(m, N - m - 1, &i, &j);
        {
            size_t ij = (size_t)i * N + j;
            P[ij] = buf_complex[ij] / buf_float[ij];
            if (m != 0) P[(size_t)j * N + i] = -conj(P[ij]);
        }
        for (int k = N - m - 2; k >= 0; --k) {
            mk2ij(m, k, &i, &j);
            size_t ij = (size_t)i * N + j, pp = (size_t)(i + 1) * N + (j + 1);
            P[ij] = (buf_complex[ij] - lap[pp * 2 + 1] * P[pp]) / buf_float[ij];
            if (m != 0) P[(size_t)j * N + i] = -conj(P[ij]);
        }
        /* cpu.py:342-352 */
        if (m == 0 && (remove_trace & 2)) {
            zc trP = P[0];
            for (int k = 1; k < N; ++k) trP += P[(size_t)k * N + k];
            trP /= N;
            for (int k = 0; k < N; ++k) P[(size_t)k * N + k] -= trP;
        }
    }
}

/* quflow/laplacian/cpu.py:98-108  _dot_cpu_generic  (the kernel behind laplace()) */
void qfo_laplace(int N, const double *lap, const zc *P, zc *W)
{
#pragma omp parallel for
    for (int i = 0; i < N; ++i) {
        for (int j = 0; j < N; ++j) {
            size_t ij = (size_t)i * N + j;
            zc w = lap[ij * 2] * P[ij];
            if (i < N - 1 && j < N - 1) {
                size_t pp = (size_t)(i + 1) * N + (j + 1);
                w += lap[pp * 2 + 1] * P[pp];
            }
            if (i > 0 && j > 0) {
                size_t pm = (size_t)(i - 1) * N + (j - 1);
                w += lap[ij * 2 + 1] * P[pm];
            }
            W[ij] = w;
        }
    }
}

/* quflow/integrators/isospectral.py:66-81  conj_subtract_ (2-D branch), in place allowed */
void qfo_conj_subtract(int N, const zc *a, zc *out)
{
    for (int i = 0; i < N; ++i) {
        out[(size_t)i * N + i] = a[(size_t)i * N + i] - conj(a[(size_t)i * N + i]);
        for (int j = 0; j < i; ++j) {
            zc v = a[(size_t)i * N + j] - conj(a[(size_t)j * N + i]);
            out[(size_t)i * N + j] = v;
            out[(size_t)j * N + i] = -conj(v);
        }
    }
}

/* scipy.linalg.norm(A, ord=inf) as used at isospectral.py:534 and np.linalg.norm(W, inf) at :448:
 * max over rows of the sum of complex moduli. */
double qfo_norm_inf(int N, const zc *A)
{
    double best = 0.0;
#pragma omp parallel for reduction(max : best)
    for (int i = 0; i < N; ++i) {
        double s = 0.0;
        for (int j = 0; j < N; ++j) s += cabs(A[(size_t)i * N + j]);
        if (s > best) best = s;
    }
    return best;
}
