// Complex-FP64 GEMM  C = A * B  (N x N, row-major interleaved complex128) on the FP64 tensor path of
// sm_100a: DMMA  mma.sync.aligned.m8n8k4.f64  (the only FP64 MMA shape Blackwell's SASS has — the larger
// PTX shapes are split into DMMA.8x8x4 by ptxas; tcgen05 has no f64 kind).
//
// Replaces the two zgemm calls of the fixed-point iteration, np.matmul(Phalf, Whalf, out=PWcomm) and
// np.matmul(PWcomm, Phalf, out=dW)  (quflow/integrators/isospectral.py:496,499).
//
// Layout.  CTA tile 128 x 64 complex, K step 16, 8 warps (4 along M x 2 along N), warp tile 32 x 32:
// 4 x 4 sub-tiles of 8 x 8, each holding a real and an imaginary accumulator fragment (128 registers).
// A complex product is four real DMMAs:  Cre += Are*Bre + (-Aim)*Bim,  Cim += Are*Bim + Aim*Bre.
// Operands sit in shared memory as 128-byte rows in the TMA SWIZZLE_128B pattern (16-byte chunk c of row r
// is stored at chunk c ^ (r & 7)); one LDS.128 per fragment delivers (re, im) of one complex element.
// The K index inside an MMA is permuted (lane%4 = t uses k = 2t + s) so that the eight lanes of every
// quarter-warp hit eight different chunks: all fragment loads are bank-conflict free.
#include <algorithm>

#include "qf_common.cuh"

namespace {

constexpr int BM = 128, BN = 64, BK = 16;
constexpr int STAGES = 3;
constexpr int A_STAGE_BYTES = BM * BK * 16;          // 32 KiB: 2 boxes [128 rows][128 B]
constexpr int B_STAGE_BYTES = BK * BN * 16;          // 16 KiB: 8 boxes [16 rows][128 B]
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_SMEM = STAGES * STAGE_BYTES + 1024;   // + slack for 1024-byte alignment

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ double flip_sign(double x)
{
    return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x));
}

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void *gptr, bool valid)
{
    const int sz = valid ? 16 : 0;   // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gptr), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ double2 lds128(uint32_t addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

// Load one K tile (k0 .. k0+15) of the A panel (rows row0 .. row0+127) and of the B panel (cols col0 .. col0+63).
__device__ __forceinline__ void load_stage(uint32_t sA, uint32_t sB, const double2 *__restrict__ A,
                                           const double2 *__restrict__ B, int N, int row0, int row_end, int col0, int k0,
                                           int tid)
{
#pragma unroll
    for (int q = 0; q < (BM * BK) / GEMM_THREADS; ++q) {
        const int idx = tid + q * GEMM_THREADS;
        const int k = idx & (BK - 1), r = idx >> 4;
        const int gr = row0 + r, gk = k0 + k;
        const bool ok = (gr < row_end) && (gk < N);
        const double2 *src = A + (ok ? ((size_t)gr * N + gk) : 0);
        const uint32_t dst = sA + (k >> 3) * (BM * 128) + r * 128 + ((((k & 7) ^ (r & 7))) << 4);
        cp_async16(dst, src, ok);
    }
#pragma unroll
    for (int q = 0; q < (BK * BN) / GEMM_THREADS; ++q) {
        const int idx = tid + q * GEMM_THREADS;
        const int n = idx & (BN - 1), k = idx >> 6;
        const int gk = k0 + k, gc = col0 + n;
        const bool ok = (gk < N) && (gc < N);
        const double2 *src = B + (ok ? ((size_t)gk * N + gc) : 0);
        const uint32_t dst = sB + (n >> 3) * (BK * 128) + k * 128 + ((((n & 7) ^ (k & 7))) << 4);
        cp_async16(dst, src, ok);
    }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
k_zgemm(const double2 *__restrict__ Ag, const double2 *__restrict__ Bg, double2 *__restrict__ Cg, int N, int row_begin,
        int row_end, int upper_only, const QfCtrl *__restrict__ ctrl, int gated)
{
    const int b = blockIdx.z;
    if (gated && !ctrl[b].active) return;
    const int row0 = row_begin + blockIdx.y * BM;
    const int col0 = blockIdx.x * BN;
    if (upper_only && (col0 + BN - 1 < row0)) return;   // tile entirely below the diagonal

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = ((uint32_t)__cvta_generic_to_shared(smem_raw) + 1023u) & ~1023u;

    const size_t moff = (size_t)b * N * N;
    const double2 *A = Ag + moff;
    const double2 *B = Bg + moff;
    double2 *C = Cg + moff;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;   // 4 x 2 warps
    const int g = lane >> 2, t = lane & 3;

    double acc_re[4][4][2], acc_im[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc_re[i][j][0] = acc_re[i][j][1] = 0.0;
            acc_im[i][j][0] = acc_im[i][j][1] = 0.0;
        }

    const int KT = (N + BK - 1) / BK;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT)
            load_stage(smem_base + s * STAGE_BYTES, smem_base + s * STAGE_BYTES + A_STAGE_BYTES, A, B, N, row0, row_end, col0,
                       s * BK, tid);
        cp_async_commit();
    }

    // per-thread fragment base offsets inside a stage (see file header for the k permutation)
    uint32_t a_off[2], b_off[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int kk = 2 * t + s;
        a_off[s] = (wm * 32 + g) * 128 + ((kk ^ g) << 4);
        b_off[s] = A_STAGE_BYTES + wn * 4 * (BK * 128) + kk * 128 + ((g ^ kk) << 4);
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + STAGES - 1;
            if (nk < KT) {
                const uint32_t sb = smem_base + (nk % STAGES) * STAGE_BYTES;
                load_stage(sb, sb + A_STAGE_BYTES, A, B, N, row0, row_end, col0, nk * BK, tid);
            }
            cp_async_commit();
        }
        const uint32_t sb = smem_base + (kt % STAGES) * STAGE_BYTES;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                double a_re[4], a_im[4], a_in[4], b_re[4], b_im[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double2 v = lds128(sb + a_off[s] + hh * (BM * 128) + i * 1024);
                    a_re[i] = v.x;
                    a_im[i] = v.y;
                    a_in[i] = flip_sign(v.y);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double2 v = lds128(sb + b_off[s] + hh * 1024 + j * (BK * 128));
                    b_re[j] = v.x;
                    b_im[j] = v.y;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        dmma884(acc_re[i][j][0], acc_re[i][j][1], a_re[i], b_re[j]);
                        dmma884(acc_im[i][j][0], acc_im[i][j][1], a_re[i], b_im[j]);
                    }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        dmma884(acc_re[i][j][0], acc_re[i][j][1], a_in[i], b_im[j]);
                        dmma884(acc_im[i][j][0], acc_im[i][j][1], a_im[i], b_re[j]);
                    }
            }
        }
    }
    cp_async_wait<0>();

    // epilogue: each thread owns, per 8x8 sub-tile, row g and the two adjacent columns 2t, 2t+1
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = row0 + wm * 32 + i * 8 + g;
        if (r >= row_end) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = col0 + wn * 32 + j * 8 + 2 * t;
            double2 *dst = C + (size_t)r * N + c;
            if (c < N) dst[0] = make_double2(acc_re[i][j][0], acc_im[i][j][0]);
            if (c + 1 < N) dst[1] = make_double2(acc_re[i][j][1], acc_im[i][j][1]);
        }
    }
}

}   // namespace

struct QfGemmPlan {
    int smem_bytes;
};

int qf_gemm_create(qf_handle_s *h)
{
    h->gemm = new QfGemmPlan{GEMM_SMEM};
    QF_CUDA(cudaFuncSetAttribute(k_zgemm, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
    return QF_OK;
}

void qf_gemm_destroy(qf_handle_s *h)
{
    delete h->gemm;
    h->gemm = nullptr;
}

int qf_launch_zgemm(qf_handle_s *h, const double2 *A, const double2 *B, double2 *C, bool upper_only, bool gated,
                    int row_begin, int row_end, cudaStream_t st)
{
    const int N = h->N;
    dim3 grid((N + BN - 1) / BN, (row_end - row_begin + BM - 1) / BM, h->batch);
    k_zgemm<<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(A, B, C, N, row_begin, row_end, upper_only ? 1 : 0, h->ctrl, gated ? 1 : 0);
    h->launches++;
    QF_CUDA(cudaGetLastError());
    return QF_OK;
}
