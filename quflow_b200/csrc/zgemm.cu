// Complex-FP64 GEMM  C = A * B  (N x N, row-major interleaved complex128) on the FP64 tensor path of
// sm_100a: DMMA  mma.sync.aligned.m8n8k4.f64  (the only FP64 MMA shape Blackwell's SASS has — the larger
// PTX shapes are split into DMMA.8x8x4 by ptxas; tcgen05 has no f64 kind).
//
// Replaces the two zgemm calls of the fixed-point iteration, np.matmul(Phalf, Whalf, out=PWcomm) and
// np.matmul(PWcomm, Phalf, out=dW)  (quflow/integrators/isospectral.py:496,499).
//
// Two arithmetic variants share all of the machinery below (template parameter M3):
//   4M  CTA tile 128 x 64, 8 warps as 4 x 2, warp tile 32 x 32.  A complex product is four real DMMAs:
//         Cre += Are*Bre + (-Aim)*Bim,  Cim += Are*Bim + Aim*Bre          (8 real flop per complex MAC)
//   3M  CTA tile 64 x 64, 8 warps as 2 x 4, warp tile 32 x 16.  Karatsuba / "ZGEMM3M": three real DMMAs
//         T1 += Are*Bre,  T2 += Aim*Bim,  T3 += (Are+Aim)*(Bre+Bim);   Cre = T1-T2,  Cim = T3-T1-T2   (6 real flop)
//       25 % fewer FP64 tensor instructions for a normwise error of the same order.
//
// Operands sit in shared memory as 128-byte rows in the TMA SWIZZLE_128B pattern (16-byte chunk c of row r
// is stored at chunk c ^ (r & 7)); one LDS.128 per fragment delivers (re, im) of one complex element.
// The K index inside an MMA is permuted (lane%4 = t uses k = 2t + s) so that the eight lanes of every
// quarter-warp hit eight different chunks: all fragment loads are bank-conflict free.
// Tiles are fed by TMA (cp.async.bulk.tensor, one mbarrier per stage) with a cp.async fallback, and scheduled
// stream-K: persistent CTAs split the (tile, k) iteration space evenly (see k_zgemm_sk).
#include <cuda.h>

#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "qf_common.cuh"
#include "qf_tma.cuh"

namespace {

constexpr int GEMM_THREADS = 256;
constexpr int MI = 4;   // 8-row sub-tiles per warp (warp tile is 32 rows in both variants)

// M3: 0 = 4M arithmetic (128 x 64 tiles), 1 = 3M arithmetic (64 x 64 tiles), 2 = 3M arithmetic with 64 x 32 tiles for
// small problems (twice as many tiles: every SM owns a whole, unsplit tile where 64 x 64 tiles would fill half the GPU)
template <int M3>
struct Cfg;
template <>
struct Cfg<0> {
    static constexpr int BM = 128, BN = 64, BK = 16, WN = 2, NJ = 4, NACC = 2, STAGES = 3;
};
template <>
struct Cfg<1> {
    static constexpr int BM = 64, BN = 64, BK = 32, WN = 4, NJ = 2, NACC = 3, STAGES = 3;
};
template <>
struct Cfg<2> {
    static constexpr int BM = 64, BN = 32, BK = 32, WN = 4, NJ = 1, NACC = 3, STAGES = 3;
};
template <int M3>
struct Geo {
    using C = Cfg<M3>;
    static constexpr int A_STAGE_BYTES = C::BM * C::BK * 16;   // BK/8 boxes [BM rows][128 B]
    static constexpr int B_STAGE_BYTES = C::BK * C::BN * 16;   // BN/8 boxes [BK rows][128 B]
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int SMEM = C::STAGES * STAGE_BYTES + 1024;   // + slack for 1024-byte alignment
    static constexpr int ACC_D2 = MI * C::NJ * C::NACC;           // double2 per thread in a partial tile
};
constexpr int WS_D2_PER_THREAD = 32;   // workspace slot = 32 double2 per thread (max over the variants)

// Stream-K finisher: wait for a contributor's partial tile.  The contributor computed its share FIRST, so with all CTAs
// resident (cooperative launch) the wait is short; without co-residency (QF_GEMM_COOP=0 on a busy GPU) it could never
// end, so it is bounded: after about five seconds the kernel traps (a launch failure the host reports) instead of hanging.
__device__ __forceinline__ void sk_wait_partial(int *flag)
{
    const long long t0 = clock64();
    while (atomicAdd(flag, 0) == 0) {
        __nanosleep(64);
        if (clock64() - t0 > 10000000000ll) __trap();
    }
}

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

// ---- cp.async loader (fallback when the driver has no tensor-map encoder, or QF_GEMM_LOAD=cpasync) --------
template <int M3>
__device__ __forceinline__ void load_stage(uint32_t sA, uint32_t sB, const double2 *__restrict__ A,
                                           const double2 *__restrict__ B, int N, int row0, int row_end, int col0, int k0,
                                           int tid)
{
    constexpr int BM = Cfg<M3>::BM, BN = Cfg<M3>::BN, BK = Cfg<M3>::BK;
#pragma unroll
    for (int q = 0; q < (BM * BK) / GEMM_THREADS; ++q) {
        const int idx = tid + q * GEMM_THREADS;
        const int k = idx & (BK - 1), r = idx / BK;
        const int gr = row0 + r, gk = k0 + k;
        const bool ok = (gr < row_end) && (gk < N);
        const double2 *src = A + (ok ? ((size_t)gr * N + gk) : 0);
        const uint32_t dst = sA + (k >> 3) * (BM * 128) + r * 128 + ((((k & 7) ^ (r & 7))) << 4);
        cp_async16(dst, src, ok);
    }
#pragma unroll
    for (int q = 0; q < (BK * BN) / GEMM_THREADS; ++q) {
        const int idx = tid + q * GEMM_THREADS;
        const int n = idx & (BN - 1), k = idx / BN;
        const int gk = k0 + k, gc = col0 + n;
        const bool ok = (gk < N) && (gc < N);
        const double2 *src = B + (ok ? ((size_t)gk * N + gc) : 0);
        const uint32_t dst = sB + (n >> 3) * (BK * 128) + k * 128 + ((((n & 7) ^ (k & 7))) << 4);
        cp_async16(dst, src, ok);
    }
}

// ---- the DMMA inner product over one resident stage (BK/4 k-steps of 4) -----------------------------------
// Fragments are double-buffered in registers: the LDS.128 (and, for 3M, the DADDs forming Are+Aim / Bre+Bim) of
// k-step ks+1 are issued before the DMMAs of k-step ks, so the tensor pipe never waits on shared-memory latency
// inside a stage.
template <int M3>
struct Frag {
    double a_re[MI], a_im[MI], a_x[MI], b_re[Cfg<M3>::NJ], b_im[Cfg<M3>::NJ], b_x[Cfg<M3>::NJ];
};

template <int M3>
__device__ __forceinline__ void gemm_load_frag(Frag<M3> &f, uint32_t sb, const uint32_t (&a_off)[2], const uint32_t (&b_off)[2],
                                               int ks)
{
    constexpr int BM = Cfg<M3>::BM, NJ = Cfg<M3>::NJ, BK = Cfg<M3>::BK;
    const int hh = ks >> 1, s = ks & 1;
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const double2 v = lds128(sb + a_off[s] + hh * (BM * 128) + i * 1024);
        f.a_re[i] = v.x;
        f.a_im[i] = v.y;
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const double2 v = lds128(sb + b_off[s] + hh * 1024 + j * (BK * 128));
        f.b_re[j] = v.x;
        f.b_im[j] = v.y;
    }
}

// derived operands: 3M  Are+Aim, Bre+Bim (DADD on the FP64 pipe);  4M  -Aim (integer XOR)
template <int M3>
__device__ __forceinline__ void gemm_finish_frag(Frag<M3> &f)
{
#pragma unroll
    for (int i = 0; i < MI; ++i) f.a_x[i] = M3 ? (f.a_re[i] + f.a_im[i]) : flip_sign(f.a_im[i]);
#pragma unroll
    for (int j = 0; j < Cfg<M3>::NJ; ++j) f.b_x[j] = M3 ? (f.b_re[j] + f.b_im[j]) : 0.0;
}

template <int M3>
__device__ __forceinline__ void gemm_mma_frag(double (&acc)[MI][Cfg<M3>::NJ][Cfg<M3>::NACC][2], const Frag<M3> &f)
{
    constexpr int NJ = Cfg<M3>::NJ;
    if (M3) {
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0][0], acc[i][j][0][1], f.a_re[i], f.b_re[j]);   // T1
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][1][0], acc[i][j][1][1], f.a_im[i], f.b_im[j]);   // T2
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][2][0], acc[i][j][2][1], f.a_x[i], f.b_x[j]);     // T3
    } else {
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                dmma884(acc[i][j][0][0], acc[i][j][0][1], f.a_re[i], f.b_re[j]);
                dmma884(acc[i][j][1][0], acc[i][j][1][1], f.a_re[i], f.b_im[j]);
            }
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                dmma884(acc[i][j][0][0], acc[i][j][0][1], f.a_x[i], f.b_im[j]);
                dmma884(acc[i][j][1][0], acc[i][j][1][1], f.a_im[i], f.b_re[j]);
            }
    }
}

template <int M3>
__device__ __forceinline__ void gemm_compute_stage(double (&acc)[MI][Cfg<M3>::NJ][Cfg<M3>::NACC][2], uint32_t sb,
                                                   const uint32_t (&a_off)[2], const uint32_t (&b_off)[2])
{
    constexpr int KS = Cfg<M3>::BK / 4;
    if (M3) {
        // order per k-step: LDS(ks+1) -> DMMAs(ks) -> DADDs(ks+1): the in-order warp never waits on the LDS
        Frag<M3> f[2];
        gemm_load_frag<M3>(f[0], sb, a_off, b_off, 0);
        gemm_finish_frag<M3>(f[0]);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            if (ks + 1 < KS) gemm_load_frag<M3>(f[(ks + 1) & 1], sb, a_off, b_off, ks + 1);
            gemm_mma_frag<M3>(acc, f[ks & 1]);
            if (ks + 1 < KS) gemm_finish_frag<M3>(f[(ks + 1) & 1]);
        }
    } else {
        // the 4M variant already runs at the register limit (128 accumulator registers): single-buffered fragments
        Frag<M3> f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            gemm_load_frag<M3>(f, sb, a_off, b_off, ks);
            gemm_finish_frag<M3>(f);
            gemm_mma_frag<M3>(acc, f);
        }
    }
}

template <int M3>
__device__ __forceinline__ void gemm_frag_offsets(uint32_t (&a_off)[2], uint32_t (&b_off)[2], int wm, int wn, int g, int t)
{
    // per-thread fragment base offsets inside a stage (see file header for the k permutation)
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int kk = 2 * t + s;
        a_off[s] = (wm * 32 + g) * 128 + ((kk ^ g) << 4);
        b_off[s] = Geo<M3>::A_STAGE_BYTES + wn * Cfg<M3>::NJ * (Cfg<M3>::BK * 128) + kk * 128 + ((g ^ kk) << 4);
    }
}

// Accumulate k-tiles [kt_begin, kt_end) of one output tile; cp.async multi-stage pipeline.
template <int M3>
__device__ __forceinline__ void gemm_mainloop(double (&acc)[MI][Cfg<M3>::NJ][Cfg<M3>::NACC][2], uint32_t smem_base,
                                              const double2 *__restrict__ A, const double2 *__restrict__ B, int N, int row0,
                                              int row_end, int col0, int kt_begin, int kt_end, int tid, int wm, int wn, int g, int t)
{
    constexpr int STAGES = Cfg<M3>::STAGES, STAGE_BYTES = Geo<M3>::STAGE_BYTES, A_STAGE_BYTES = Geo<M3>::A_STAGE_BYTES;
    constexpr int BK = Cfg<M3>::BK;
    const int nkt = kt_end - kt_begin;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nkt)
            load_stage<M3>(smem_base + s * STAGE_BYTES, smem_base + s * STAGE_BYTES + A_STAGE_BYTES, A, B, N, row0, row_end, col0,
                           (kt_begin + s) * BK, tid);
        cp_async_commit();
    }
    uint32_t a_off[2], b_off[2];
    gemm_frag_offsets<M3>(a_off, b_off, wm, wn, g, t);
    for (int kt = 0; kt < nkt; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + STAGES - 1;
            if (nk < nkt) {
                const uint32_t sb = smem_base + (nk % STAGES) * STAGE_BYTES;
                load_stage<M3>(sb, sb + A_STAGE_BYTES, A, B, N, row0, row_end, col0, (kt_begin + nk) * BK, tid);
            }
            cp_async_commit();
        }
        gemm_compute_stage<M3>(acc, smem_base + (kt % STAGES) * STAGE_BYTES, a_off, b_off);
    }
    cp_async_wait<0>();
    __syncthreads();   // all warps are done with shared memory: the next segment may refill it
}

// Same main loop fed by TMA: one elected thread issues 2 (A) + BN/8 (B) cp.async.bulk.tensor boxes of 128-byte rows
// per stage into the SWIZZLE_128B layout; completion is tracked by one mbarrier per stage (expect_tx = stage bytes).
// `gk` counts the k-tiles this CTA has consumed since kernel start: stage = gk % STAGES, phase = (gk / STAGES) & 1.
template <int M3>
__device__ __forceinline__ void gemm_mainloop_tma(double (&acc)[MI][Cfg<M3>::NJ][Cfg<M3>::NACC][2], uint32_t smem_base,
                                                  uint32_t bars, const CUtensorMap *tmA, const CUtensorMap *tmB, int member,
                                                  int a_mem_row0, int col0, int kt_begin, int kt_end, uint32_t &gk, int tid,
                                                  int wm, int wn, int g, int t)
{
    constexpr int STAGES = Cfg<M3>::STAGES, STAGE_BYTES = Geo<M3>::STAGE_BYTES, A_STAGE_BYTES = Geo<M3>::A_STAGE_BYTES;
    constexpr int BM = Cfg<M3>::BM, BN = Cfg<M3>::BN, BK = Cfg<M3>::BK;
    const int nkt = kt_end - kt_begin;
    auto issue = [&](int i) {
        const uint32_t idx = gk + (uint32_t)i;
        const uint32_t stage = idx % STAGES;
        const uint32_t sb = smem_base + stage * STAGE_BYTES;
        const uint32_t bar = bars + 8 * stage;
        const int k0 = (kt_begin + i) * BK;
        mbar_arrive_expect_tx(bar, STAGE_BYTES);
#pragma unroll
        for (int hh = 0; hh < BK / 8; ++hh) tma_load_3d(sb + hh * (BM * 128), tmA, 2 * (k0 + 8 * hh), a_mem_row0, member, bar);
#pragma unroll
        for (int c = 0; c < BN / 8; ++c) tma_load_3d(sb + A_STAGE_BYTES + c * (BK * 128), tmB, 2 * (col0 + 8 * c), k0, member, bar);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s)
            if (s < nkt) issue(s);
    }
    uint32_t a_off[2], b_off[2];
    gemm_frag_offsets<M3>(a_off, b_off, wm, wn, g, t);
    for (int kt = 0; kt < nkt; ++kt) {
        const uint32_t idx = gk + (uint32_t)kt;
        const uint32_t stage = idx % STAGES;
        __syncthreads();   // every warp finished k-tile kt-1: its stage may be refilled
        if (tid == 0 && kt + STAGES - 1 < nkt) issue(kt + STAGES - 1);
        mbar_wait(bars + 8 * stage, (idx / STAGES) & 1);
        gemm_compute_stage<M3>(acc, smem_base + stage * STAGE_BYTES, a_off, b_off);
    }
    gk += (uint32_t)nkt;
    __syncthreads();
}

// each thread owns, per 8x8 sub-tile, row g and the two adjacent columns 2t, 2t+1 (row0 / row_end: OUTPUT rows)
template <int M3>
__device__ __forceinline__ void gemm_store_tile(const double (&acc)[MI][Cfg<M3>::NJ][Cfg<M3>::NACC][2], double2 *__restrict__ C,
                                                int N, int row0, int row_end, int col0, int wm, int wn, int g, int t)
{
    constexpr int NJ = Cfg<M3>::NJ;
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int r = row0 + wm * 32 + i * 8 + g;
        if (r >= row_end) continue;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int c = col0 + wn * (8 * NJ) + j * 8 + 2 * t;
            double2 *dst = C + (size_t)r * N + c;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double re, im;
                if (M3) {
                    re = acc[i][j][0][e] - acc[i][j][1][e];
                    im = (acc[i][j][2][e] - acc[i][j][0][e]) - acc[i][j][1][e];
                } else {
                    re = acc[i][j][0][e];
                    im = acc[i][j][1][e];
                }
                if (c + e < N) dst[e] = make_double2(re, im);
            }
        }
    }
}

// ---- work schedule of a persistent CTA: data-parallel waves, then a stream-K tail ---------------------------------
// With G CTAs and ntiles tiles, the first floor(ntiles/G) "waves" are plain data-parallel: in wave w CTA c computes the
// whole tile w*G + c.  The tiles of a wave are consecutive in the row-major tile list, so at any time the CTAs share
// a few row panels of A and stream through B together: both stay L2-resident (handing every CTA a contiguous run of
// tiles instead keeps ALL of A and B live at once, 134 MB at N=2048, and re-reads them ~5x from DRAM).
// The remaining R < G tiles are split stream-K style: CTA c owns k-tile iterations [Tt c/G, Tt (c+1)/G) of Tt = R*KT.
struct SkSched {
    int G, cta, KT, nfull, w;
    long long Tt, it, it_end;
    int tail0;   // index of the first tail tile
    __device__ __forceinline__ void init(int ntiles, int KT_, int cta_, int G_)
    {
        G = G_; cta = cta_; KT = KT_;
        nfull = ntiles / G;
        w = 0;
        tail0 = nfull * G;
        Tt = (long long)(ntiles - tail0) * KT;
        it = Tt * cta / G;
        it_end = Tt * (cta + 1) / G;
    }
    // next segment: tile index, k-tile range [ka, kb); tail_local >= 0 for tail tiles (index inside the tail)
    __device__ __forceinline__ bool next(int &tile, int &ka, int &kb, int &tail_local)
    {
        if (w < nfull) {
            tile = w * G + cta;
            ka = 0;
            kb = KT;
            tail_local = -1;
            ++w;
            return true;
        }
        if (it < it_end) {
            tail_local = (int)(it / KT);
            tile = tail0 + tail_local;
            ka = (int)(it - (long long)tail_local * KT);
            kb = (int)min((long long)KT, ka + (it_end - it));
            it += kb - ka;
            return true;
        }
        return false;
    }
};

// ---- stream-K kernel: persistent CTAs split the (tile, k) iteration space evenly --------------------
// CTA c owns iterations [T c / G, T (c+1) / G) of the T = ntiles * KT k-tile iterations.  A tile whose k range is
// split is finished by the CTA that computed its k = 0 part (at the END of that CTA's range); the CTAs that hold the
// rest of the tile compute it FIRST in their own range, park the partial accumulators in their workspace slot and
// raise a flag.  Waits therefore only ever target work that was started at kernel start: no dependency cycles.
// Partials are added in CTA order, so the result is deterministic for a given (N, G).
//
// One output tile: rows [a_row0, row_end) of the A operand (logical matrix rows) times columns [col0, col0+BN) of B,
// written to rows c_row0.. of C (C may use the rank-permuted row layout of the multi-GPU path, see qf_common.cuh).
struct SkTile { int member, a_row0, c_row0, col0, row_end, op_row0, pad1, pad2; };   // op_row0: first row of the A operand in memory

template <int M3, bool TMA>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
k_zgemm_sk(const double2 *__restrict__ Ag, const double2 *__restrict__ Bg, double2 *__restrict__ Cg, int N,
           const SkTile *__restrict__ tiles, int ntiles, double2 *__restrict__ ws, int *__restrict__ flags,
           const QfCtrl *__restrict__ ctrl, int gated, const __grid_constant__ CUtensorMap tmA,
           const __grid_constant__ CUtensorMap tmB)
{
    constexpr int STAGES = Cfg<M3>::STAGES, NJ = Cfg<M3>::NJ, NACC = Cfg<M3>::NACC, WN = Cfg<M3>::WN;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = ((uint32_t)__cvta_generic_to_shared(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) unsigned long long mbar_storage[STAGES];
    const uint32_t bars = (uint32_t)__cvta_generic_to_shared(mbar_storage);
    uint32_t gk = 0;
    if (TMA) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; ++s) mbar_init(bars + 8 * s, 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp / WN, wn = warp % WN;
    const int g = lane >> 2, t = lane & 3;
    const int cta = blockIdx.x, G = gridDim.x;
    constexpr int BK = Cfg<M3>::BK;
    const int KT = (N + BK - 1) / BK;
    SkSched sched;
    sched.init(ntiles, KT, cta, G);
    int tile, ka, kb, tail_local;

    while (sched.next(tile, ka, kb, tail_local)) {
        const SkTile ti = tiles[tile];
        if (gated && !ctrl[ti.member].active) continue;
        const size_t moff = (size_t)ti.member * N * N;

        double acc[MI][NJ][NACC][2];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int c = 0; c < NACC; ++c) acc[i][j][c][0] = acc[i][j][c][1] = 0.0;

        // the A operand may itself be stored in the rank-permuted row layout (second GEMM): op_row0 is its memory row
        if (TMA)
            gemm_mainloop_tma<M3>(acc, smem_base, bars, &tmA, &tmB, ti.member, ti.op_row0, ti.col0, ka, kb, gk, tid, wm, wn, g, t);
        else
            gemm_mainloop<M3>(acc, smem_base, Ag + moff + ((ptrdiff_t)ti.op_row0 - ti.a_row0) * N, Bg + moff, N, ti.a_row0,
                              ti.row_end, ti.col0, ka, kb, tid, wm, wn, g, t);

        if (ka > 0) {
            // contributor: park the partial tile in this CTA's slot ([reg][thread] layout: coalesced)
            double2 *slot = ws + (size_t)cta * (WS_D2_PER_THREAD * GEMM_THREADS);
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j)
#pragma unroll
                    for (int c = 0; c < NACC; ++c)
                        __stcg(slot + ((i * NJ + j) * NACC + c) * GEMM_THREADS + tid, make_double2(acc[i][j][c][0], acc[i][j][c][1]));
            __threadfence();
            __syncthreads();
            if (tid == 0) atomicExch(flags + cta, 1);
        } else {
            if (kb < KT) {
                // finisher: add the parts computed by the following CTAs, in CTA order
                const long long tile_end = (long long)(tail_local + 1) * KT;
                int peer = cta + 1;
                long long covered = sched.it_end;
                while (covered < tile_end) {
                    if (tid == 0) {
                        sk_wait_partial(flags + peer);
                        atomicExch(flags + peer, 0);   // consume: flags are all zero again when the kernel ends
                    }
                    __syncthreads();
                    __threadfence();
                    const double2 *slot = ws + (size_t)peer * (WS_D2_PER_THREAD * GEMM_THREADS);
#pragma unroll
                    for (int i = 0; i < MI; ++i)
#pragma unroll
                        for (int j = 0; j < NJ; ++j)
#pragma unroll
                            for (int c = 0; c < NACC; ++c) {
                                const double2 p = __ldcg(slot + ((i * NJ + j) * NACC + c) * GEMM_THREADS + tid);
                                acc[i][j][c][0] += p.x;
                                acc[i][j][c][1] += p.y;
                            }
                    covered = sched.Tt * (peer + 1) / G;
                    ++peer;
                }
            }
            gemm_store_tile<M3>(acc, Cg + moff, N, ti.c_row0, ti.c_row0 + (ti.row_end - ti.a_row0), ti.col0, wm, wn, g, t);
        }
    }
}

// ---- fused tail of the fixed-point iteration (QfEpiPost, qf_common.cuh) --------------------------------------------
// Fragment layout of the 3M warp tile: sub-tile (i, j) holds row  r = row0 + wm*32 + i*8 + g  and the two adjacent columns
// c = col0 + wn*16 + j*8 + 2t + e.  Direct accesses: the four lanes of a quad cover 128 contiguous bytes of row r.
// Transposed accesses (A_ji, W_ji, dW_ji, W~_ji): the eight lanes with equal t cover 128 contiguous bytes of row c.
// Only elements on or above the diagonal (r <= c) are processed; their mirrors are written from the same registers.
// Tile-exchange path (xg.nranks > 1): the rank that owns the tile pair also stores the residual partials into every peer's
// copy (plain stores through the NVLink peer mappings), so that after the exchange barrier every rank evaluates the
// stopping rule on identical numbers; the W~ tiles follow in a copy kernel of their own (comm.cu: k_xchg_push_wh).
__device__ __forceinline__ void epi_post_tile(const double (&acc)[MI][Cfg<1>::NJ][Cfg<1>::NACC][2], const SkTile &ti, int N,
                                              int wm, int wn, int g, int t, const QfEpiPost &E, const QfXchg &xg)
{
    constexpr int NJ = Cfg<1>::NJ;
    const size_t moff = (size_t)ti.member * N * N;
    const double2 *__restrict__ A = E.A + moff;
    const double2 *__restrict__ W = E.W + moff;
    double2 *__restrict__ dW = E.dW + moff;
    double2 *__restrict__ Wh = E.Wh + moff;
    double *__restrict__ pd = E.part_direct + (size_t)ti.member * N * E.nsd;
    double *__restrict__ pm = E.part_mirror + (size_t)ti.member * N * E.nsm;
    const int cbase = ti.col0 + wn * (8 * NJ) + 2 * t;
    double colsum[NJ][2];
#pragma unroll
    for (int j = 0; j < NJ; ++j) colsum[j][0] = colsum[j][1] = 0.0;
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int r = ti.a_row0 + wm * 32 + i * 8 + g;
        const bool rok = r < ti.row_end;
        double rowsum = 0.0;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            double2 ad[2], at[2], dwo[2], wd[2], wt[2];
            bool ok[2];
            // the ten loads of this sub-tile are issued before the first use
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = cbase + j * 8 + e;
                ok[e] = rok && c < N && r <= c;
                if (ok[e]) {
                    const size_t rc = (size_t)r * N + c, cr = (size_t)c * N + r;
                    ad[e] = A[rc];
                    at[e] = __ldcg(A + cr);      // may have been stored by a peer (tile-exchange path): not through L1
                    dwo[e] = dW[rc];
                    wd[e] = W[rc];
                    wt[e] = W[cr];
                }
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = cbase + j * 8 + e;
                if (ok[e]) {
                    double2 sv = make_double2(acc[i][j][0][e] - acc[i][j][1][e], (acc[i][j][2][e] - acc[i][j][0][e]) - acc[i][j][1][e]);
                    if (r == c) sv.x = 0.0;                                        // P W P is skew-Hermitian
                    const double2 cm = zsub(ad[e], zconj(at[e]));                  // isospectral.py:66-81
                    const double2 d = zadd(sv, cm);                                // :499, :509
                    const double res = zabs(zsub(dwo[e], d));                      // :526, :534
                    const size_t rc = (size_t)r * N + c;
                    dW[rc] = d;
                    const double2 whd = zadd(wd[e], d);                            // :481-482
                    Wh[rc] = whd;
                    rowsum += res;
                    if (r < c) {
                        const size_t cr = (size_t)c * N + r;
                        const double2 dm = make_double2(-d.x, d.y);                // dW_ji = -conj(dW_ij)
                        dW[cr] = dm;
                        const double2 whm = zadd(wt[e], dm);
                        Wh[cr] = whm;
                        colsum[j][e] += res;                                       // the mirrored element has the same residual
                    }
                }
            }
        }
        // row r, columns of this warp: fixed shuffle tree over the quad
        rowsum += __shfl_xor_sync(0xffffffffu, rowsum, 1);
        rowsum += __shfl_xor_sync(0xffffffffu, rowsum, 2);
        if (t == 0 && rok) {
            const size_t at_ = (size_t)r * E.nsd + ((ti.col0 >> 4) + wn);
            pd[at_] = rowsum;
            if (xg.nranks > 1) {
                const size_t off = (size_t)xg.part2_off + (size_t)(pd - E.part_direct) + at_;
                for (int p = 0; p < xg.nranks; ++p)
                    if (p != xg.rank) xg.peerPart[p][off] = rowsum;
            }
        }
    }
    // column c (= row c of the mirrored half), the 32 rows of this warp: fixed tree over g
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double v = colsum[j][e];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            const int c = cbase + j * 8 + e;
            if (g == 0 && c < N) {
                const size_t at_ = (size_t)c * E.nsm + ((ti.a_row0 >> 5) + wm);
                pm[at_] = v;
                if (xg.nranks > 1) {
                    // the peers' partial arrays have the layout of this rank's: [direct | mirrored] in one allocation
                    const size_t off = (size_t)xg.part2_off + (size_t)(pm - E.part_direct) + at_;
                    for (int p = 0; p < xg.nranks; ++p)
                        if (p != xg.rank) xg.peerPart[p][off] = v;
                }
            }
        }
}

__device__ __forceinline__ void consumer_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// ---- warp-specialised stream-K kernel (3M arithmetic, TMA) ------------------------------------------------------
// Warp 8 is the producer: one lane walks the same (tile, k) segments as the consumers and keeps the STAGES-deep ring
// full (wait `empty`, arm `full` with expect_tx, issue the TMA boxes).  Warps 0-7 are the DMMA consumers: wait `full`,
// multiply, one arrive per warp on `empty`.  There is no CTA-wide barrier in the main loop, the producer runs ahead
// across tile boundaries while the consumers do the stream-K fix-up / store, and only the 256 consumer threads meet
// at the (named) barrier around the fix-up.
constexpr int WS_THREADS = GEMM_THREADS + 32;

template <int M3, bool POST>
__global__ void __launch_bounds__(WS_THREADS, 1)
k_zgemm3m_ws(double2 *__restrict__ Cg, int N, const SkTile *__restrict__ tiles, int ntiles, double2 *__restrict__ ws,
             int *__restrict__ flags, const QfCtrl *__restrict__ ctrl, int gated, const __grid_constant__ CUtensorMap tmA,
             const __grid_constant__ CUtensorMap tmB, const QfEpiPost epi, const QfXchg xg)
{
    const CUtensorMap *tmAp = &tmA;
    static_assert(M3 == 1 || M3 == 2, "warp-specialised kernel: 3M arithmetic only");
    static_assert(!POST || M3 == 1, "the fused tail is written for the 64 x 64 tile");
    constexpr int STAGES = Cfg<M3>::STAGES, NJ = Cfg<M3>::NJ, NACC = Cfg<M3>::NACC, WN = Cfg<M3>::WN;
    constexpr int BM = Cfg<M3>::BM, BN = Cfg<M3>::BN, BK = Cfg<M3>::BK;
    constexpr int STAGE_BYTES = Geo<M3>::STAGE_BYTES, A_STAGE_BYTES = Geo<M3>::A_STAGE_BYTES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = ((uint32_t)__cvta_generic_to_shared(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) unsigned long long bar_full[STAGES], bar_empty[STAGES];
    const uint32_t full = (uint32_t)__cvta_generic_to_shared(bar_full);
    const uint32_t empty = (uint32_t)__cvta_generic_to_shared(bar_empty);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full + 8 * s, 1);
            mbar_init(empty + 8 * s, GEMM_THREADS / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int cta = blockIdx.x, G = gridDim.x;
    const int KT = (N + BK - 1) / BK;
    SkSched sched;
    sched.init(ntiles, KT, cta, G);
    int tile, ka, kb, tail_local;
    uint32_t gk = 0;

    bool g1_seen = false;
    if (warp == GEMM_THREADS / 32) {
        // ===== producer =====
        if (lane == 0) {
            while (sched.next(tile, ka, kb, tail_local)) {
                const SkTile ti = tiles[tile];
                if (gated && !ctrl[ti.member].active) continue;
                for (int kt = ka; kt < kb; ++kt, ++gk) {
                    const uint32_t stage = gk % STAGES;
                    const uint32_t round = gk / STAGES;
                    if (round > 0) mbar_wait(empty + 8 * stage, (round - 1) & 1);   // consumers released the previous fill
                    const uint32_t sb = smem_base + stage * STAGE_BYTES;
                    const uint32_t bar = full + 8 * stage;
                    const int k0 = kt * BK;
                    mbar_arrive_expect_tx(bar, STAGE_BYTES);
#pragma unroll
                    for (int hh = 0; hh < BK / 8; ++hh)
                        tma_load_3d(sb + hh * (BM * 128), tmAp, 2 * (k0 + 8 * hh), ti.op_row0, ti.member, bar);
#pragma unroll
                    for (int c = 0; c < BN / 8; ++c)
                        tma_load_3d(sb + A_STAGE_BYTES + c * (BK * 128), &tmB, 2 * (ti.col0 + 8 * c), k0, ti.member, bar);
                }
            }
        }
        return;
    }

    // ===== consumers =====
    const int wm = warp / WN, wn = warp % WN;
    const int g = lane >> 2, t = lane & 3;
    uint32_t a_off[2], b_off[2];
    gemm_frag_offsets<M3>(a_off, b_off, wm, wn, g, t);
    while (sched.next(tile, ka, kb, tail_local)) {
        const SkTile ti = tiles[tile];
        if (gated && !ctrl[ti.member].active) continue;
        const size_t moff = (size_t)ti.member * N * N;

        double acc[MI][NJ][NACC][2];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int c = 0; c < NACC; ++c) acc[i][j][c][0] = acc[i][j][c][1] = 0.0;

        for (int kt = ka; kt < kb; ++kt, ++gk) {
            const uint32_t stage = gk % STAGES;
            mbar_wait(full + 8 * stage, (gk / STAGES) & 1);
            gemm_compute_stage<M3>(acc, smem_base + stage * STAGE_BYTES, a_off, b_off);
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + 8 * stage);
        }

        if (ka > 0) {
            double2 *slot = ws + (size_t)cta * (WS_D2_PER_THREAD * GEMM_THREADS);
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j)
#pragma unroll
                    for (int c = 0; c < NACC; ++c)
                        __stcg(slot + ((i * NJ + j) * NACC + c) * GEMM_THREADS + tid, make_double2(acc[i][j][c][0], acc[i][j][c][1]));
            __threadfence();
            consumer_bar_sync();
            if (tid == 0) atomicExch(flags + cta, 1);
        } else {
            if (kb < KT) {
                const long long tile_end = (long long)(tail_local + 1) * KT;
                int peer = cta + 1;
                long long covered = sched.it_end;
                while (covered < tile_end) {
                    if (tid == 0) {
                        sk_wait_partial(flags + peer);
                        atomicExch(flags + peer, 0);
                    }
                    consumer_bar_sync();
                    __threadfence();
                    const double2 *slot = ws + (size_t)peer * (WS_D2_PER_THREAD * GEMM_THREADS);
#pragma unroll
                    for (int i = 0; i < MI; ++i)
#pragma unroll
                        for (int j = 0; j < NJ; ++j)
#pragma unroll
                            for (int c = 0; c < NACC; ++c) {
                                const double2 p = __ldcg(slot + ((i * NJ + j) * NACC + c) * GEMM_THREADS + tid);
                                acc[i][j][c][0] += p.x;
                                acc[i][j][c][1] += p.y;
                            }
                    covered = sched.Tt * (peer + 1) / G;
                    ++peer;
                }
            }
            if constexpr (POST) {
                // GEMM 2 of the fixed-point iteration: dW, W~ and the residual partials straight from the accumulators
                if (xg.nranks > 1 && !g1_seen) {
                    // the transposed A tiles this rank needs were pushed by their owners during THEIR first GEMM: wait (once
                    // per CTA) until every peer has signalled that its GEMM 1 of this iteration is complete
                    if (tid == 0 && ctrl[0].nonfinite != 2) xchg_wait_flags(xg, QF_XF_G1, ctrl[0].gseq + 1ull);
                    consumer_bar_sync();
                    g1_seen = true;
                }
                epi_post_tile(acc, ti, N, wm, wn, g, t, epi, xg);
                continue;
            }
            gemm_store_tile<M3>(acc, Cg + moff, N, ti.c_row0, ti.c_row0 + (ti.row_end - ti.a_row0), ti.col0, wm, wn, g, t);
            // tile-exchange path, GEMM 1: a tile strictly below the diagonal is needed (transposed) by the rank that owns
            // its column block; it goes there as plain stores through the NVLink peer mapping while the next tile is
            // being multiplied
            if (xg.nranks > 1 && ti.col0 + BN <= ti.a_row0) {
                const int oc = qf_owner_of_row(ti.col0, xg.hb, xg.nranks);
                if (oc != xg.rank)
                    gemm_store_tile<M3>(acc, xg.peerA[oc] + moff, N, ti.c_row0, ti.c_row0 + (ti.row_end - ti.a_row0), ti.col0, wm, wn, g, t);
            }
        }
    }
    // peer stores need no fence here: the kernel boundary orders them before the signal kernel, whose system-scope fence
    // precedes the flag
}

}   // namespace

PFN_tmapEncodeTiled qf_tmap_encoder()
{
    static bool tried = false;
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && p &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_tmapEncodeTiled)p;
        else
            cudaGetLastError();
    }
    return fn;
}

struct QfGemmPlan {
    bool m3 = true;             // 3M (Karatsuba) arithmetic; QF_GEMM_3M=0 selects the 4-multiplication variant
    bool tma = true;
    bool cooperative = true;    // QF_GEMM_COOP=0 uses a plain launch
    bool warp_spec = true;      // QF_GEMM_WS=0: no producer warp (all 8 warps load through one elected thread)
    PFN_tmapEncodeTiled encode = nullptr;
    int max_ctas = 0;
    double2 *ws = nullptr;      // [max_ctas][32][256] partial tiles
    int *flags = nullptr;       // [max_ctas]
    // cached tile lists keyed by (upper_only, rank, nranks, a_permuted); rank < 0 = all ranks (single-GPU emulation)
    struct List { int upper, rank, nranks, aperm, natural, bn, ntiles; SkTile *dev; };
    std::vector<List> lists;
    int small_tiles = 1;        // QF_GEMM_SMALL=0: never use the 64 x 32 tile variant, 2: always (experiments)
    int BM() const { return m3 ? Cfg<1>::BM : Cfg<0>::BM; }
    int BN(bool small = false) const { return m3 ? (small ? Cfg<2>::BN : Cfg<1>::BN) : Cfg<0>::BN; }
    int BK() const { return m3 ? Cfg<1>::BK : Cfg<0>::BK; }
};

// The 64 x 32 tile variant pays off when the 64 x 64 tiles of a launch would occupy at most half of the SMs (N <= 512 for
// one simulation): twice as many tiles, every SM owns a whole tile (or a clean share of one) instead of a stream-K
// fragment.  Single GPU, warp-specialised 3M TMA kernel only.
static bool use_small_tiles(const qf_handle_s *h, bool upper_only, int nranks)
{
    const QfGemmPlan *p = h->gemm;
    if (!(p->small_tiles && p->m3 && p->tma && p->warp_spec && nranks == 1 && h->N >= 8)) return false;
    if (p->small_tiles == 2) return true;
    const int nt = (h->N + 63) / 64;
    const long long tiles64 = (long long)h->batch * (upper_only ? (long long)nt * (nt + 1) / 2 : (long long)nt * nt);
    return 2 * tiles64 <= p->max_ctas;
}

int qf_gemm_create(qf_handle_s *h)
{
    QfGemmPlan *p = new QfGemmPlan();
    h->gemm = p;
    const char *env = getenv("QF_GEMM_3M");
    p->m3 = !(env && env[0] == '0');
    p->max_ctas = h->sm_count;
    {
        const char *sm = getenv("QF_GEMM_SMALL");
        p->small_tiles = sm ? atoi(sm) : 1;
    }
    {
        const char *c = getenv("QF_GEMM_COOP");
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device);
        p->cooperative = coop && !(c && c[0] == '0');
    }
    QF_CUDA(cudaFuncSetAttribute(k_zgemm_sk<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<0>::SMEM));
    QF_CUDA(cudaFuncSetAttribute(k_zgemm_sk<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<0>::SMEM));
    QF_CUDA(cudaFuncSetAttribute(k_zgemm_sk<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<1>::SMEM));
    QF_CUDA(cudaFuncSetAttribute(k_zgemm_sk<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<1>::SMEM));
    QF_CUDA(cudaFuncSetAttribute(k_zgemm3m_ws<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<1>::SMEM));
    QF_CUDA(cudaFuncSetAttribute(k_zgemm3m_ws<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<1>::SMEM));
    QF_CUDA(cudaFuncSetAttribute(k_zgemm3m_ws<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<2>::SMEM));
    {
        const char *w = getenv("QF_GEMM_WS");
        p->warp_spec = !(w && w[0] == '0');
    }
    {
        const char *ld = getenv("QF_GEMM_LOAD");
        p->tma = !(ld && strcmp(ld, "cpasync") == 0);
        PFN_tmapEncodeTiled fn = qf_tmap_encoder();
        if (!fn) p->tma = false;   // no TMA descriptor encoder in this driver: keep the cp.async loader
        p->encode = fn;
    }
    QF_CUDA(cudaMalloc(&p->ws, sizeof(double2) * WS_D2_PER_THREAD * GEMM_THREADS * (size_t)p->max_ctas));
    QF_CUDA(cudaMalloc(&p->flags, sizeof(int) * p->max_ctas));
    QF_CUDA(cudaMemset(p->flags, 0, sizeof(int) * p->max_ctas));
    return QF_OK;
}

void qf_gemm_destroy(qf_handle_s *h)
{
    if (!h->gemm) return;
    for (auto &l : h->gemm->lists) cudaFree(l.dev);
    if (h->gemm->ws) cudaFree(h->gemm->ws);
    if (h->gemm->flags) cudaFree(h->gemm->flags);
    delete h->gemm;
    h->gemm = nullptr;
}

extern "C" int qf_gemm_is_3m(qf_handle_t h) { return h && h->gemm && h->gemm->m3 ? 1 : 0; }

// Executed real FP64 flops of one launch (for the roofline): tiles * BM * BN * N * (6 or 8).
extern "C" double qf_gemm_executed_flops(qf_handle_t h, int upper_only)
{
    if (!h || !h->gemm) return 0.0;
    const int N = h->N, BMv = h->gemm->BM(), BNv = h->gemm->BN(use_small_tiles(h, upper_only != 0, h->nranks));
    long long tiles = 0;
    for (int r0 = 0; r0 < N; r0 += BMv)
        for (int c0 = 0; c0 < N; c0 += BNv)
            if (!upper_only || c0 + BNv - 1 >= r0) ++tiles;
    const int BK = h->gemm->BK();
    const double kpad = (double)((N + BK - 1) / BK) * BK;
    return (double)tiles * BMv * BNv * kpad * (h->gemm->m3 ? 6.0 : 8.0) * h->batch;
}

// 3-D tensor map over `batch` row-major N x N complex128 matrices seen as doubles: dims {2N, N, batch}, box {16, rows, 1}
// (= 8 complex = one 128-byte swizzle span per row), SWIZZLE_128B, out-of-bounds elements read as zero.
static int make_tmap(qf_handle_s *h, const double2 *base, int box_rows, CUtensorMap *out)
{
    const cuuint64_t N = (cuuint64_t)h->N;
    cuuint64_t dims[3] = {2 * N, N, (cuuint64_t)h->batch};
    cuuint64_t strides[2] = {16 * N, 16 * N * N};
    cuuint32_t box[3] = {16, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = h->gemm->encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double2 *>(base), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        qf_set_error("cuTensorMapEncodeTiled failed with CUresult %d (N=%d)", (int)r, h->N);
        return QF_ERR_CUDA;
    }
    return QF_OK;
}

// natural: output rows keep their natural positions (single GPU: always true in effect; tile-exchange path); otherwise the
// output uses the rank-permuted row layout of the legacy all-gather paths (qf_prow).
static int get_tile_list(qf_handle_s *h, bool upper_only, int rank, int nranks, bool a_permuted, bool natural, const SkTile **dev,
                         int *ntiles, bool small = false)
{
    QfGemmPlan *p = h->gemm;
    if (nranks == 1) { natural = true; a_permuted = false; }      // one rank: the permutation is the identity
    const int BNv = p->BN(small);
    for (auto &l : p->lists)
        if (l.upper == (int)upper_only && l.rank == rank && l.nranks == nranks && l.aperm == (int)a_permuted &&
            l.natural == (int)natural && l.bn == BNv) {
            *dev = l.dev;
            *ntiles = l.ntiles;
            return QF_OK;
        }
    std::vector<SkTile> tl;
    const int N = h->N, BMv = p->BM();
    const int hb = qf_block_rows(N, nranks);
    for (int b = 0; b < h->batch; ++b)
        for (int r = 0; r < nranks; ++r) {
            if (rank >= 0 && r != rank) continue;
            // rank r owns logical row blocks r and 2G-1-r (balanced for the upper-triangular GEMM)
            const int blocks[2] = {r, 2 * nranks - 1 - r};
            for (int q = 0; q < (nranks == 1 ? 1 : 2); ++q) {
                const int rb = (nranks == 1) ? 0 : blocks[q] * hb;
                const int re = (nranks == 1) ? N : std::min(N, rb + hb);
                // Tile order = L2 blocking.  Column groups of CG tile-columns, row-major inside a group: the B panel of a
                // group (CG * BN columns, 33 MB at N=2048) stays L2-resident while the waves stream the A row panels past
                // it, so A is read ncols/CG times and B once (all 32 tile-columns at once re-reads B every wave).
                const int CG = 1024;      // complex columns per group
                for (int cg = 0; cg < N; cg += CG)
                    for (int r0 = rb; r0 < re; r0 += BMv)
                        for (int c0 = cg; c0 < std::min(N, cg + CG); c0 += BNv) {
                            if (upper_only && (c0 + BNv - 1 < r0)) continue;
                            tl.push_back(SkTile{b, r0, natural ? r0 : qf_prow(r0, hb, nranks), c0, std::min(re, r0 + BMv),
                                                a_permuted ? qf_prow(r0, hb, nranks) : r0, 0, 0});
                        }
            }
        }
    QfGemmPlan::List l{(int)upper_only, rank, nranks, (int)a_permuted, (int)natural, BNv, (int)tl.size(), nullptr};
    QF_CUDA(cudaMalloc(&l.dev, sizeof(SkTile) * std::max<size_t>(tl.size(), 1)));
    QF_CUDA(cudaMemcpy(l.dev, tl.data(), sizeof(SkTile) * tl.size(), cudaMemcpyHostToDevice));
    p->lists.push_back(l);
    *dev = l.dev;
    *ntiles = l.ntiles;
    return QF_OK;
}

int qf_gemm_prepare(qf_handle_s *h, int rank, int nranks)
{
    const SkTile *t;
    int n;
    if (h->comm_mode == 5 || nranks == 1) {
        QF_CHECK(get_tile_list(h, false, rank, nranks, false, true, &t, &n));
        QF_CHECK(get_tile_list(h, true, rank, nranks, false, true, &t, &n));    // fused GEMM-2 tail
        if (use_small_tiles(h, false, nranks)) QF_CHECK(get_tile_list(h, false, rank, nranks, false, true, &t, &n, true));
        if (use_small_tiles(h, true, nranks)) QF_CHECK(get_tile_list(h, true, rank, nranks, false, true, &t, &n, true));
    }
    if (h->comm_mode != 5) {
        QF_CHECK(get_tile_list(h, false, rank, nranks, false, false, &t, &n));
        QF_CHECK(get_tile_list(h, true, rank, nranks, true, false, &t, &n));
    }
    return QF_OK;
}

// Tile lists of the all-gather data path (rank-permuted output rows) for this handle's rank: used by the host-stepped
// driver on several GPUs whatever the handle's default data path is.
int qf_gemm_prepare_gather(qf_handle_s *h)
{
    const SkTile *t;
    int n;
    QF_CHECK(get_tile_list(h, false, h->rank, h->nranks, false, false, &t, &n));
    QF_CHECK(get_tile_list(h, true, h->rank, h->nranks, true, false, &t, &n));
    return QF_OK;
}

// The stream-K CTAs wait on one another (finisher <- contributors), so they must all be resident at the same time:
// the launch is cooperative, which makes the driver co-schedule the whole grid (grid <= #SMs, 1 CTA/SM).
template <int M3, bool TMA>
static cudaError_t launch_sk_impl(qf_handle_s *h, int G, const double2 *A, const double2 *B, double2 *C, const SkTile *tiles,
                                  int ntiles, int gated, const CUtensorMap &tmA, const CUtensorMap &tmB, cudaStream_t st)
{
    QfGemmPlan *p = h->gemm;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = Geo<M3>::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = p->cooperative ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k_zgemm_sk<M3, TMA>, A, B, C, h->N, tiles, ntiles, p->ws, p->flags,
                              (const QfCtrl *)h->ctrl, gated, tmA, tmB);
}

template <int M3>
static cudaError_t launch_sk(qf_handle_s *h, bool tma, int G, const double2 *A, const double2 *B, double2 *C, const SkTile *tiles,
                             int ntiles, int gated, const CUtensorMap &tmA, const CUtensorMap &tmB, cudaStream_t st)
{
    return tma ? launch_sk_impl<M3, true>(h, G, A, B, C, tiles, ntiles, gated, tmA, tmB, st)
               : launch_sk_impl<M3, false>(h, G, A, B, C, tiles, ntiles, gated, tmA, tmB, st);
}

// rank/nranks select the row blocks (see qf_prow); rank < 0 computes every rank's blocks (emulation on one GPU).
int qf_launch_zgemm(qf_handle_s *h, const double2 *A, const double2 *B, double2 *C, bool upper_only, bool gated,
                    int rank, int nranks, bool a_permuted, cudaStream_t st, bool natural, const QfXchg *xg)
{
    const int N = h->N;
    QfGemmPlan *p = h->gemm;
    const SkTile *tiles;
    int ntiles;
    const bool small = !xg && use_small_tiles(h, upper_only, nranks);
    QF_CHECK(get_tile_list(h, upper_only, rank, nranks, a_permuted, natural, &tiles, &ntiles, small));
    if (ntiles == 0) return QF_OK;
    const int BK = p->BK();
    const int KT = (N + BK - 1) / BK;
    const long long T = (long long)ntiles * KT;
    // at least 128 k per CTA so that the fix-up traffic stays small
    const int G = (int)std::min<long long>(p->max_ctas, std::max<long long>(1, T / (128 / BK)));
    // TMA needs 16-byte aligned rows (always true) and N >= 8 so that a 128-byte box fits the row pitch
    const bool tma = p->tma && N >= 8;
    CUtensorMap tmA, tmB;
    memset(&tmA, 0, sizeof(tmA));
    memset(&tmB, 0, sizeof(tmB));
    if (tma) {
        QF_CHECK(make_tmap(h, A, p->BM(), &tmA));
        QF_CHECK(make_tmap(h, B, BK, &tmB));   // B boxes: BK rows x 8 complex
    }
    if (xg && !(p->m3 && tma && p->warp_spec)) {
        qf_set_error("the tile-exchange data path needs the warp-specialised 3M TMA kernel (QF_GEMM_3M/QF_GEMM_WS/QF_GEMM_LOAD defaults)");
        return QF_ERR_UNSUPPORTED;
    }
    if (p->m3 && tma && p->warp_spec) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(G);
        cfg.blockDim = dim3(WS_THREADS);
        cfg.dynamicSmemBytes = small ? Geo<2>::SMEM : Geo<1>::SMEM;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.attrs = attr;
        cfg.numAttrs = p->cooperative ? 1 : 0;
        const QfEpiPost none = {};
        const QfXchg solo;
        if (small)
            QF_CUDA(cudaLaunchKernelEx(&cfg, k_zgemm3m_ws<2, false>, C, N, tiles, ntiles, p->ws, p->flags, (const QfCtrl *)h->ctrl,
                                       gated ? 1 : 0, tmA, tmB, none, solo));
        else
            QF_CUDA(cudaLaunchKernelEx(&cfg, k_zgemm3m_ws<1, false>, C, N, tiles, ntiles, p->ws, p->flags, (const QfCtrl *)h->ctrl,
                                       gated ? 1 : 0, tmA, tmB, none, xg ? *xg : solo));
    } else {
        QF_CUDA(p->m3 ? launch_sk<true>(h, tma, G, A, B, C, tiles, ntiles, gated ? 1 : 0, tmA, tmB, st)
                      : launch_sk<false>(h, tma, G, A, B, C, tiles, ntiles, gated ? 1 : 0, tmA, tmB, st));
    }
    h->launches++;
    QF_CUDA(cudaGetLastError());
    return QF_OK;
}

// Second GEMM of the fixed-point iteration with the fused tail (QfEpiPost): upper tiles of S = A P~ only, nothing is
// written for S itself.  Needs the warp-specialised 3M TMA kernel; the caller falls back to qf_launch_zgemm + k_post
// (QF_FUSE_POST=0) otherwise.
bool qf_gemm_can_fuse_post(qf_handle_s *h)
{
    QfGemmPlan *p = h->gemm;
    return p && p->m3 && p->tma && p->warp_spec && h->N >= 8;
}

int qf_launch_zgemm_post(qf_handle_s *h, const double2 *A, const double2 *B, const QfEpiPost &epi, bool gated, int rank,
                         int nranks, cudaStream_t st, const QfXchg *xg)
{
    const int N = h->N;
    QfGemmPlan *p = h->gemm;
    if (!qf_gemm_can_fuse_post(h)) { qf_set_error("the fused GEMM-2 tail needs the warp-specialised 3M TMA kernel"); return QF_ERR_UNSUPPORTED; }
    const SkTile *tiles;
    int ntiles;
    QF_CHECK(get_tile_list(h, true, rank, nranks, false, true, &tiles, &ntiles));
    if (ntiles == 0) return QF_OK;
    const int BK = p->BK();
    const int KT = (N + BK - 1) / BK;
    const long long T = (long long)ntiles * KT;
    const int G = (int)std::min<long long>(p->max_ctas, std::max<long long>(1, T / (128 / BK)));
    CUtensorMap tmA, tmB;
    QF_CHECK(make_tmap(h, A, p->BM(), &tmA));
    QF_CHECK(make_tmap(h, B, BK, &tmB));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G);
    cfg.blockDim = dim3(WS_THREADS);
    cfg.dynamicSmemBytes = Geo<1>::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = p->cooperative ? 1 : 0;
    double2 *nullC = nullptr;
    const QfXchg solo;
    QF_CUDA(cudaLaunchKernelEx(&cfg, k_zgemm3m_ws<1, true>, nullC, N, tiles, ntiles, p->ws, p->flags, (const QfCtrl *)h->ctrl,
                               gated ? 1 : 0, tmA, tmB, epi, xg ? *xg : solo));
    h->launches++;
    QF_CUDA(cudaGetLastError());
    return QF_OK;
}
