// Hoppe–Yau Laplacian on the device: P = eps * Delta_N^{-1} (W + dW), and W = Delta_N P.
//
// Replaces quflow/laplacian/cpu.py: `_compute_cpu_laplacian` (:55-95), `_solve_cpu_skewh`
// (:281-362), `solve_poisson` (:681-734), `laplace`/`_dot_cpu_generic` (:628-669, :98-108).
//
// Math.  Element (k, k+m) of the upper triangle is position k of the tridiagonal system of
// diagonal m (length N-m):  o_k x_{k-1} + d_k x_k + o_{k+1} x_{k+1} = r_k  with
//   d_k = -((N-1)(2k+1+m) - 2k(k+m)),   o_k = sqrt((k+m)(N-k-m) k (N-k)),   d_0 -= 1/2 for m = 0.
// The matrices do not depend on W, so their LU factors  w_k = o_k/u_{k-1}, u_k = d_k - w_k o_k
// are built once per N on the host (the reference recomputes them every call) and kept in HBM packed
// in the order the solve kernel reads them (k_poisson_band below).
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "qf_common.cuh"

int qf_poisson_prepare(qf_handle_s *h);

// ---------------------------------------------------------------------------------------
// host: the work plan of k_poisson_band (pure host code, also exported for the CPU tests: qf_poisson_plan)
// ---------------------------------------------------------------------------------------
// A band is M adjacent diagonals m = M b + s.  The triangle is cut into UNITS of PC positions x M diagonals, one per
// CTA: a "long piece" (positions [posbase, posbase + PC) of band bL) followed, from local position PS on, by a whole
// short band bS that fills the space the long piece leaves (diagonal lengths fall linearly, so band b and band
// nbands - b together have about N positions: every CTA is full).  A band longer than PC spans `nlink` consecutive
// ranks of one thread-block cluster, starting at rank `rank0`; several such groups and independent units share a
// cluster.  Returns false when N needs more than 8 CTAs per band (no cluster that large).
static bool qf_poisson_plan_host(int N, int &L, int &M, int &NT, int &CL, std::vector<int> &units)
{
    // Measured at N = 2048: 16 positions per thread, 4 diagonals per band and CTAs of up to 512
    // threads are the fastest combination (8 positions: +13 us; 8 diagonals: +7 us; 256-thread cluster pairs: +3 us).
    // Up to N = 1024 a band fits one CTA even with 8 positions per thread, and the solve is a pure latency chain there:
    // half the sequential depth per thread wins (measured 14.2 -> 12.5 us at N = 512, 19.3 -> 17.0 us at N = 1024).
    L = (N <= 1024) ? 8 : 16;
    {
        const char *env = getenv("QF_POISSON_L");          // 8 or 16: override for experiments
        if (env && (atoi(env) == 8 || atoi(env) == 16)) L = atoi(env);
    }
    M = 4;
    const int NTMAX = 512;
    const int chunks = (N + L - 1) / L;
    NT = NTMAX;
    CL = 1;
    if (M * chunks <= NTMAX) {
        NT = ((M * chunks + 31) / 32) * 32;
    } else {
        const int need = (M * chunks + NTMAX - 1) / NTMAX;
        while (CL < need) CL *= 2;
    }
    if (CL > 8) return false;
    const int PC = (NT / M) * L;
    const int nbands = (N + M - 1) / M;
    std::vector<char> taken(nbands, 0);
    units.clear();
    auto filler = [&](int used) -> int {      // longest free band that fits into PC - used positions
        const int room = PC - used;
        if (room <= 0) return -1;
        int b0 = (N - room + M - 1) / M;
        if (b0 < 0) b0 = 0;
        for (int bb = b0; bb < nbands; ++bb)
            if (!taken[bb] && N - M * bb <= room) return bb;
        return -1;
    };
    auto push_single = [&](int bb) {
        taken[bb] = 1;
        const int used = N - M * bb;
        int bS = filler(used), PS = PC;
        if (bS >= 0) { taken[bS] = 1; PS = used; }
        units.insert(units.end(), {bb, 0, bS, PS, 1, 0, 0, 0});
    };
    // Linked bands (longer than PC) take k = ceil(len / PC) consecutive ranks of a cluster; several of them share a
    // cluster when they fit (largest first, then the largest that still fits), the remaining ranks take single bands.
    std::vector<int> linked_bands;
    for (int bb = 0; bb < nbands && N - M * bb > PC; ++bb) {
        linked_bands.push_back(bb);
        taken[bb] = 1;
    }
    std::vector<char> placed(linked_bands.size(), 0);
    for (size_t g0 = 0; g0 < linked_bands.size(); ++g0) {
        if (placed[g0]) continue;
        int room = CL;
        for (size_t g = g0; g < linked_bands.size() && room >= 2; ++g) {
            if (placed[g]) continue;
            const int bb = linked_bands[g];
            const int len = N - M * bb;
            const int k = (len + PC - 1) / PC;
            if (k > room) continue;
            placed[g] = 1;
            const int rank0 = CL - room;
            for (int r = 0; r < k; ++r) {
                int bS = -1, PS = PC;
                if (r == k - 1) {
                    const int used = len - r * PC;
                    bS = filler(used);
                    if (bS >= 0) { taken[bS] = 1; PS = used; }
                }
                units.insert(units.end(), {bb, r * PC, bS, PS, k, rank0, 0, 0});
            }
            room -= k;
        }
        for (; room > 0; --room) {                                     // spare ranks: independent short bands
            int nb = -1;
            for (int q = nbands - 1; q >= 0; --q) if (!taken[q] && N - M * q <= PC) nb = q;   // longest free short band
            if (nb >= 0) push_single(nb); else units.insert(units.end(), {-1, 0, -1, PC, 1, 0, 0, 0});
        }
    }
    for (int bb = 0; bb < nbands; ++bb)                                // the rest: one band (+ filler) per CTA
        if (!taken[bb]) push_single(bb);
    while ((units.size() / 8) % CL) units.insert(units.end(), {-1, 0, -1, PC, 1, 0, 0, 0});
    return true;
}

// params_out[6] = L, M, NT (threads per CTA), CL (CTAs per cluster), PC (positions per CTA), number of units;
// units_out receives 8 ints per unit (bL, posbase, bS, PS, nlink, rank0, 0, 0) if it has room for them (cap ints).
// Returns the number of units, 0 if N needs more than 8 CTAs per band (unsupported), or a negative qf_status.  No CUDA call.
extern "C" int qf_poisson_plan(int N, int *params_out, int *units_out, int cap)
{
    if (N < 2 || !params_out) { qf_set_error("qf_poisson_plan: bad arguments"); return QF_ERR_INVALID; }
    int L, M, NT, CL;
    std::vector<int> units;
    if (!qf_poisson_plan_host(N, L, M, NT, CL, units)) return 0;
    const int nunits = (int)(units.size() / 8);
    const int p[6] = {L, M, NT, CL, (NT / M) * L, nunits};
    memcpy(params_out, p, sizeof(p));
    if (units_out && cap >= (int)units.size()) memcpy(units_out, units.data(), units.size() * sizeof(int));
    return nunits;
}

// ---------------------------------------------------------------------------------------
// host: coefficient / factor tables
// ---------------------------------------------------------------------------------------
int qf_build_tables(qf_handle_s *h)
{
    const int N = h->N;
    const size_t n2 = (size_t)N * N;
    std::vector<double> tw(n2, 0.0), tiu(n2, 0.0);
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
            u_prev = u;
        }
    }
    // ---- unit-packed tables for k_poisson_band (work plan: qf_poisson_plan_host above).  The factor entries of a unit
    // are contiguous and ordered [warp block][i][chunk in block][s]: the 32 lanes of a warp read 32 consecutive doubles
    // for every i.  Entries outside the diagonals are 0 (which also decouples the pieces of a unit).
    {
        int L, M, NT, CL;
        std::vector<int> units;     // 8 ints per unit: bL, posbase, bS, PS, nlink, rank0, 0, 0
        if (!qf_poisson_plan_host(N, L, M, NT, CL, units)) {
            qf_set_error("N=%d needs more than 8 CTAs per band of diagonals: not supported", N);
            return QF_ERR_UNSUPPORTED;
        }
        const int NTMAX = 512;
        const int PC = (NT / M) * L, WB = (32 / M) * L;
        const size_t nunits = units.size() / 8;
        const size_t total = nunits * (size_t)PC * M;
        if (total >= (1ull << 31)) { qf_set_error("factor tables of N=%d exceed 2^31 entries", N); return QF_ERR_UNSUPPORTED; }
        const int CPW = 32 / M;
        std::vector<double> pw(total, 0.0), piu(total, 0.0);
        for (size_t u = 0; u < nunits; ++u) {
            const int bL = units[8 * u], posbase = units[8 * u + 1], bS = units[8 * u + 2], PS = units[8 * u + 3];
            for (int sl = 0; sl < M; ++sl) {
                for (int pl = 0; pl < PC; ++pl) {
                    int m = -1, k = 0;
                    if (pl < PS) {
                        if (bL >= 0) { m = M * bL + sl; k = posbase + pl; }
                    } else if (bS >= 0) {
                        m = M * bS + sl;
                        k = pl - PS;
                    }
                    if (m < 0 || m >= N || k >= N - m) continue;
                    const size_t src = (size_t)k * N + (k + m);
                    const size_t dst = u * (size_t)PC * M + (size_t)(pl / WB) * WB * M + (size_t)(pl % L) * 32 +
                                       (size_t)((pl / L) % CPW) * M + sl;
                    pw[dst] = tw[src];
                    piu[dst] = tiu[src];
                }
            }
        }
        QF_CUDA(cudaMalloc(&h->ptab_w, total * sizeof(double)));
        QF_CUDA(cudaMalloc(&h->ptab_iu, total * sizeof(double)));
        QF_CUDA(cudaMalloc(&h->ptab_units, units.size() * sizeof(int)));
        QF_CUDA(cudaMemcpy(h->ptab_w, pw.data(), total * sizeof(double), cudaMemcpyHostToDevice));
        QF_CUDA(cudaMemcpy(h->ptab_iu, piu.data(), total * sizeof(double), cudaMemcpyHostToDevice));
        QF_CUDA(cudaMemcpy(h->ptab_units, units.data(), units.size() * sizeof(int), cudaMemcpyHostToDevice));
        h->p_L = L;
        h->p_M = M;
        h->p_CL = CL;
        h->p_NT = NT;
        h->p_NTMAX = NTMAX;
        h->p_nunits = (int)nunits;
    }
    return qf_poisson_prepare(h);
}

// ---------------------------------------------------------------------------------------
// helpers
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

// ---------------------------------------------------------------------------------------
// The solve: band kernel (chunked affine scan).
//
// With the LDL^T factors (w_k, 1/u_k) both sweeps are first-order linear recurrences,
//     forward   c_k = r_k - w_k c_{k-1}                 backward  x_k = c_k/u_k - w_{k+1} x_{k+1},
// so a chunk of L positions is an affine map of its carry-in.  Thread (s, c) keeps chunk c of diagonal s of its
// band in registers: pass 1 evaluates the chunk with carry-in 0 plus the map's slope (a running product), a
// warp-shuffle + shared-memory scan composes the maps along the diagonal, pass 2 re-runs the chunk from its true
// carry-in.  Sequential depth: 4 L FMAs + two log-depth scans instead of 2 N.
//   * a CTA is up to 512 threads = M diagonals x chunks of L = 16 positions (M = 4: 2048 positions);
//   * the work is cut into FULL units (qf_build_tables): a long piece of one band plus a whole short band, so
//     the triangle costs N^2/2 thread slots, not N^2; warps past the end of their pieces do nothing;
//   * a band longer than one CTA's positions is solved by a thread-block CLUSTER: every CTA reduces its part
//     of the diagonal to an affine map; the maps travel through distributed shared memory with st.async, which
//     signals an mbarrier in the receiving CTA.  No cluster-wide barrier and no memory fence sits in the solve
//     (a fence would wait for every load and prefetch the CTA has in flight);
//   * factor tables are unit-packed: every warp load is 256 contiguous bytes; 1/u goes straight to shared
//     memory with cp.async and is consumed between the sweeps;
//   * the skew-Hermitian mirror P[j,i] = -conj(P[i,j]) is transposed through shared memory, so the mirrored
//     stores are 16 M-byte row pieces like the direct ones (no scattered 16-byte stores);
//   * once its own loads have landed a CTA prefetches into L2 the inputs of the unit that will follow it on its
//     SM slot (pf_stride units ahead), so HBM keeps streaming while the resident CTAs run their sweeps.
// HBM traffic: upper triangle of W~ (8 N^2 B) + w and 1/u (8 N^2 B) + full P (16 N^2 B) = 32 N^2 B.
// ---------------------------------------------------------------------------------------
#ifdef QF_PTRACE
__device__ unsigned long long g_ptrace[4096 * 16];
#define PT(k) do { if (tid == 0 && mem == 0 && blockIdx.x < 4096) { unsigned long long _t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t)); g_ptrace[blockIdx.x * 16 + (k)] = _t; } } while (0)
#else
#define PT(k)
#endif

__device__ __forceinline__ void pb_mbar_wait(uint32_t mb)
{
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(mb), "r"(0u) : "memory");
}
// remote store of one double into CTA `dst` of the cluster; completes 8 bytes on that CTA's mbarrier
__device__ __forceinline__ void pb_send(double *local_slot, uint64_t *local_mbar, unsigned dst, double v)
{
    uint32_t ra, rm;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"((uint32_t)__cvta_generic_to_shared(local_slot)), "r"(dst));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rm) : "r"((uint32_t)__cvta_generic_to_shared(local_mbar)), "r"(dst));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(ra), "d"(v), "r"(rm) : "memory");
}

// L2 prefetch of the inputs of unit `un` with this CTA's thread mapping: the factor tables as two bulk requests, the
// row pieces of W~ one request per end of a piece (fire and forget: no data comes back to the SM, so these requests do
// not occupy the L1 miss queue that throttles ordinary loads to about 30 GB/s per SM at DRAM latency).
template <int L, int M>
__device__ __forceinline__ void pb_prefetch_unit(int un, const int4 *__restrict__ units, const double *__restrict__ tw,
                                                 const double *__restrict__ tiu, const double2 *__restrict__ R, int N, int PC,
                                                 int tid, int s, int plo)
{
    const unsigned stride = (unsigned)N + 1u;
    const int4 nd = __ldg(units + 2 * un);
    if (tid < 2) {
        const double *t = (tid == 0 ? tw : tiu) + (size_t)un * PC * M;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(t), "r"(PC * M * 8) : "memory");
    }
    if (s == 0 || s == M - 1) {
        const int nmL = nd.x * M + s, nmS = nd.z * M + s;
        const int nnL = (nd.x >= 0) ? min(nd.w, max(0, N - nmL - nd.y)) : 0;
        const int nnS = (nd.z >= 0) ? max(0, N - nmS) : 0;
        const unsigned noL = (unsigned)nd.y * stride + (unsigned)nmL;
        const unsigned noS = (unsigned)nmS - (unsigned)nd.w * stride;
#pragma unroll
        for (int i = 0; i < L; ++i) {
            const int pl = plo + i;
            const bool sh = pl >= nd.w;
            const bool ok = sh ? (pl - nd.w < nnS) : (pl < nnL);
            if (ok) asm volatile("prefetch.global.L2 [%0];" ::"l"(R + ((unsigned)pl * stride + (sh ? noS : noL))));
        }
    }
}

template <int L, int M, int CL, int NTMAX>
__global__ void __launch_bounds__(NTMAX, 512 / NTMAX)
k_poisson_band(const double2 *__restrict__ Wh, double2 *__restrict__ P, const double *__restrict__ tw,
               const double *__restrict__ tiu, const int4 *__restrict__ units, int N, int nunits, int pf_stride,
               double eps, const QfCtrl *__restrict__ ctrl, int gated)
{
    const int mem = blockIdx.y;
    if (gated && !ctrl[mem].active) return;      // uniform over the grid row: whole clusters leave together
    constexpr int LOGM = (M == 8) ? 3 : 2;
    constexpr int CPW = 32 / M;                  // chunks of one diagonal per warp
    constexpr int WB = CPW * L;                  // positions per warp
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    const int PC = (blockDim.x >> LOGM) * L;     // positions per CTA
    const int unit = blockIdx.x;
    const int rank = (CL > 1) ? unit % CL : 0;   // cluster dims are (CL, 1, 1): this is %cluster_ctarank
    const int4 ud = __ldg(units + 2 * unit);
    const int bL = ud.x, posbase = ud.y, bS = ud.z, PS = ud.w;
    const int4 ue = (CL > 1) ? __ldg(units + 2 * unit + 1) : make_int4(1, 0, 0, 0);
    const int nlink = ue.x;                      // ranks of the cluster that share this CTA's long band ...
    const int rank0 = ue.y;                      // ... starting at this cluster rank
    const int grank = rank - rank0;              // this CTA's place among them
    const bool linked = (CL > 1) && nlink > 1;
    const bool diag0 = (bL == 0);                // the band of the main diagonal (all its ranks see bL == 0)

    __shared__ double totA[NTMAX / 32][M], totBx[NTMAX / 32][M], totBy[NTMAX / 32][M];
    __shared__ double xchF[CL][M][3], xchB[CL][M][3];     // per-rank chunk maps, written by the peers (DSMEM)
    __shared__ double redx[NTMAX / 32], redy[NTMAX / 32];
    __shared__ double sumR[CL][2], sumX[CL][2];
    __shared__ uint64_t mbar[4];                          // arrival of xchF, xchB, sumR, sumX (each used once)
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    double *iu_s = reinterpret_cast<double *>(dyn_smem);     // [L][blockDim.x], consumed by the forward sweep
    double2 *tr_s = reinterpret_cast<double2 *>(dyn_smem);   // mirror transpose buffer (reuses the same bytes later)

    if (CL > 1) {
        if (tid == 0) {
            const uint32_t nF = linked ? grank * M * 24 : 0, nB = linked ? (nlink - 1 - grank) * M * 24 : 0;
            const uint32_t nS = (linked && diag0) ? (nlink - 1) * 16 : 0;
            const uint32_t tx[4] = {nF, nB, nS, nS};
            for (int q = 0; q < 4; ++q) {
                const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&mbar[q]);
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(tx[q]) : "memory");
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        // every CTA of the cluster arrives here (nothing is in flight yet); the wait sits before the first send
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    }
    if (bL < 0 && bS < 0) return;                // padding unit
#ifdef QF_PTRACE
    if (tid == 0 && mem == 0 && blockIdx.x < 4096) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); g_ptrace[blockIdx.x * 16 + 15] = sm; }
#endif
    PT(0);

    const int s = tid & (M - 1), c = tid >> LOGM;
    const int plo = c * L;                         // first local position of this thread's chunk
    const unsigned stride = (unsigned)N + 1u;
    // long piece: system m = M bL + s, positions posbase + pl, pl < nL;  short piece: m' = M bS + s, positions pl - PS < nS
    const int mL = bL * M + s, mS = bS * M + s;
    const int nL = (bL >= 0) ? min(PS, max(0, N - mL - posbase)) : 0;
    const int nS = (bS >= 0) ? max(0, N - mS) : 0;
    // element (k, k+m) sits at k (N+1) + m, its mirror (k+m, k) at k (N+1) + m N; k = posbase + pl or pl - PS
    const unsigned offL = (unsigned)posbase * stride + (unsigned)mL;
    const unsigned offS = (unsigned)mS - (unsigned)PS * stride;
    const unsigned offML = (unsigned)posbase * stride + (unsigned)mL * (unsigned)N;
    const unsigned offMS = (unsigned)mS * (unsigned)N - (unsigned)PS * stride;
    const int extent = max((bL >= 0) ? min(PS, max(0, N - bL * M - posbase)) : 0, (bS >= 0) ? PS + N - bS * M : 0);
    const bool warp_work = warp * WB < extent;     // slot 0 is the longest thing a warp owns
    const bool in_short = plo >= PS;               // a chunk that starts in the short piece lies entirely in it
    const bool straddle = !in_short && plo + L > PS;
    // number of leading valid positions of a chunk that does not straddle the two pieces
    const int nvalid = straddle ? 0 : min(L, max(0, in_short ? nS - (plo - PS) : nL - plo));
    const unsigned offD = in_short ? offS : offL;
    const size_t off = (size_t)mem * N * N;
    const double2 *R = Wh + off;
    double2 *X = P + off;

    double2 r[L];
    double w[L + 1];                               // w[L]: w of the first position after the chunk
    const uint32_t iu_base = (uint32_t)__cvta_generic_to_shared(iu_s) + (uint32_t)tid * 8u;
    const uint32_t iu_pitch = blockDim.x * 8u;
    if (warp_work) {
        const size_t tb = (size_t)unit * PC * M + (size_t)warp * (WB * M) + lane;
        const double *wp = tw + tb;
        const double *up = tiu + tb;
#pragma unroll
        for (int i = 0; i < L; ++i) {
            w[i] = __ldg(wp + i * 32);             // zero outside the diagonals
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(iu_base + i * iu_pitch), "l"(up + i * 32));
        }
        if (nvalid == L) {
            const double2 *rp = R + ((unsigned)plo * stride + offD);
#pragma unroll
            for (int i = 0; i < L; ++i) r[i] = rp[(size_t)i * stride];
        } else {
#pragma unroll
            for (int i = 0; i < L; ++i) {
                const int pl = plo + i;
                const bool sh = pl >= PS;
                const bool ok = sh ? (pl - PS < nS) : (pl < nL);
                r[i] = ok ? R[(unsigned)pl * stride + (sh ? offS : offL)] : make_double2(0.0, 0.0);
            }
        }
        // w of the position after the chunk: next chunk of this unit, or the first chunk of the next linked rank
        w[L] = 0.0;
        if (plo + L < PC) {
            const int cn = c + 1;
            w[L] = __ldg(tw + (size_t)unit * PC * M + (size_t)(cn / CPW) * (WB * M) + (cn % CPW) * M + s);
        } else if (linked && grank + 1 < nlink) {
            w[L] = __ldg(tw + (size_t)(unit + 1) * PC * M + s);
        }
    } else {
#pragma unroll
        for (int i = 0; i < L; ++i) { r[i] = make_double2(0.0, 0.0); w[i] = 0.0; }
        w[L] = 0.0;
    }
    asm volatile("cp.async.commit_group;");
    PT(1);

    // ---- m = 0: remove the mean of the diagonal from the right-hand side (cpu.py:311-317,327-328)
    if (diag0) {
        double sx = 0.0, sy = 0.0;
        if (s == 0) {
#pragma unroll
            for (int i = 0; i < L; ++i)
                if (plo + i < nL) { sx += r[i].x; sy += r[i].y; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, o);
            sy += __shfl_xor_sync(0xffffffffu, sy, o);
        }
        if (lane == 0) { redx[warp] = sx; redy[warp] = sy; }
        if (linked) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");   // peers' mbarriers are ready
        __syncthreads();
        if (tid == 0) {
            double ax = 0.0, ay = 0.0;
            for (int q = 0; q < nwarps; ++q) { ax += redx[q]; ay += redy[q]; }
            sumR[grank][0] = ax;
            sumR[grank][1] = ay;
            if (linked) {
                for (int d = 0; d < nlink; ++d) {
                    if (d == grank) continue;
                    pb_send(&sumR[grank][0], &mbar[2], rank0 + d, ax);
                    pb_send(&sumR[grank][1], &mbar[2], rank0 + d, ay);
                }
            }
        }
        __syncthreads();
        if (linked) pb_mbar_wait((uint32_t)__cvta_generic_to_shared(&mbar[2]));
        double tx = 0.0, ty = 0.0;
        for (int d = 0; d < nlink; ++d) { tx += sumR[d][0]; ty += sumR[d][1]; }
        tx /= N;
        ty /= N;
        if (s == 0) {
#pragma unroll
            for (int i = 0; i < L; ++i)
                if (plo + i < nL) { r[i].x -= tx; r[i].y -= ty; }
        }
    } else if (linked) {
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");               // peers' mbarriers are ready
    }

    // ---- forward, pass 1: chunk map  c_out = A c_in + B
    double A = 1.0;
    double2 B = make_double2(0.0, 0.0);
    double xA = 1.0, xBx = 0.0, xBy = 0.0;
    if (warp_work) {
#pragma unroll
        for (int i = 0; i < L; ++i) {
            B.x = r[i].x - w[i] * B.x;
            B.y = r[i].y - w[i] * B.y;
            A = -w[i] * A;
        }
        // inclusive scan over the chunks of this warp (lanes with equal s are M apart)
#pragma unroll
        for (int d = M; d < 32; d <<= 1) {
            const double eA = __shfl_up_sync(0xffffffffu, A, d);
            const double eBx = __shfl_up_sync(0xffffffffu, B.x, d);
            const double eBy = __shfl_up_sync(0xffffffffu, B.y, d);
            if (lane >= d) {
                B.x = A * eBx + B.x;
                B.y = A * eBy + B.y;
                A = A * eA;
            }
        }
        xA = __shfl_up_sync(0xffffffffu, A, M);
        xBx = __shfl_up_sync(0xffffffffu, B.x, M);
        xBy = __shfl_up_sync(0xffffffffu, B.y, M);
        if (lane < M) { xA = 1.0; xBx = 0.0; xBy = 0.0; }
    }
    PT(2);
    // ---- L2 prefetch of the unit pf_stride ahead: it will follow this one on the same SM slot
    if (pf_stride > 0 && unit + pf_stride < nunits)
        pb_prefetch_unit<L, M>(unit + pf_stride, units, tw, tiu, R, N, PC, tid, s, plo);
    if (lane >= 32 - M) { totA[warp][s] = A; totBx[warp][s] = B.x; totBy[warp][s] = B.y; }
    __syncthreads();
    PT(3);
    if (linked) {
        if (tid < M && grank + 1 < nlink) {        // the map of this CTA's whole part of diagonal s -> higher ranks
            double tA = 1.0, tBx = 0.0, tBy = 0.0;
            for (int q = 0; q < nwarps; ++q) {
                const double a = totA[q][s];
                tBx = a * tBx + totBx[q][s];
                tBy = a * tBy + totBy[q][s];
                tA = a * tA;
            }
            for (int d = grank + 1; d < nlink; ++d) {
                pb_send(&xchF[grank][s][0], &mbar[0], rank0 + d, tA);
                pb_send(&xchF[grank][s][1], &mbar[0], rank0 + d, tBx);
                pb_send(&xchF[grank][s][2], &mbar[0], rank0 + d, tBy);
            }
        }
        if (grank > 0) pb_mbar_wait((uint32_t)__cvta_generic_to_shared(&mbar[0]));
    }
    PT(4);
    double2 carry;
    {
        double pBx = 0.0, pBy = 0.0;
        if (linked) {
            for (int d = 0; d < grank; ++d) {
                const double a = xchF[d][s][0];
                pBx = a * pBx + xchF[d][s][1];
                pBy = a * pBy + xchF[d][s][2];
            }
        }
        for (int q = 0; q < warp; ++q) {
            const double a = totA[q][s];
            pBx = a * pBx + totBx[q][s];
            pBy = a * pBy + totBy[q][s];
        }
        carry.x = xA * pBx + xBx;
        carry.y = xA * pBy + xBy;
    }
    // ---- forward, pass 2 from the true carry-in; r becomes z = c / u
    asm volatile("cp.async.wait_group 0;" ::: "memory");   // this thread's own 1/u values (no cross-thread sharing)
    if (warp_work) {
#pragma unroll
        for (int i = 0; i < L; ++i) {
            carry.x = r[i].x - w[i] * carry.x;
            carry.y = r[i].y - w[i] * carry.y;
            const double iu = iu_s[i * blockDim.x + tid];
            r[i].x = carry.x * iu;
            r[i].y = carry.y * iu;
        }
    }
    PT(5);
    __syncthreads();   // tot* and the iu_s bytes are reused below
    PT(6);

    // ---- backward, pass 1: x_out = A x_in + B  (x_in = value just after the chunk)
    A = 1.0;
    B = make_double2(0.0, 0.0);
    xA = 1.0; xBx = 0.0; xBy = 0.0;
    if (warp_work) {
#pragma unroll
        for (int i = L - 1; i >= 0; --i) {
            B.x = r[i].x - w[i + 1] * B.x;
            B.y = r[i].y - w[i + 1] * B.y;
            A = -w[i + 1] * A;
        }
#pragma unroll
        for (int d = M; d < 32; d <<= 1) {
            const double eA = __shfl_down_sync(0xffffffffu, A, d);
            const double eBx = __shfl_down_sync(0xffffffffu, B.x, d);
            const double eBy = __shfl_down_sync(0xffffffffu, B.y, d);
            if (lane + d < 32) {
                B.x = A * eBx + B.x;
                B.y = A * eBy + B.y;
                A = A * eA;
            }
        }
        xA = __shfl_down_sync(0xffffffffu, A, M);
        xBx = __shfl_down_sync(0xffffffffu, B.x, M);
        xBy = __shfl_down_sync(0xffffffffu, B.y, M);
        if (lane >= 32 - M) { xA = 1.0; xBx = 0.0; xBy = 0.0; }
    }
    PT(7);
    if (lane < M) { totA[warp][s] = A; totBx[warp][s] = B.x; totBy[warp][s] = B.y; }
    __syncthreads();
    PT(8);
    if (linked) {
        if (tid < M && grank > 0) {                // -> lower ranks
            double tA = 1.0, tBx = 0.0, tBy = 0.0;
            for (int q = nwarps - 1; q >= 0; --q) {
                const double a = totA[q][s];
                tBx = a * tBx + totBx[q][s];
                tBy = a * tBy + totBy[q][s];
                tA = a * tA;
            }
            for (int d = 0; d < grank; ++d) {
                pb_send(&xchB[grank][s][0], &mbar[1], rank0 + d, tA);
                pb_send(&xchB[grank][s][1], &mbar[1], rank0 + d, tBx);
                pb_send(&xchB[grank][s][2], &mbar[1], rank0 + d, tBy);
            }
        }
        if (grank + 1 < nlink) pb_mbar_wait((uint32_t)__cvta_generic_to_shared(&mbar[1]));
    }
    PT(9);
    {
        double pBx = 0.0, pBy = 0.0;
        if (linked) {
            for (int d = nlink - 1; d > grank; --d) {
                const double a = xchB[d][s][0];
                pBx = a * pBx + xchB[d][s][1];
                pBy = a * pBy + xchB[d][s][2];
            }
        }
        for (int q = nwarps - 1; q > warp; --q) {
            const double a = totA[q][s];
            pBx = a * pBx + totBx[q][s];
            pBy = a * pBy + totBy[q][s];
        }
        carry.x = xA * pBx + xBx;
        carry.y = xA * pBy + xBy;
    }
    // ---- backward, pass 2; r becomes x
    if (warp_work) {
#pragma unroll
        for (int i = L - 1; i >= 0; --i) {
            carry.x = r[i].x - w[i + 1] * carry.x;
            carry.y = r[i].y - w[i + 1] * carry.y;
            r[i] = carry;
        }
    }

    // ---- m = 0: remove the mean of diag(P) (cpu.py:342-352)
    if (diag0) {
        double sx = 0.0, sy = 0.0;
        if (s == 0) {
#pragma unroll
            for (int i = 0; i < L; ++i)
                if (plo + i < nL) { sx += r[i].x; sy += r[i].y; }
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
            sumX[grank][0] = ax;
            sumX[grank][1] = ay;
            if (linked) {
                for (int d = 0; d < nlink; ++d) {
                    if (d == grank) continue;
                    pb_send(&sumX[grank][0], &mbar[3], rank0 + d, ax);
                    pb_send(&sumX[grank][1], &mbar[3], rank0 + d, ay);
                }
            }
        }
        __syncthreads();
        if (linked) pb_mbar_wait((uint32_t)__cvta_generic_to_shared(&mbar[3]));
        double tx = 0.0, ty = 0.0;
        for (int d = 0; d < nlink; ++d) { tx += sumX[d][0]; ty += sumX[d][1]; }
        tx /= N;
        ty /= N;
        if (s == 0) {
#pragma unroll
            for (int i = 0; i < L; ++i)
                if (plo + i < nL) { r[i].x -= tx; r[i].y -= ty; }
        }
    }

    PT(10);
    // ---- store P = eps x (cpu.py:334; isospectral.py:492) and stage it for the mirror
    // tr_s[(pl, s)] with pl the local position; 64 bytes of padding per chunk keep both the chunk-major writes
    // and the row-major reads below free of bank conflicts.
    if (warp_work) {
        double2 *ts = tr_s + (plo * M + s + c * 4);
        if (nvalid == L) {
            double2 *xp = X + ((unsigned)plo * stride + offD);
#pragma unroll
            for (int i = 0; i < L; ++i) {
                const double2 v = make_double2(eps * r[i].x, eps * r[i].y);
                xp[(size_t)i * stride] = v;
                ts[i * M] = v;
            }
        } else {
#pragma unroll
            for (int i = 0; i < L; ++i) {
                const int pl = plo + i;
                const bool sh = pl >= PS;
                const bool ok = sh ? (pl - PS < nS) : (pl < nL);
                const double2 v = make_double2(eps * r[i].x, eps * r[i].y);
                if (ok) X[(unsigned)pl * stride + (sh ? offS : offL)] = v;
                ts[i * M] = v;
            }
        }
    }
    PT(11);
    __syncthreads();
    PT(12);
    // ---- mirror P[k+m, k] = -conj(P[k, k+m]) (cpu.py:340).  Row j of the mirror block holds the positions
    // pl = j - s of the M diagonals in columns j - s: M lanes write M * 16 contiguous bytes.
    if (warp_work) {
#pragma unroll
        for (int i = 0; i < L; ++i) {
            const int pl = plo + i - s;
            const bool sh = pl >= PS;
            const bool ok = pl >= 0 && (sh ? (pl - PS < nS) : (pl < nL)) && (sh ? mS : mL) != 0;
            if (ok) {
                const double2 v = tr_s[pl * M + s + (pl / L) * 4];
                X[(unsigned)pl * stride + (sh ? offMS : offML)] = make_double2(-v.x, v.y);
            }
        }
    }
    if (tid < M * (M - 1)) {   // rows PC .. PC+M-2 of the mirror block hold the last positions of diagonals s >= 1
        const int pl = PC + (tid >> LOGM) - s;
        const bool sh = pl >= PS;
        const bool ok = pl < PC && (sh ? (pl - PS < nS) : (pl < nL)) && (sh ? mS : mL) != 0;
        if (ok) {
            const double2 v = tr_s[pl * M + s + (pl / L) * 4];
            X[(unsigned)pl * stride + (sh ? offMS : offML)] = make_double2(-v.x, v.y);
        }
    }
    PT(13);
}

#ifdef QF_PTRACE
extern "C" int qf_ptrace_read(unsigned long long *out, int n)
{
    return (int)cudaMemcpyFromSymbol(out, g_ptrace, sizeof(unsigned long long) * (size_t)n);
}
#endif

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
static void *poisson_band_fn(const qf_handle_s *h)
{
    void *fn = nullptr;
#define QF_PB(LL, CC) if (h->p_L == LL && h->p_M == 4 && h->p_CL == CC && h->p_NTMAX == 512) fn = (void *)k_poisson_band<LL, 4, CC, 512>;
    QF_PB(16, 1) QF_PB(16, 2) QF_PB(16, 4) QF_PB(16, 8)
    QF_PB(8, 1) QF_PB(8, 2) QF_PB(8, 4) QF_PB(8, 8)
#undef QF_PB
    return fn;
}

// Per-handle (= per-device) kernel set-up: the opt-in to more than 48 KB of dynamic shared memory is a property of the
// function ON THE CURRENT DEVICE, so it is made once for every handle, right after its tables are built.
int qf_poisson_prepare(qf_handle_s *h)
{
    void *fn = poisson_band_fn(h);
    if (!fn) { qf_set_error("no k_poisson_band instantiation for L=%d M=%d CL=%d", h->p_L, h->p_M, h->p_CL); return QF_ERR_UNSUPPORTED; }
    QF_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, h->p_NTMAX == 512 ? 144 * 1024 : 72 * 1024));
    const char *env = getenv("QF_POISSON_PF");
    h->p_pf = env ? atoi(env) : 1;
    return QF_OK;
}

int qf_launch_poisson(qf_handle_s *h, const double2 *W, const double2 *dW, double2 *Wh, double2 *P, double eps,
                      bool gated, cudaStream_t st, int members)
{
    const int N = h->N;
    const int nmem = (members > 0 && members < h->batch) ? members : h->batch;
    const size_t n2 = h->mat_elems;
    const int g = gated ? 1 : 0;
    if (Wh != W) {
        dim3 gw((unsigned)std::min<size_t>((n2 + 255) / 256, (size_t)h->sm_count * 8), h->batch);
        k_whalf<<<gw, 256, 0, st>>>(W, dW, Wh, n2, h->ctrl, g);
        h->launches++;
    }
    if (h->p_L) {
        const int L = h->p_L, M = h->p_M, CL = h->p_CL, NT = h->p_NT;
        const int PC = (NT / M) * L;
        const size_t smem = std::max((size_t)NT * L * 8, (size_t)(PC * M + (PC / L) * 4) * sizeof(double2));
        void *fn = poisson_band_fn(h);
        if (!fn) { qf_set_error("no k_poisson_band instantiation for L=%d M=%d CL=%d", L, M, CL); return QF_ERR_UNSUPPORTED; }
        const int pf = h->p_pf;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)h->p_nunits, (unsigned)nmem);
        cfg.blockDim = dim3((unsigned)NT);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)CL;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        const double2 *a0 = Wh;
        double2 *a1 = P;
        const double *a2 = h->ptab_w, *a3 = h->ptab_iu;
        const int4 *a4 = reinterpret_cast<const int4 *>(h->ptab_units);
        int a5 = N, a6 = h->p_nunits;
        // L2 prefetch distance = units resident at once (one 512-thread CTA per SM), a multiple of the cluster size
        int a7 = pf ? (pf > 1 ? pf : (512 / h->p_NTMAX) * h->sm_count) / CL * CL : 0;
        double a8 = eps;
        const QfCtrl *a9 = h->ctrl;
        int a10 = g;
        void *args[] = {&a0, &a1, &a2, &a3, &a4, &a5, &a6, &a7, &a8, &a9, &a10};
        QF_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
        h->launches++;
    } else {
        qf_set_error("qf_launch_poisson: no work plan for N=%d", N);
        return QF_ERR_UNSUPPORTED;
    }
    QF_CUDA(cudaGetLastError());
    return QF_OK;
}

// Wh = W + dW (all members, ungated)
int qf_launch_whalf(qf_handle_s *h, const double2 *W, const double2 *dW, double2 *Wh, cudaStream_t st)
{
    const size_t n2 = h->mat_elems;
    dim3 gw((unsigned)std::min<size_t>((n2 + 255) / 256, (size_t)h->sm_count * 8), h->batch);
    k_whalf<<<gw, 256, 0, st>>>(W, dW, Wh, n2, h->ctrl, 0);
    h->launches++;
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
