// NVLink peer-store microbenchmark (single process, GPU 0 -> GPU 1): what bandwidth can kernels on the SMs reach when
// they store into a peer's memory, as a function of the access pattern?  Calibrates the W~ exchange of the tile path.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o p2p_store_probe p2p_store_probe.cu && ./p2p_store_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// (a) store only: every lane writes 16 B, consecutive lanes consecutive addresses
__global__ void k_store16(double2 *dst, size_t n)
{
    const double2 v = make_double2(1.0, 2.0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}
// (b) copy local -> peer, U independent 16-B loads per lane in flight, then U stores
template <int U>
__global__ void k_copy16(double2 *__restrict__ dst, const double2 *__restrict__ src, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += stride * U) {
        double2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (i0 + u * stride < n) v[u] = src[i0 + u * stride];
#pragma unroll
        for (int u = 0; u < U; ++u) if (i0 + u * stride < n) dst[i0 + u * stride] = v[u];
    }
}
// (c) copy with 32-B accesses per lane
__global__ void k_copy32(double4 *__restrict__ dst, const double4 *__restrict__ src, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += stride * 4) {
        double4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (i0 + u * stride < n) v[u] = src[i0 + u * stride];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (i0 + u * stride < n) dst[i0 + u * stride] = v[u];
    }
}
// (d) bulk copies: one thread per CTA stages 16 KB pieces through shared memory with cp.async.bulk (global->shared,
// shared->peer global), double-buffered
__global__ void __launch_bounds__(32)
k_bulk(char *__restrict__ dst, const char *__restrict__ src, size_t bytes, int piece)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar[2];
    if (threadIdx.x != 0) return;
    const unsigned b0 = (unsigned)__cvta_generic_to_shared(&bar[0]);
    for (int s = 0; s < 2; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b0 + 8 * s));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const unsigned s0 = (unsigned)__cvta_generic_to_shared(sm);
    const size_t npieces = bytes / piece;
    unsigned it = 0;
    for (size_t p = blockIdx.x; p < npieces; p += gridDim.x, ++it) {
        const unsigned st = it & 1, ph = (it >> 1) & 1;
        if (it >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");     // the store that used this stage has read it
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b0 + 8 * st), "r"(piece) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(s0 + st * piece), "l"(src + p * piece), "r"(piece), "r"(b0 + 8 * st) : "memory");
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }"
                         : "=r"(done) : "r"(b0 + 8 * st), "r"(ph) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + p * piece), "r"(s0 + st * piece), "r"(piece) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
static float best_ms(F f)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { printf("needs 2 GPUs\n"); return 0; }
    CK(cudaSetDevice(1));
    const size_t bytes = 64ull << 20;
    char *peer; CK(cudaMalloc(&peer, bytes));
    CK(cudaSetDevice(0));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    char *local, *local2; CK(cudaMalloc(&local, bytes)); CK(cudaMalloc(&local2, bytes));
    CK(cudaMemset(local, 1, bytes));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount;
    auto gbs = [&](float ms) { return bytes / (ms * 1e-3) / 1e9; };
    printf("GPU0 -> GPU1, %zu MiB per transfer, %d SMs\n", bytes >> 20, sms);
    { float ms = best_ms([&] { CK(cudaMemcpyAsync(peer, local, bytes, cudaMemcpyDeviceToDevice)); }); printf("cudaMemcpyAsync peer               %7.1f us  %6.1f GB/s\n", ms * 1e3, gbs(ms)); }
    { float ms = best_ms([&] { CK(cudaMemcpy2DAsync(peer, 32768, local, 32768, 16384, bytes / 32768, cudaMemcpyDeviceToDevice)); }); printf("cudaMemcpy2DAsync 16 KB rows (half) %7.1f us  %6.1f GB/s\n", ms * 1e3, gbs(ms) / 2); }
    { float ms = best_ms([&] { CK(cudaMemcpy2DAsync(peer, 32768, local, 32768, 2048, bytes / 32768, cudaMemcpyDeviceToDevice)); }); printf("cudaMemcpy2DAsync 2 KB rows (1/16)  %7.1f us  %6.1f GB/s\n", ms * 1e3, gbs(ms) / 16); }
    { float ms = best_ms([&] { CK(cudaMemcpyAsync(local2, local, bytes, cudaMemcpyDeviceToDevice)); }); printf("cudaMemcpyAsync local (reference)   %7.1f us  %6.1f GB/s\n", ms * 1e3, gbs(ms)); }
    for (int mult : {1, 2, 4, 8}) {
        const int g = sms * mult;
        { float ms = best_ms([&] { k_store16<<<g, 256>>>((double2 *)peer, bytes / 16); }); printf("store16   grid %4d x 256            %7.1f us  %6.1f GB/s\n", g, ms * 1e3, gbs(ms)); }
        { float ms = best_ms([&] { k_copy16<4><<<g, 256>>>((double2 *)peer, (const double2 *)local, bytes / 16); }); printf("copy16 U4 grid %4d x 256            %7.1f us  %6.1f GB/s\n", g, ms * 1e3, gbs(ms)); }
        { float ms = best_ms([&] { k_copy16<8><<<g, 256>>>((double2 *)peer, (const double2 *)local, bytes / 16); }); printf("copy16 U8 grid %4d x 256            %7.1f us  %6.1f GB/s\n", g, ms * 1e3, gbs(ms)); }
        { float ms = best_ms([&] { k_copy32<<<g, 256>>>((double4 *)peer, (const double4 *)local, bytes / 32); }); printf("copy32 U4 grid %4d x 256            %7.1f us  %6.1f GB/s\n", g, ms * 1e3, gbs(ms)); }
    }
    for (int piece : {4096, 16384, 32768}) {
        CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 32768));
        for (int mult : {1, 2, 4}) {
            const int g = sms * mult;
            float ms = best_ms([&] { k_bulk<<<g, 32, 2 * piece>>>(peer, local, bytes, piece); });
            printf("bulk G2S+S2G piece %5d grid %4d      %7.1f us  %6.1f GB/s\n", piece, g, ms * 1e3, gbs(ms));
        }
    }
    // the same to local memory, for reference
    { float ms = best_ms([&] { k_copy16<8><<<sms * 4, 256>>>((double2 *)local2, (const double2 *)local, bytes / 16); }); printf("copy16 U8 LOCAL grid %4d            %7.1f us  %6.1f GB/s\n", sms * 4, ms * 1e3, gbs(ms)); }
    { float ms = best_ms([&] { k_copy16<8><<<sms * 4, 256>>>((double2 *)local2, (const double2 *)peer, bytes / 16); }); printf("PULL copy16 U8 grid %4d (peer->local) %5.1f us  %6.1f GB/s\n", sms * 4, ms * 1e3, gbs(ms)); }
    return 0;
}
