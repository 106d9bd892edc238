// Roofline denominator for the FP64 tensor path, measured at run time.
//
// MEASURED_PEAKS.json carries no FP64 figure, so bench.py measures the DMMA issue peak of the device it runs on, in
// the same process and at the clocks of the same run: every warp issues independent chains of
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4, the only FP64 MMA sm_100a has), 16 warps per SM and 8 chains per warp — far
// past the point where the FP64 tensor pipe saturates (profiles/r01_fp64_pipes.txt: 4 warps x 2 chains suffice).
#include "qf_common.cuh"

namespace {

template <int ILP>
__global__ void __launch_bounds__(512)
k_dmma_issue(double *out, double a, double b, int iters)
{
    double c[ILP][2];
#pragma unroll
    for (int j = 0; j < ILP; ++j) c[j][0] = c[j][1] = 0.0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[j][0]), "+d"(c[j][1])
                         : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += c[j][0] + c[j][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}   // namespace

// *tflops_out = best of `reps` timed launches (CUDA events on `stream`), 2*8*8*4 = 512 flop per warp instruction.
extern "C" int qf_measure_fp64_tensor_peak(int device, int reps, double *tflops_out, void *stream)
{
    if (!tflops_out || reps < 1) { qf_set_error("qf_measure_fp64_tensor_peak: bad arguments"); return QF_ERR_INVALID; }
    QF_ON_DEVICE(device);
    cudaStream_t st = (cudaStream_t)stream;
    int sms = 0;
    QF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    constexpr int ILP = 8, WARPS = 16, ITERS = 8000;
    double *out = nullptr;
    QF_CUDA(cudaMalloc(&out, sizeof(double) * (size_t)sms * WARPS * 32));
    cudaEvent_t e0, e1;
    QF_CUDA(cudaEventCreate(&e0));
    QF_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = -1; r < reps; ++r) {      // r = -1: warm-up
        QF_CUDA(cudaEventRecord(e0, st));
        k_dmma_issue<ILP><<<sms, WARPS * 32, 0, st>>>(out, 1.0, 1.0, ITERS);
        QF_CUDA(cudaEventRecord(e1, st));
        QF_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        QF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 0 && ms < best) best = ms;
    }
    QF_CUDA(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops_out = (double)sms * WARPS * (double)ITERS * ILP * 512.0 / (best * 1e-3) / 1e12;
    return QF_OK;
}
