// Library comparator for the Poisson solve: cusparseDgtsv2StridedBatch with the packing the reference's
// own GPU prototype uses (quflow/experimental/cuda.py:10-44, 123-166: diagonals m and N-m share one system of length N,
// batch = N/2 + 1 systems, real and imaginary parts of the right-hand side as two calls).  Only the two library solves are
// timed — the pack / unpack kernels the prototype needs around them (extract_body / insert_body) are NOT included, so
// this is a lower bound for the incumbent.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o cusparse_gtsv_ref cusparse_gtsv_ref.cu -lcusparse
#include <cuda_runtime.h>
#include <cusparse.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { auto _e = (x); if (_e != 0) { printf("error %d at line %d\n", (int)_e, __LINE__); return 1; } } while (0)

int main(int argc, char **argv)
{
    cusparseHandle_t hs;
    CK(cusparseCreate(&hs));
    std::vector<int> sizes = {512, 1024, 2048};
    if (argc > 1) sizes = {atoi(argv[1])};      // bench.py passes the size it is timing
    for (int N : sizes) {
        const int batch = N / 2 + 1, m = N;
        const size_t tot = (size_t)batch * m;
        std::vector<double> dl(tot), d(tot), du(tot), x(tot);
        for (int b = 0; b < batch; ++b)          // system b: diagonal b (length N-b) followed by diagonal N-b (length b)
            for (int p = 0; p < m; ++p) {
                const bool first = p < N - b;
                const double mm = first ? b : N - b, k = first ? p : p - (N - b);
                const double dN = N;
                const size_t i = (size_t)b * m + p;
                d[i] = -((dN - 1.0) * (2.0 * k + 1.0 + mm) - 2.0 * k * (k + mm)) - (b == 0 && p == 0 ? 0.5 : 0.0);
                const double o = sqrt(((k + mm) * (dN - k - mm)) * (k * (dN - k)));
                dl[i] = (k > 0) ? o : 0.0;
                const double kn = k + 1.0;
                const bool last = first ? (p == N - b - 1) : (p == m - 1);
                du[i] = last ? 0.0 : sqrt(((kn + mm) * (dN - kn - mm)) * (kn * (dN - kn)));
                x[i] = sin(0.37 * (double)i);
            }
        double *ddl, *dd, *ddu, *dx, *dx2;
        CK(cudaMalloc(&ddl, tot * 8)); CK(cudaMalloc(&dd, tot * 8)); CK(cudaMalloc(&ddu, tot * 8));
        CK(cudaMalloc(&dx, tot * 8)); CK(cudaMalloc(&dx2, tot * 8));
        CK(cudaMemcpy(ddl, dl.data(), tot * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dd, d.data(), tot * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ddu, du.data(), tot * 8, cudaMemcpyHostToDevice));
        size_t bufsz = 0;
        CK(cusparseDgtsv2StridedBatch_bufferSizeExt(hs, m, ddl, dd, ddu, dx, batch, m, &bufsz));
        void *buf;
        CK(cudaMalloc(&buf, bufsz));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e9f;
        for (int rep = 0; rep < 12; ++rep) {
            CK(cudaMemcpy(dx, x.data(), tot * 8, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(dx2, x.data(), tot * 8, cudaMemcpyHostToDevice));
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            CK(cusparseDgtsv2StridedBatch(hs, m, ddl, dd, ddu, dx, batch, m, buf));     // real parts
            CK(cusparseDgtsv2StridedBatch(hs, m, ddl, dd, ddu, dx2, batch, m, buf));    // imaginary parts
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep >= 2 && ms < best) best = ms;
        }
        printf("N=%d  cusparseDgtsv2StridedBatch x2 (batch=%d, m=%d): %.1f us  (%.0f GB/s of the 32 N^2 algorithmic bytes; pack/unpack not included)\n",
               N, batch, m, best * 1e3, 32.0 * N * N / (best * 1e-3) / 1e9);
        cudaFree(ddl); cudaFree(dd); cudaFree(ddu); cudaFree(dx); cudaFree(dx2); cudaFree(buf);
    }
    cusparseDestroy(hs);
    return 0;
}
