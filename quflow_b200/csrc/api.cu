// C-ABI entry points (see include/quflow_b200.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "qf_common.cuh"

void qf_comm_destroy(qf_handle_s *h);   // comm.cu

static thread_local char g_err[512] = "";

void qf_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *qf_last_error(void) { return g_err; }
extern "C" const char *qf_version(void) { return "quflow_b200 0.1.0 (sm_100a)"; }

extern "C" int qf_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        qf_set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return QF_ERR_CUDA;
    }
    return n;
}

static int alloc_mat(qf_handle_s *h, double2 **p)
{
    QF_CUDA(cudaMalloc(p, sizeof(double2) * h->mat_elems * h->batch));
    QF_CUDA(cudaMemset(*p, 0, sizeof(double2) * h->mat_elems * h->batch));
    return QF_OK;
}

extern "C" int qf_create(int N, int batch, int device, qf_handle_t *out)
{
    if (!out) { qf_set_error("qf_create: out is null"); return QF_ERR_INVALID; }
    *out = nullptr;
    if (N < 2 || N > 16384) { qf_set_error("qf_create: N=%d out of range [2, 16384]", N); return QF_ERR_INVALID; }
    if (batch < 1) { qf_set_error("qf_create: batch must be >= 1"); return QF_ERR_INVALID; }
    int ndev = qf_device_count();
    if (ndev <= 0) {
        if (ndev == 0) qf_set_error("qf_create: no CUDA device (quflow_b200 has no CPU fallback)");
        return QF_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { qf_set_error("qf_create: device %d out of range (%d devices)", device, ndev); return QF_ERR_INVALID; }
    QF_ON_DEVICE(device);
    cudaDeviceProp prop;
    QF_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        qf_set_error("qf_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return QF_ERR_UNSUPPORTED;
    }
    qf_handle_s *h = new qf_handle_s();
    h->N = N;
    h->batch = batch;
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->mat_elems = (size_t)N * N;
    h->nslots = (N + 31) / 32;
    {
        const char *env = getenv("QF_GRAPH");
        h->use_graph = !(env && env[0] == '0');
    }
    int rc = qf_build_tables(h);
    if (rc == QF_OK) rc = alloc_mat(h, &h->dW);
    if (rc == QF_OK) rc = alloc_mat(h, &h->Wh);
    if (rc == QF_OK) rc = alloc_mat(h, &h->P);
    if (rc == QF_OK) rc = alloc_mat(h, &h->A);
    if (rc == QF_OK) rc = alloc_mat(h, &h->S);
    if (rc == QF_OK) rc = alloc_mat(h, &h->scratch);
    if (rc == QF_OK) rc = qf_gemm_create(h);
    const size_t rp = sizeof(double) * (size_t)batch * 2 * h->nslots * N;
    auto cu = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && rc == QF_OK) {
            qf_set_error("qf_create: %s failed: %s", what, cudaGetErrorString(e));
            rc = QF_ERR_CUDA;
        }
    };
    if (rc == QF_OK) cu(cudaMalloc(&h->rowpart, rp), "cudaMalloc(rowpart)");
    if (rc == QF_OK) cu(cudaMemset(h->rowpart, 0, rp), "cudaMemset(rowpart)");
    h->nsd = (N + 15) / 16;
    h->nsm = (N + 31) / 32;
    const size_t rp2 = sizeof(double) * (size_t)batch * N * (h->nsd + h->nsm);
    if (rc == QF_OK) cu(cudaMalloc(&h->rowpart2, rp2), "cudaMalloc(rowpart2)");
    if (rc == QF_OK) cu(cudaMemset(h->rowpart2, 0, rp2), "cudaMemset(rowpart2)");
    {
        // The fused GEMM-2 tail is parity-green but measured slower than the separate k_post launch at every size
        // (DESIGN.md section 3.4), so it is opt-in: QF_FUSE_POST=1 or qf_set_fuse_post.
        const char *env = getenv("QF_FUSE_POST");
        h->fuse_post = (env && env[0] == '1') ? 1 : 0;
    }
    if (rc == QF_OK) cu(cudaMalloc(&h->ctrl, sizeof(QfCtrl) * batch), "cudaMalloc(ctrl)");
    if (rc == QF_OK) cu(cudaMemset(h->ctrl, 0, sizeof(QfCtrl) * batch), "cudaMemset(ctrl)");
    if (rc == QF_OK) cu(cudaMallocHost(&h->ctrl_host, sizeof(QfCtrl) * batch), "cudaMallocHost(ctrl_host)");
    if (rc != QF_OK) { qf_destroy(h); return rc; }      // frees whatever was allocated so far
    *out = h;
    return QF_OK;
}

extern "C" int qf_destroy(qf_handle_t h)
{
    if (!h) return QF_OK;
    QfDeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    qf_graph_destroy(h);
    qf_gemm_destroy(h);
    qf_comm_destroy(h);
    qf_p2p_destroy(h);
    void *ptrs[] = {h->ptab_w, h->ptab_iu, h->ptab_units, h->dW, h->Wh, h->P, h->A, h->S, h->scratch, h->kahan_c,
                    h->io, h->io2, h->rowpart, h->rowpart2, h->inner_part, h->ctrl, h->iters_dev};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (h->ctrl_host) cudaFreeHost(h->ctrl_host);
    delete h;
    return QF_OK;
}

extern "C" int64_t qf_launch_count(qf_handle_t h) { return h ? h->launches : 0; }

extern "C" int qf_set_fuse_post(qf_handle_t h, int enable)
{
    if (!h) { qf_set_error("qf_set_fuse_post: null handle"); return QF_ERR_INVALID; }
    if (enable && !qf_gemm_can_fuse_post(h)) { qf_set_error("the fused GEMM-2 tail needs the warp-specialised 3M TMA kernel"); return QF_ERR_UNSUPPORTED; }
    h->fuse_post = enable ? 1 : 0;
    qf_graph_destroy(h);        // the step graph bakes the choice in
    return QF_OK;
}

extern "C" int qf_solve_poisson(qf_handle_t h, const void *W_dev, void *P_dev, void *stream)
{
    if (!h || !W_dev || !P_dev) { qf_set_error("qf_solve_poisson: null argument"); return QF_ERR_INVALID; }
    if (W_dev == P_dev) { qf_set_error("qf_solve_poisson: W and P may not alias"); return QF_ERR_INVALID; }
    QF_ON_DEVICE(h->device);
    const double2 *W = (const double2 *)W_dev;
    // Wh == W: the solve reads W directly (no W + dW pass), eps = 1, not gated
    return qf_launch_poisson(h, W, nullptr, const_cast<double2 *>(W), (double2 *)P_dev, 1.0, false, (cudaStream_t)stream);
}

extern "C" int qf_laplace(qf_handle_t h, const void *P_dev, void *W_dev, void *stream)
{
    if (!h || !W_dev || !P_dev) { qf_set_error("qf_laplace: null argument"); return QF_ERR_INVALID; }
    if (W_dev == P_dev) { qf_set_error("qf_laplace: P and W may not alias"); return QF_ERR_INVALID; }
    QF_ON_DEVICE(h->device);
    return qf_launch_laplace(h, (const double2 *)P_dev, (double2 *)W_dev, (cudaStream_t)stream);
}

extern "C" int qf_norm_inf(qf_handle_t h, const void *W_dev, double *out_host, void *stream)
{
    if (!h || !W_dev || !out_host) { qf_set_error("qf_norm_inf: null argument"); return QF_ERR_INVALID; }
    QF_ON_DEVICE(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    QF_CHECK(qf_launch_norm_inf(h, (const double2 *)W_dev, st));
    QF_CUDA(cudaMemcpyAsync(h->ctrl_host, h->ctrl, sizeof(QfCtrl) * h->batch, cudaMemcpyDeviceToHost, st));
    QF_CUDA(cudaStreamSynchronize(st));
    for (int b = 0; b < h->batch; ++b) out_host[b] = h->ctrl_host[b].norm0;
    return QF_OK;
}

extern "C" int qf_zgemm(qf_handle_t h, const void *A_dev, const void *B_dev, void *C_dev, void *stream)
{
    if (!h || !A_dev || !B_dev || !C_dev) { qf_set_error("qf_zgemm: null argument"); return QF_ERR_INVALID; }
    QF_ON_DEVICE(h->device);
    return qf_launch_zgemm(h, (const double2 *)A_dev, (const double2 *)B_dev, (double2 *)C_dev, false, false, -1, 1, false,
                           (cudaStream_t)stream);
}

// ---- host-buffer variants ---------------------------------------------------------------
static int ensure_io(qf_handle_s *h, bool two)
{
    const size_t bytes = sizeof(double2) * h->mat_elems * h->batch;
    if (!h->io) QF_CUDA(cudaMalloc(&h->io, bytes));
    if (two && !h->io2) QF_CUDA(cudaMalloc(&h->io2, bytes));
    return QF_OK;
}

extern "C" int qf_isomp_host(qf_handle_t h, void *W_host, double dt, int steps, double tol, int maxit, int minit,
                             unsigned flags, qf_stats *stats, int32_t *iters_per_step)
{
    if (!h || !W_host) { qf_set_error("qf_isomp_host: null argument"); return QF_ERR_INVALID; }
    QF_ON_DEVICE(h->device);
    const size_t bytes = sizeof(double2) * h->mat_elems * h->batch;
    if (flags & QF_FLAG_HOST_ROWS_OWN) {
        // Row-distributed host state (tile-exchange path): this rank's host array is read and written on its own two
        // row blocks only.  Upload 1/G of the matrix over this GPU's PCIe link, complete the state on every rank over
        // NVLink, run, download the own rows again.
        if (!(h->nranks > 1 && h->comm_mode == 5)) { qf_set_error("QF_FLAG_HOST_ROWS_OWN needs the tile-exchange data path"); return QF_ERR_INVALID; }
        const int G = h->nranks, hb = qf_block_rows(h->N, G);
        const size_t blk = sizeof(double2) * (size_t)hb * h->N;
        const int blocks[2] = {h->rank, 2 * G - 1 - h->rank};
        // everybody has finished reading the previous call's state before anyone overwrites it
        QF_CHECK(qf_xchg_barrier(h, QF_XF_X, false, 0));
        for (int q = 0; q < 2; ++q)
            QF_CUDA(cudaMemcpyAsync((char *)h->Wst + blocks[q] * blk, (const char *)W_host + blocks[q] * blk, blk, cudaMemcpyHostToDevice, 0));
        QF_CHECK(qf_xchg_push_rows(h, 0));
        QF_CHECK(qf_xchg_barrier(h, QF_XF_X, false, 0));
        int rc = qf_isomp_impl(h, nullptr, dt, steps, tol, maxit, minit, flags, stats, iters_per_step, 0);
        if (rc != QF_OK && rc != QF_ERR_NONFINITE) return rc;
        for (int q = 0; q < 2; ++q)
            QF_CUDA(cudaMemcpyAsync((char *)W_host + blocks[q] * blk, (const char *)h->Wst + blocks[q] * blk, blk, cudaMemcpyDeviceToHost, 0));
        QF_CUDA(cudaStreamSynchronize(0));
        return rc;
    }
    QF_CHECK(ensure_io(h, false));
    QF_CUDA(cudaMemcpyAsync(h->io, W_host, bytes, cudaMemcpyHostToDevice, 0));
    int rc = qf_isomp(h, h->io, dt, steps, tol, maxit, minit, flags, stats, iters_per_step, 0);
    if (rc != QF_OK && rc != QF_ERR_NONFINITE) return rc;
    QF_CUDA(cudaMemcpy(W_host, h->io, bytes, cudaMemcpyDeviceToHost));
    return rc;
}

extern "C" int qf_solve_poisson_host(qf_handle_t h, const void *W_host, void *P_host)
{
    if (!h || !W_host || !P_host) { qf_set_error("qf_solve_poisson_host: null argument"); return QF_ERR_INVALID; }
    QF_ON_DEVICE(h->device);
    QF_CHECK(ensure_io(h, true));
    const size_t bytes = sizeof(double2) * h->mat_elems * h->batch;
    QF_CUDA(cudaMemcpyAsync(h->io, W_host, bytes, cudaMemcpyHostToDevice, 0));
    QF_CHECK(qf_solve_poisson(h, h->io, h->io2, 0));
    QF_CUDA(cudaMemcpy(P_host, h->io2, bytes, cudaMemcpyDeviceToHost));
    return QF_OK;
}

extern "C" int qf_laplace_host(qf_handle_t h, const void *P_host, void *W_host)
{
    if (!h || !W_host || !P_host) { qf_set_error("qf_laplace_host: null argument"); return QF_ERR_INVALID; }
    QF_ON_DEVICE(h->device);
    QF_CHECK(ensure_io(h, true));
    const size_t bytes = sizeof(double2) * h->mat_elems * h->batch;
    QF_CUDA(cudaMemcpyAsync(h->io, P_host, bytes, cudaMemcpyHostToDevice, 0));
    QF_CHECK(qf_laplace(h, h->io, h->io2, 0));
    QF_CUDA(cudaMemcpy(W_host, h->io2, bytes, cudaMemcpyDeviceToHost));
    return QF_OK;
}

