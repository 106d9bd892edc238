// Multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.
//
// New functionality (the reference is single-process).  The state (W, dW, W~, P~) is replicated;
// the two GEMMs are sharded by row blocks (qf_prow in qf_common.cuh) and each is completed by ONE in-place
// ncclAllGather of the rank-permuted output.  Everything downstream (k_post, k_control, k_update) runs replicated
// on identical bytes, so all ranks take the same convergence decisions without a scalar all-reduce.
#include <stdlib.h>
#include <string.h>

#include "qf_common.cuh"

#ifdef QF_WITH_NCCL
#include <nccl.h>

#define QF_NCCL(call)                                                                            \
    do {                                                                                         \
        ncclResult_t _r = (call);                                                                \
        if (_r != ncclSuccess) {                                                                 \
            qf_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, ncclGetErrorString(_r)); \
            return QF_ERR_NCCL;                                                                  \
        }                                                                                        \
    } while (0)

static_assert(sizeof(ncclUniqueId) == QF_UNIQUE_ID_BYTES, "ncclUniqueId size changed");

extern "C" int qf_comm_get_unique_id(void *id_out)
{
    if (!id_out) { qf_set_error("qf_comm_get_unique_id: null"); return QF_ERR_INVALID; }
    ncclUniqueId id;
    QF_NCCL(ncclGetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return QF_OK;
}

extern "C" int qf_comm_init(qf_handle_t h, const void *unique_id, int rank, int nranks)
{
    if (!h || !unique_id || nranks < 1 || rank < 0 || rank >= nranks) { qf_set_error("qf_comm_init: bad arguments"); return QF_ERR_INVALID; }
    if (h->batch != 1) { qf_set_error("row sharding is for a single large-N simulation (batch == 1); ensembles shard per member"); return QF_ERR_INVALID; }
    if (nranks > 1 && h->N % (2 * nranks) != 0) { qf_set_error("row sharding needs N divisible by 2*nranks (N=%d, nranks=%d)", h->N, nranks); return QF_ERR_INVALID; }
    QF_ON_DEVICE(h->device);
    if (h->nccl_comm) { ncclCommDestroy((ncclComm_t)h->nccl_comm); h->nccl_comm = nullptr; }
    h->rank = rank;
    h->nranks = nranks;
    h->comm_mode = 0;
    if (nranks == 1) return QF_OK;
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    ncclComm_t comm;
    QF_NCCL(ncclCommInitRank(&comm, nranks, id, rank));
    h->nccl_comm = comm;
    h->comm_mode = 1;
    h->use_graph = 0;   // NCCL nodes are not allowed inside a conditional (WHILE) graph body
    return QF_OK;
}

// In-place all-gather of a rank-permuted N x N matrix: rank r contributes permuted rows [2 r hb, 2 (r+1) hb).
int qf_comm_allgather_rows(qf_handle_s *h, double2 *M, cudaStream_t st)
{
    const int hb = qf_block_rows(h->N, h->nranks);
    const size_t count = (size_t)2 * hb * h->N * 2;   // doubles per rank
    double *base = reinterpret_cast<double *>(M);
    QF_NCCL(ncclAllGather(base + (size_t)h->rank * count, base, count, ncclDouble, (ncclComm_t)h->nccl_comm, st));
    return QF_OK;
}

void qf_comm_destroy(qf_handle_s *h)
{
    if (h->nccl_comm) ncclCommDestroy((ncclComm_t)h->nccl_comm);
    h->nccl_comm = nullptr;
}
#else
extern "C" int qf_comm_get_unique_id(void *) { qf_set_error("built without NCCL"); return QF_ERR_UNSUPPORTED; }
extern "C" int qf_comm_init(qf_handle_t, const void *, int, int) { qf_set_error("built without NCCL"); return QF_ERR_UNSUPPORTED; }
int qf_comm_allgather_rows(qf_handle_s *, double2 *, cudaStream_t) { qf_set_error("built without NCCL"); return QF_ERR_UNSUPPORTED; }
void qf_comm_destroy(qf_handle_s *) {}
#endif

// ---------------------------------------------------------------------------------------
// Peer-memory all-gather over NVLink (default data path for the row-sharded GEMM outputs).
//
// NCCL cannot live inside the body of a CUDA conditional (WHILE) node and cannot be gated by a device flag, so the
// per-iteration gathers are done by a plain kernel that PULLS the peers' row blocks through P2P-mapped memory
// (cudaIpc handles exchanged once by the host side).  Synchronisation is a monotone sequence number per source
// rank written into every peer's flag array: a rank signals "my rows of gather #seq are complete" at the start of
// its gather kernel (its GEMM finished earlier on the same stream) and every CTA waits until all sources reached
// seq before copying.  Write-after-read safety needs no extra handshake: a rank can only start overwriting its
// rows of A in iteration i+1 after it passed the S gather of iteration i, which required every peer's "S ready"
// signal, which each peer raises only after its own A gather of iteration i has finished (stream order); the same
// chain protects S.
// ---------------------------------------------------------------------------------------
struct QfP2P {
    int nranks = 0, rank = 0;
    unsigned long long *flags = nullptr;               // [2 * MAXR] local, written by peers ([MAXR..): push barrier)
    double2 *AS2 = nullptr;                            // [A2 | S2]: odd-iteration copies for the push mode (one allocation)
    double2 *peerA[QF_MAX_RANKS] = {}, *peerS[QF_MAX_RANKS] = {}, *peerAS2[QF_MAX_RANKS] = {};
    unsigned long long *peerFlags[QF_MAX_RANKS] = {};
    // device copies of the pointer tables
    double2 **peerA_dev = nullptr, **peerS_dev = nullptr;
    double2 **pushA_dev = nullptr, **pushS_dev = nullptr;   // [2][MAXR], parity-major
    unsigned long long **peerFlags_dev = nullptr;
};

struct QfP2PBlob {
    cudaIpcMemHandle_t A, S, flags, AS2;
};
static_assert(sizeof(QfP2PBlob) <= QF_P2P_BLOB_BYTES, "blob too large");

extern "C" int qf_comm_p2p_export(qf_handle_t h, void *blob_out)
{
    if (!h || !blob_out) { qf_set_error("qf_comm_p2p_export: null"); return QF_ERR_INVALID; }
    QF_ON_DEVICE(h->device);
    QfP2P *p = reinterpret_cast<QfP2P *>(h->p2p);
    if (!p) {
        p = new QfP2P();
        h->p2p = p;
        QF_CUDA(cudaMalloc(&p->flags, sizeof(unsigned long long) * 2 * QF_MAX_RANKS));
        QF_CUDA(cudaMemset(p->flags, 0, sizeof(unsigned long long) * 2 * QF_MAX_RANKS));
        QF_CUDA(cudaMalloc(&p->AS2, sizeof(double2) * 2 * h->mat_elems));
        QF_CUDA(cudaMemset(p->AS2, 0, sizeof(double2) * 2 * h->mat_elems));
        h->A2 = p->AS2;
        h->S2 = p->AS2 + h->mat_elems;
    }
    QfP2PBlob b;
    memset(&b, 0, sizeof(b));
    QF_CUDA(cudaIpcGetMemHandle(&b.A, h->A));
    QF_CUDA(cudaIpcGetMemHandle(&b.S, h->S));
    QF_CUDA(cudaIpcGetMemHandle(&b.flags, p->flags));
    QF_CUDA(cudaIpcGetMemHandle(&b.AS2, p->AS2));
    memset(blob_out, 0, QF_P2P_BLOB_BYTES);
    memcpy(blob_out, &b, sizeof(b));
    return QF_OK;
}

extern "C" int qf_comm_p2p_import(qf_handle_t h, const void *blobs, int rank, int nranks)
{
    if (!h || !blobs || nranks < 1 || nranks > QF_MAX_RANKS || rank < 0 || rank >= nranks) { qf_set_error("qf_comm_p2p_import: bad arguments"); return QF_ERR_INVALID; }
    if (h->batch != 1) { qf_set_error("row sharding needs batch == 1"); return QF_ERR_INVALID; }
    if (nranks > 1 && h->N % (2 * nranks) != 0) { qf_set_error("row sharding needs N divisible by 2*nranks (N=%d, nranks=%d)", h->N, nranks); return QF_ERR_INVALID; }
    QfP2P *p = reinterpret_cast<QfP2P *>(h->p2p);
    if (!p) { qf_set_error("qf_comm_p2p_import: call qf_comm_p2p_export first"); return QF_ERR_INVALID; }
    QF_ON_DEVICE(h->device);
    p->nranks = nranks;
    p->rank = rank;
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) {
            p->peerA[r] = h->A;
            p->peerS[r] = h->S;
            p->peerFlags[r] = p->flags;
            p->peerAS2[r] = p->AS2;
            continue;
        }
        QfP2PBlob b;
        memcpy(&b, (const char *)blobs + (size_t)r * QF_P2P_BLOB_BYTES, sizeof(b));
        QF_CUDA(cudaIpcOpenMemHandle((void **)&p->peerA[r], b.A, cudaIpcMemLazyEnablePeerAccess));
        QF_CUDA(cudaIpcOpenMemHandle((void **)&p->peerS[r], b.S, cudaIpcMemLazyEnablePeerAccess));
        QF_CUDA(cudaIpcOpenMemHandle((void **)&p->peerFlags[r], b.flags, cudaIpcMemLazyEnablePeerAccess));
        QF_CUDA(cudaIpcOpenMemHandle((void **)&p->peerAS2[r], b.AS2, cudaIpcMemLazyEnablePeerAccess));
    }
    QF_CUDA(cudaMalloc(&p->peerA_dev, sizeof(void *) * QF_MAX_RANKS));
    QF_CUDA(cudaMalloc(&p->peerS_dev, sizeof(void *) * QF_MAX_RANKS));
    QF_CUDA(cudaMalloc(&p->peerFlags_dev, sizeof(void *) * QF_MAX_RANKS));
    QF_CUDA(cudaMemcpy(p->peerA_dev, p->peerA, sizeof(void *) * QF_MAX_RANKS, cudaMemcpyHostToDevice));
    QF_CUDA(cudaMemcpy(p->peerS_dev, p->peerS, sizeof(void *) * QF_MAX_RANKS, cudaMemcpyHostToDevice));
    QF_CUDA(cudaMemcpy(p->peerFlags_dev, p->peerFlags, sizeof(void *) * QF_MAX_RANKS, cudaMemcpyHostToDevice));
    {
        double2 *tabA[2 * QF_MAX_RANKS] = {}, *tabS[2 * QF_MAX_RANKS] = {};
        for (int r = 0; r < nranks; ++r) {
            tabA[r] = p->peerA[r];
            tabS[r] = p->peerS[r];
            tabA[QF_MAX_RANKS + r] = p->peerAS2[r];
            tabS[QF_MAX_RANKS + r] = p->peerAS2[r] + h->mat_elems;
        }
        QF_CUDA(cudaMalloc(&p->pushA_dev, sizeof(tabA)));
        QF_CUDA(cudaMalloc(&p->pushS_dev, sizeof(tabS)));
        QF_CUDA(cudaMemcpy(p->pushA_dev, tabA, sizeof(tabA), cudaMemcpyHostToDevice));
        QF_CUDA(cudaMemcpy(p->pushS_dev, tabS, sizeof(tabS), cudaMemcpyHostToDevice));
    }
    h->rank = rank;
    h->nranks = nranks;
    // Default data path: the GEMM epilogue pushes its tiles to the peers (fused all-gather).  QF_COMM=pull keeps the
    // separate pull kernels (also used when the warp-specialised 3M TMA GEMM is switched off).
    const char *env = getenv("QF_COMM");
    bool pull = env && strcmp(env, "pull") == 0;
    if (!env || (strcmp(env, "pull") != 0 && strcmp(env, "push") != 0)) {
        // The push hides behind the GEMM only while finished tiles leave early: with fewer than two data-parallel waves
        // per rank the stream-K schedule completes every tile at the very end and the pull (overlapped with the second
        // GEMM) is faster (measured: push +3 % at 2 GPUs, -33 % at 8 GPUs, N = 2048).
        const long long tiles = ((long long)h->N / nranks / 64) * (h->N / 64);
        pull = tiles < 2LL * h->sm_count;
    }
    h->comm_mode = (nranks > 1) ? (pull ? 2 : 3) : 0;
    if (nranks > 1 && env && strcmp(env, "pushcopy") == 0) h->comm_mode = 4;
    return QF_OK;
}

void qf_p2p_destroy(qf_handle_s *h)
{
    QfP2P *p = reinterpret_cast<QfP2P *>(h->p2p);
    if (!p) return;
    for (int r = 0; r < p->nranks; ++r) {
        if (r == p->rank) continue;
        if (p->peerA[r]) cudaIpcCloseMemHandle(p->peerA[r]);
        if (p->peerS[r]) cudaIpcCloseMemHandle(p->peerS[r]);
        if (p->peerFlags[r]) cudaIpcCloseMemHandle(p->peerFlags[r]);
        if (p->peerAS2[r]) cudaIpcCloseMemHandle(p->peerAS2[r]);
    }
    if (p->pushA_dev) cudaFree(p->pushA_dev);
    if (p->pushS_dev) cudaFree(p->pushS_dev);
    if (p->AS2) cudaFree(p->AS2);
    h->A2 = h->S2 = nullptr;
    if (p->peerA_dev) cudaFree(p->peerA_dev);
    if (p->peerS_dev) cudaFree(p->peerS_dev);
    if (p->peerFlags_dev) cudaFree(p->peerFlags_dev);
    if (p->flags) cudaFree(p->flags);
    delete p;
    h->p2p = nullptr;
}

// kind 0: gather of A (full rows), kind 1: gather of S (only the columns at or right of the diagonal block are
// needed by k_post).  seq = 2 * gseq + kind + 1 is monotone over the life of the handle.
// One warp per pulled row, 8 independent 16-byte loads in flight per lane.
__global__ void __launch_bounds__(256)
k_p2p_allgather(double2 *const *__restrict__ peers, unsigned long long *const *__restrict__ peer_flags,
                volatile unsigned long long *my_flags, int rank, int nranks, int N, int hb, int kind,
                const QfCtrl *__restrict__ ctrl, int gated)
{
    if (gated && !ctrl[0].active) return;
    const unsigned long long seq = 2ull * ctrl[0].gseq + (unsigned long long)kind + 1ull;
    if (blockIdx.x == 0 && threadIdx.x < nranks && threadIdx.x != rank) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(peer_flags[threadIdx.x] + rank) = seq;   // "my rows are ready"
    }
    if (threadIdx.x == 0) {
        for (int p = 0; p < nranks; ++p) {
            if (p == rank) continue;
            while (my_flags[p] < seq) __nanosleep(200);
        }
        __threadfence_system();
    }
    __syncthreads();
    double2 *mine = peers[rank];
    const int lane = threadIdx.x & 31;
    const int rows_per_rank = 2 * hb;
    const int total_rows = rows_per_rank * (nranks - 1);
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total_rows; w += warps) {
        const int q = w / rows_per_rank;
        const int src = q < rank ? q : q + 1;
        const int lr = w - q * rows_per_rank;                 // row inside the source's permuted region
        const int prow = src * rows_per_rank + lr;            // permuted row index
        int c0 = 0;
        if (kind == 1) {
            // logical row of this permuted row: slot 0 -> block src, slot 1 -> block 2G-1-src
            const int blk = lr < hb ? src : 2 * nranks - 1 - src;
            c0 = ((blk * hb) / 64) * 64;                      // k_post reads S only at columns >= its row (64-aligned tiles)
        }
        const double2 *__restrict__ s = peers[src] + (size_t)prow * N;
        double2 *__restrict__ d = mine + (size_t)prow * N;
        // 8 independent 16-byte NVLink loads in flight per lane before the first store
        for (int c = c0 + lane; c < N; c += 32 * 8) {
            double2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (c + 32 * u < N) v[u] = __ldcg(s + c + 32 * u);
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (c + 32 * u < N) d[c + 32 * u] = v[u];
        }
    }
}

int qf_comm_p2p_allgather(qf_handle_s *h, int kind, bool gated, cudaStream_t st)
{
    QfP2P *p = reinterpret_cast<QfP2P *>(h->p2p);
    const int hb = qf_block_rows(h->N, h->nranks);
    k_p2p_allgather<<<h->sm_count * 2, 256, 0, st>>>(kind == 0 ? p->peerA_dev : p->peerS_dev, p->peerFlags_dev, p->flags, h->rank,
                                                 h->nranks, h->N, hb, kind, h->ctrl, gated ? 1 : 0);
    h->launches++;
    QF_CUDA(cudaGetLastError());
    return QF_OK;
}

// ---------------------------------------------------------------------------------------
// Push mode (default): the GEMM kernel stores every finished tile of its row blocks into all peers' copies of the
// output as well (zgemm.cu, k_zgemm3m_ws epilogue), so the all-gather rides on the GEMM and needs no kernel of its
// own.  What remains is one barrier per fixed-point iteration, after both GEMMs: "my tiles of iteration #seq have
// landed everywhere" / "everybody's have landed here".
// Write-after-read safety: A and S are double-buffered by the parity of the iteration counter.  A rank reaches the
// GEMMs of iteration i+2 only after the barrier of iteration i+1, which needs every peer's signal i+1, which a peer
// raises after its GEMMs of i+1, i.e. after its k_post of iteration i has finished reading buffer i mod 2.
// ---------------------------------------------------------------------------------------
__global__ void k_push_barrier(unsigned long long *const *__restrict__ peer_flags, volatile unsigned long long *my_flags,
                               int rank, int nranks, const QfCtrl *__restrict__ ctrl, int gated)
{
    if (gated && !ctrl[0].active) return;
    const unsigned long long seq = ctrl[0].gseq + 1ull;
    const int t = threadIdx.x;
    if (t < nranks && t != rank) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(peer_flags[t] + QF_MAX_RANKS + rank) = seq;
        while (my_flags[QF_MAX_RANKS + t] < seq) __nanosleep(100);
        __threadfence_system();
    }
}

int qf_comm_push_barrier(qf_handle_s *h, bool gated, cudaStream_t st)
{
    QfP2P *p = reinterpret_cast<QfP2P *>(h->p2p);
    k_push_barrier<<<1, 32, 0, st>>>(p->peerFlags_dev, p->flags, h->rank, h->nranks, h->ctrl, gated ? 1 : 0);
    h->launches++;
    QF_CUDA(cudaGetLastError());
    return QF_OK;
}

// Push-copy mode: the GEMMs write their (double-buffered) outputs locally; afterwards one kernel copies this rank's
// rows of A and of S into every peer's copy with plain remote stores (NVLink writes run faster than the reads of the
// pull kernels), followed by the same flag barrier.  One warp per (matrix, row): the row is read once from local
// memory, 8 x 16 bytes per lane in flight, and stored to all peers.
__global__ void __launch_bounds__(256)
k_push_rows(double2 *const *__restrict__ peersA, double2 *const *__restrict__ peersS, int rank, int nranks, int N, int hb,
            const QfCtrl *__restrict__ ctrl, int gated)
{
    if (gated && !ctrl[0].active) return;
    const int par = (int)(ctrl[0].gseq & 1ull);
    double2 *const *pa = peersA + par * QF_MAX_RANKS;
    double2 *const *ps = peersS + par * QF_MAX_RANKS;
    const int lane = threadIdx.x & 31;
    const int rows_per_rank = 2 * hb;
    const int total = 2 * rows_per_rank;                   // A rows, then S rows
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total; w += warps) {
        const int kind = w / rows_per_rank;
        const int lr = w - kind * rows_per_rank;           // row inside this rank's permuted region
        const int prow = rank * rows_per_rank + lr;
        int c0 = 0;
        if (kind == 1) {                                   // k_post reads S only at columns >= its row (64-aligned tiles)
            const int blk = lr < hb ? rank : 2 * nranks - 1 - rank;
            c0 = ((blk * hb) / 64) * 64;
        }
        double2 *const *tab = kind ? ps : pa;
        const double2 *__restrict__ src = tab[rank] + (size_t)prow * N;
        for (int c = c0 + lane; c < N; c += 32 * 8) {
            double2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (c + 32 * u < N) v[u] = __ldcg(src + c + 32 * u);
            for (int p = 0; p < nranks; ++p) {
                if (p == rank) continue;
                double2 *__restrict__ d = tab[p] + (size_t)prow * N;
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (c + 32 * u < N) d[c + 32 * u] = v[u];
            }
        }
    }
    __threadfence_system();
}

int qf_comm_push_rows(qf_handle_s *h, bool gated, cudaStream_t st)
{
    QfP2P *p = reinterpret_cast<QfP2P *>(h->p2p);
    const int hb = qf_block_rows(h->N, h->nranks);
    k_push_rows<<<h->sm_count * 2, 256, 0, st>>>(p->pushA_dev, p->pushS_dev, h->rank, h->nranks, h->N, hb, h->ctrl, gated ? 1 : 0);
    h->launches++;
    QF_CUDA(cudaGetLastError());
    return QF_OK;
}

int qf_comm_push_args(qf_handle_s *h, int kind, QfGemmPush *out)
{
    QfP2P *p = reinterpret_cast<QfP2P *>(h->p2p);
    if (!p || !p->pushA_dev) { qf_set_error("push mode needs qf_comm_p2p_import"); return QF_ERR_INVALID; }
    out->C1 = kind == 0 ? h->A2 : h->S2;
    out->A1 = kind == 0 ? nullptr : h->A2;      // the second GEMM multiplies the A of the same iteration
    out->peers = (h->comm_mode == 4) ? nullptr : (kind == 0 ? p->pushA_dev : p->pushS_dev);   // mode 4 copies after the GEMMs
    out->nranks = h->nranks;
    out->rank = h->rank;
    return QF_OK;
}

// Switch between the fused push (1, default) and the separate pull kernels (0) after qf_comm_p2p_import.
extern "C" int qf_comm_set_push(qf_handle_t h, int enable)
{
    if (!h || !h->p2p || h->nranks < 2 || h->comm_mode < 2) {
        qf_set_error("qf_comm_set_push: the handle has no peer-memory communicator");
        return QF_ERR_INVALID;
    }
    h->comm_mode = enable == 2 ? 4 : (enable ? 3 : 2);   // 2: push-copy after the GEMMs
    qf_graph_destroy(h);   // the step graph bakes the data path in
    return QF_OK;
}

// 0: single GPU / emulated ranks, 1: NCCL all-gather, 2: pull kernels, 3: fused GEMM + push
extern "C" int qf_comm_mode(qf_handle_t h) { return h ? h->comm_mode : 0; }

// Single-GPU emulation of G ranks (tests): same tile lists, same permuted layout, no communication.
extern "C" int qf_set_emulated_ranks(qf_handle_t h, int nranks)
{
    if (!h || nranks < 1) { qf_set_error("qf_set_emulated_ranks: bad arguments"); return QF_ERR_INVALID; }
    if (h->comm_mode != 0) { qf_set_error("qf_set_emulated_ranks: handle already has a communicator"); return QF_ERR_INVALID; }
    if (h->batch != 1) { qf_set_error("row sharding needs batch == 1"); return QF_ERR_INVALID; }
    if (nranks > 1 && h->N % (2 * nranks) != 0) { qf_set_error("row sharding needs N divisible by 2*nranks"); return QF_ERR_INVALID; }
    h->rank = 0;
    h->nranks = nranks;
    return QF_OK;
}
