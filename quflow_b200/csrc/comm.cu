// Multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.
//
// New functionality (the reference is single-process).  The state (W, dW, W~, P~) is replicated;
// the two GEMMs are sharded by row blocks (qf_prow in qf_common.cuh) and each is completed by ONE in-place
// ncclAllGather of the rank-permuted output.  Everything downstream (k_post, k_control, k_update) runs replicated
// on identical bytes, so all ranks take the same convergence decisions without a scalar all-reduce.
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
    QF_CUDA(cudaSetDevice(h->device));
    if (h->nccl_comm) { ncclCommDestroy((ncclComm_t)h->nccl_comm); h->nccl_comm = nullptr; }
    h->rank = rank;
    h->nranks = nranks;
    if (nranks == 1) return QF_OK;
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    ncclComm_t comm;
    QF_NCCL(ncclCommInitRank(&comm, nranks, id, rank));
    h->nccl_comm = comm;
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

// Single-GPU emulation of G ranks (tests): same tile lists, same permuted layout, no communication.
extern "C" int qf_set_emulated_ranks(qf_handle_t h, int nranks)
{
    if (!h || nranks < 1) { qf_set_error("qf_set_emulated_ranks: bad arguments"); return QF_ERR_INVALID; }
    if (h->nccl_comm) { qf_set_error("qf_set_emulated_ranks: handle already has a communicator"); return QF_ERR_INVALID; }
    if (h->batch != 1) { qf_set_error("row sharding needs batch == 1"); return QF_ERR_INVALID; }
    if (nranks > 1 && h->N % (2 * nranks) != 0) { qf_set_error("row sharding needs N divisible by 2*nranks"); return QF_ERR_INVALID; }
    h->rank = 0;
    h->nranks = nranks;
    return QF_OK;
}
