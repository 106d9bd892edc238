// Multi-GPU plumbing: one process per GPU; new functionality (the reference is single-process).
//
// One large-N simulation is sharded by row blocks of the two GEMMs (rank r owns blocks r and 2G-1-r of the 2G blocks).
// Three data paths complete the iteration, all selected per handle (comm_mode):
//   5  tile exchange (default, bottom of this file): the tail of the iteration and the update are sharded by tile pairs
//      as well; per iteration a rank pushes the few lower tiles of A its peers need and its new W~ tiles over NVLink peer
//      mappings, inside the step graph; the Poisson solve is the only replicated kernel;
//   2  pull all-gather: A and S (rank-permuted rows) are completed on every rank by kernels that pull the peers' rows;
//      tail and update run replicated;
//   1  NCCL: the same gathers as one in-place ncclAllGather each, eager launches (NCCL cannot run inside the body of a
//      conditional graph node).
// On every path all ranks evaluate the stopping rule on identical bytes, so they take the same decisions without a
// scalar all-reduce.
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
// Peer memory.  Everything a peer ever touches lives in ONE allocation per rank, the arena
//     [ W~ | A | Wst | rowpart | rowpart2 | flags ]
// exported as one CUDA IPC handle (plus the separately allocated S for the pull path).  The host side all-gathers the
// blobs (torch.distributed) and every rank maps every peer once.
// ---------------------------------------------------------------------------------------
struct QfArenaLayout {
    size_t wh, a, wst, part, part2, flags, total;     // byte offsets
};
static QfArenaLayout arena_layout(const qf_handle_s *h)
{
    auto up = [](size_t x) { return (x + 1023) & ~(size_t)1023; };
    QfArenaLayout l;
    const size_t mat = up(sizeof(double2) * h->mat_elems);
    l.wh = 0;
    l.a = l.wh + mat;
    l.wst = l.a + mat;
    l.part = l.wst + mat;
    l.part2 = l.part + up(sizeof(double) * 2 * (size_t)h->nslots * h->N);
    l.flags = l.part2 + up(sizeof(double) * (size_t)h->N * (h->nsd + h->nsm));
    l.total = l.flags + up(sizeof(unsigned long long) * 4 * QF_MAX_RANKS);
    return l;
}
// flags: [0, 2 MAXR) pull all-gather (kind 0 / 1 signalled through one monotone sequence), [2 MAXR, 4 MAXR) tile exchange
constexpr int QF_FLAGS_XCHG = 2 * QF_MAX_RANKS;

struct QfP2P {
    int nranks = 0, rank = 0;
    bool local = false;                                // peers are handles of this process on the same device (tests)
    char *peerArena[QF_MAX_RANKS] = {};
    double2 *peerS[QF_MAX_RANKS] = {};
    // device copies of the pointer tables, QF_MAX_RANKS entries each
    double2 **peerA_dev = nullptr, **peerS_dev = nullptr, **peerWh_dev = nullptr, **peerWst_dev = nullptr;
    double **peerPart_dev = nullptr;
    unsigned long long **peerFlags_dev = nullptr;      // pull path: base of the flag block
    unsigned long long **peerXFlags_dev = nullptr;     // tile exchange: base + QF_FLAGS_XCHG
    QfXchg desc;
};

struct QfP2PBlob {
    cudaIpcMemHandle_t arena, S;
};
static_assert(sizeof(QfP2PBlob) <= QF_P2P_BLOB_BYTES, "blob too large");

static unsigned long long *arena_flags(qf_handle_s *h) { return reinterpret_cast<unsigned long long *>((char *)h->arena + arena_layout(h).flags); }

// Move the peer-visible buffers of the handle into one arena (idempotent).
static int ensure_arena(qf_handle_s *h)
{
    if (h->arena) return QF_OK;
    if (h->batch != 1) { qf_set_error("row sharding needs batch == 1"); return QF_ERR_INVALID; }
    const QfArenaLayout l = arena_layout(h);
    char *base = nullptr;
    QF_CUDA(cudaMalloc(&base, l.total));
    QF_CUDA(cudaMemset(base, 0, l.total));
    QF_CUDA(cudaDeviceSynchronize());
    qf_graph_destroy(h);                      // the step graph bakes the buffer addresses in
    cudaFree(h->Wh);
    cudaFree(h->A);
    cudaFree(h->rowpart);
    cudaFree(h->rowpart2);
    h->arena = base;
    h->Wh = reinterpret_cast<double2 *>(base + l.wh);
    h->A = reinterpret_cast<double2 *>(base + l.a);
    h->Wst = reinterpret_cast<double2 *>(base + l.wst);
    h->rowpart = reinterpret_cast<double *>(base + l.part);
    h->rowpart2 = reinterpret_cast<double *>(base + l.part2);
    return QF_OK;
}

extern "C" int qf_comm_p2p_export(qf_handle_t h, void *blob_out)
{
    if (!h || !blob_out) { qf_set_error("qf_comm_p2p_export: null"); return QF_ERR_INVALID; }
    QF_ON_DEVICE(h->device);
    QF_CHECK(ensure_arena(h));
    QfP2PBlob b;
    memset(&b, 0, sizeof(b));
    QF_CUDA(cudaIpcGetMemHandle(&b.arena, h->arena));
    QF_CUDA(cudaIpcGetMemHandle(&b.S, h->S));
    memset(blob_out, 0, QF_P2P_BLOB_BYTES);
    memcpy(blob_out, &b, sizeof(b));
    return QF_OK;
}

static void p2p_release(qf_handle_s *h)
{
    QfP2P *p = reinterpret_cast<QfP2P *>(h->p2p);
    if (!p) return;
    if (!p->local)
        for (int r = 0; r < p->nranks; ++r) {
            if (r == p->rank) continue;
            if (p->peerArena[r]) cudaIpcCloseMemHandle(p->peerArena[r]);
            if (p->peerS[r]) cudaIpcCloseMemHandle(p->peerS[r]);
        }
    void *tabs[] = {p->peerA_dev, p->peerS_dev, p->peerWh_dev, p->peerWst_dev, p->peerPart_dev, p->peerFlags_dev, p->peerXFlags_dev};
    for (void *t : tabs)
        if (t) cudaFree(t);
    delete p;
    h->p2p = nullptr;
}

// Build the device tables and the QfXchg descriptor from the mapped peers; picks the data path.
static int p2p_finish(qf_handle_s *h, QfP2P *p)
{
    const int nranks = p->nranks, rank = p->rank;
    const QfArenaLayout l = arena_layout(h);
    void *tA[QF_MAX_RANKS] = {}, *tS[QF_MAX_RANKS] = {}, *tWh[QF_MAX_RANKS] = {}, *tWst[QF_MAX_RANKS] = {}, *tPart[QF_MAX_RANKS] = {},
         *tF[QF_MAX_RANKS] = {}, *tXF[QF_MAX_RANKS] = {};
    for (int r = 0; r < nranks; ++r) {
        char *base = p->peerArena[r];
        tA[r] = base + l.a;
        tS[r] = p->peerS[r];
        tWh[r] = base + l.wh;
        tWst[r] = base + l.wst;
        tPart[r] = base + l.part;
        tF[r] = base + l.flags;
        tXF[r] = base + l.flags + sizeof(unsigned long long) * QF_FLAGS_XCHG;
    }
    auto upload = [&](void *const *tab, void **dev) -> int {
        QF_CUDA(cudaMalloc(dev, sizeof(void *) * QF_MAX_RANKS));
        QF_CUDA(cudaMemcpy(*dev, tab, sizeof(void *) * QF_MAX_RANKS, cudaMemcpyHostToDevice));
        return QF_OK;
    };
    QF_CHECK(upload(tA, (void **)&p->peerA_dev));
    QF_CHECK(upload(tS, (void **)&p->peerS_dev));
    QF_CHECK(upload(tWh, (void **)&p->peerWh_dev));
    QF_CHECK(upload(tWst, (void **)&p->peerWst_dev));
    QF_CHECK(upload(tPart, (void **)&p->peerPart_dev));
    QF_CHECK(upload(tF, (void **)&p->peerFlags_dev));
    QF_CHECK(upload(tXF, (void **)&p->peerXFlags_dev));
    p->desc.nranks = nranks;
    p->desc.rank = rank;
    p->desc.hb = qf_block_rows(h->N, nranks);
    p->desc.peerWh = p->peerWh_dev;
    p->desc.peerA = p->peerA_dev;
    p->desc.peerWst = p->peerWst_dev;
    p->desc.peerPart = p->peerPart_dev;
    p->desc.part2_off = (long long)((l.part2 - l.part) / sizeof(double));
    p->desc.peerFlags = p->peerXFlags_dev;
    p->desc.myFlags = arena_flags(h) + QF_FLAGS_XCHG;
    {
        // Upper-only W~ exchange + local mirror: half of the 16 N^2 (G-1)/G bytes a rank sends per exchange against one
        // replicated mirror pass (14 us at N = 2048); measured faster at 2, 4 and 8 GPUs.  QF_XCHG_UPPER=0 switches it off.
        const char *u = getenv("QF_XCHG_UPPER");
        p->desc.upper_only = u ? (u[0] == '1') : 1;
        // How the W~ tiles travel: stored to the peers by the tail / update kernels themselves ("inline", default: the NVLink
        // stores overlap those kernels' memory latency; 361.8 vs 352.8 steps/s at 2 GPUs), by a copy kernel on all SMs
        // ("sm"), or by copy engines ("ce", memcpy nodes in the step graph; slowest: the rows are too short for DMA).
        const char *m = getenv("QF_XCHG_PUSH");
        h->xchg_ce = (m && strcmp(m, "ce") == 0) ? 1 : 0;
        p->desc.push_inline = (m && (strcmp(m, "sm") == 0 || strcmp(m, "ce") == 0)) ? 0 : 1;
        h->skew_host = -1;
    }
    {
        // ranks may legitimately be seconds apart (host work between calls): the bound only has to end a real hang
        const char *t = getenv("QF_COMM_TIMEOUT_S");
        const double secs = t ? atof(t) : 30.0;
        p->desc.timeout_cycles = (long long)((secs > 0.0 ? secs : 30.0) * 2.0e9);
    }
    h->rank = rank;
    h->nranks = nranks;
    // Data path: tile exchange whenever the ownership blocks are whole 64-row tiles (N divisible by 128 * nranks) and the
    // warp-specialised 3M TMA GEMM is in use; otherwise the pull all-gather.  QF_COMM=pull|tile overrides.
    const char *env = getenv("QF_COMM");
    const bool can_tile = (h->N % (128 * nranks) == 0) && qf_gemm_can_fuse_post(h);
    bool tile = can_tile;
    if (env && strcmp(env, "pull") == 0) tile = false;
    if (env && strcmp(env, "tile") == 0 && !can_tile) {
        qf_set_error("QF_COMM=tile needs N divisible by 128*nranks (N=%d, nranks=%d) and the default GEMM kernel", h->N, nranks);
        return QF_ERR_UNSUPPORTED;
    }
    h->comm_mode = (nranks > 1) ? (tile ? 5 : 2) : 0;
    qf_graph_destroy(h);
    return QF_OK;
}

extern "C" int qf_comm_p2p_import(qf_handle_t h, const void *blobs, int rank, int nranks)
{
    if (!h || !blobs || nranks < 1 || nranks > QF_MAX_RANKS || rank < 0 || rank >= nranks) { qf_set_error("qf_comm_p2p_import: bad arguments"); return QF_ERR_INVALID; }
    if (h->batch != 1) { qf_set_error("row sharding needs batch == 1"); return QF_ERR_INVALID; }
    if (nranks > 1 && h->N % (2 * nranks) != 0) { qf_set_error("row sharding needs N divisible by 2*nranks (N=%d, nranks=%d)", h->N, nranks); return QF_ERR_INVALID; }
    if (!h->arena) { qf_set_error("qf_comm_p2p_import: call qf_comm_p2p_export first"); return QF_ERR_INVALID; }
    QF_ON_DEVICE(h->device);
    p2p_release(h);                           // a second import replaces the first: no leaked mappings or tables
    QfP2P *p = new QfP2P();
    h->p2p = p;
    p->nranks = nranks;
    p->rank = rank;
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) {
            p->peerArena[r] = (char *)h->arena;
            p->peerS[r] = h->S;
            continue;
        }
        QfP2PBlob b;
        memcpy(&b, (const char *)blobs + (size_t)r * QF_P2P_BLOB_BYTES, sizeof(b));
        QF_CUDA(cudaIpcOpenMemHandle((void **)&p->peerArena[r], b.arena, cudaIpcMemLazyEnablePeerAccess));
        QF_CUDA(cudaIpcOpenMemHandle((void **)&p->peerS[r], b.S, cudaIpcMemLazyEnablePeerAccess));
    }
    return p2p_finish(h, p);
}

// Test hook: attach G handles that live in THIS process on ONE device to each other (plain device pointers instead of
// IPC mappings).  Used with qf_isomp_lockstep to run the tile-exchange path for G ranks on a single GPU.
extern "C" int qf_comm_attach_local(qf_handle_t *hs, int G)
{
    if (!hs || G < 1 || G > QF_MAX_RANKS) { qf_set_error("qf_comm_attach_local: bad arguments"); return QF_ERR_INVALID; }
    for (int r = 0; r < G; ++r)
        if (!hs[r] || hs[r]->N != hs[0]->N || hs[r]->device != hs[0]->device || hs[r]->batch != 1) {
            qf_set_error("qf_comm_attach_local: the handles must share N and the device and have batch == 1");
            return QF_ERR_INVALID;
        }
    if (G > 1 && hs[0]->N % (2 * G) != 0) { qf_set_error("row sharding needs N divisible by 2*nranks (N=%d, nranks=%d)", hs[0]->N, G); return QF_ERR_INVALID; }
    QF_ON_DEVICE(hs[0]->device);
    for (int r = 0; r < G; ++r) QF_CHECK(ensure_arena(hs[r]));
    for (int r = 0; r < G; ++r) {
        p2p_release(hs[r]);
        QfP2P *p = new QfP2P();
        hs[r]->p2p = p;
        p->local = true;
        p->nranks = G;
        p->rank = r;
        for (int q = 0; q < G; ++q) {
            p->peerArena[q] = (char *)hs[q]->arena;
            p->peerS[q] = hs[q]->S;
        }
        QF_CHECK(p2p_finish(hs[r], p));
    }
    return QF_OK;
}

void qf_p2p_destroy(qf_handle_s *h)
{
    p2p_release(h);
    if (h->arena) {
        cudaFree(h->arena);
        h->arena = nullptr;
        h->Wh = h->A = h->Wst = nullptr;       // they lived inside the arena
        h->rowpart = h->rowpart2 = nullptr;
    }
}

// ---------------------------------------------------------------------------------------
// Pull all-gather (comm_mode 2).  NCCL cannot live inside the body of a CUDA conditional (WHILE) node and cannot be
// gated by a device flag, so the per-iteration gathers are done by a plain kernel that PULLS the peers' row blocks
// through the peer mappings.  Synchronisation is a monotone sequence number per source rank written into every peer's
// flag array: a rank signals "my rows of gather #seq are complete" at the start of its gather kernel (its GEMM finished
// earlier on the same stream) and every CTA waits until all sources reached seq before copying.  Write-after-read
// safety needs no extra handshake: a rank can only start overwriting its rows of A in iteration i+1 after it passed the
// S gather of iteration i, which required every peer's "S ready" signal, which each peer raises only after its own A
// gather of iteration i has finished (stream order); the same chain protects S.
// kind 0: gather of A (full rows), kind 1: gather of S (only the columns at or right of the diagonal block are
// needed by k_post).  seq = 2 * gseq + kind + 1 is monotone over the life of the handle.
// One warp per pulled row, 8 independent 16-byte loads in flight per lane.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_p2p_allgather(double2 *const *__restrict__ peers, unsigned long long *const *__restrict__ peer_flags,
                volatile unsigned long long *my_flags, int rank, int nranks, int N, int hb, int kind, QfCtrl *ctrl, int gated,
                long long timeout_cycles)
{
    if (gated && !ctrl[0].active) return;
    const unsigned long long seq = 2ull * ctrl[0].gseq + (unsigned long long)kind + 1ull;
    if (blockIdx.x == 0 && threadIdx.x < nranks && threadIdx.x != rank) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(peer_flags[threadIdx.x] + rank) = seq;   // "my rows are ready"
    }
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        for (int p = 0; p < nranks; ++p) {
            if (p == rank) continue;
            while (my_flags[p] < seq) {
                __nanosleep(200);
                if (clock64() - t0 > timeout_cycles) { ctrl[0].nonfinite = 2; break; }     // a peer never answered: give up, report
            }
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
    k_p2p_allgather<<<h->sm_count * 2, 256, 0, st>>>(kind == 0 ? p->peerA_dev : p->peerS_dev, p->peerFlags_dev, arena_flags(h), h->rank,
                                                 h->nranks, h->N, hb, kind, h->ctrl, gated ? 1 : 0, p->desc.timeout_cycles);
    h->launches++;
    QF_CUDA(cudaGetLastError());
    return QF_OK;
}

// ---------------------------------------------------------------------------------------
// Tile exchange (comm_mode 5): the kernels that store into peer memory are the GEMM (lower tiles of A, zgemm.cu), the
// tail of the iteration (k_post or the fused GEMM-2 tail: W~ tiles and residual partials) and k_update (W~ of the next
// step), all in isomp.cu / zgemm.cu.  What lives here is the synchronisation — a signal and a wait kernel per flag kind,
// separate launches so that the lock-step emulation on one GPU can enqueue every rank's signal before any rank's wait —
// and the two state exchanges at the ends of a call.
//   QF_XF_G1  seq = gseq + 1        raised after a rank's first GEMM; consumed inside the peers' tail kernels
//   QF_XF_X   seq = xseq + 1        raised after a rank's pushes of an exchange; xseq advances in the wait kernel
// Write-after-read safety: a rank pushes into a peer's W~ / partials / A only after that peer signalled G1 of the same
// iteration (tail kernels wait for it) or passed the previous exchange barrier (first GEMM), i.e. after the peer's last
// reads of the previous contents; DESIGN.md section 4 walks through every buffer.
// ---------------------------------------------------------------------------------------
const QfXchg *qf_xchg_desc(qf_handle_s *h)
{
    QfP2P *p = reinterpret_cast<QfP2P *>(h->p2p);
    return (p && h->comm_mode == 5) ? &p->desc : nullptr;
}

__global__ void k_xchg_signal(const QfXchg x, int kind, const QfCtrl *__restrict__ ctrl, int gated)
{
    if (gated && !ctrl[0].active) return;
    const unsigned long long seq = (kind == QF_XF_G1 ? ctrl[0].gseq : ctrl[0].xseq) + 1ull;
    const int t = threadIdx.x;
    if (t < x.nranks && t != x.rank) {
        __threadfence_system();          // release: this rank's earlier kernels (and their remote stores) come first
        *reinterpret_cast<volatile unsigned long long *>(x.peerFlags[t] + kind * QF_MAX_RANKS + x.rank) = seq;
    }
}

// signal + wait in one launch (a real multi-GPU run; the lock-step emulation needs them apart)
__global__ void k_xchg_barrier(const QfXchg x, int kind, QfCtrl *ctrl, int gated)
{
    if (gated && !ctrl[0].active) return;
    const unsigned long long seq = (kind == QF_XF_G1 ? ctrl[0].gseq : ctrl[0].xseq) + 1ull;
    const int t = threadIdx.x;
    if (t < x.nranks && t != x.rank) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(x.peerFlags[t] + kind * QF_MAX_RANKS + x.rank) = seq;
    }
    __syncwarp();
    if (t == 0) {
        if (ctrl[0].nonfinite != 2 && !xchg_wait_flags(x, kind, seq)) ctrl[0].nonfinite = 2;
        if (kind == QF_XF_X) ctrl[0].xseq = seq;
    }
}

__global__ void k_xchg_wait(const QfXchg x, int kind, QfCtrl *ctrl, int gated)
{
    if (gated && !ctrl[0].active) return;
    const unsigned long long seq = (kind == QF_XF_G1 ? ctrl[0].gseq : ctrl[0].xseq) + 1ull;
    if (ctrl[0].nonfinite != 2 && !xchg_wait_flags(x, kind, seq)) ctrl[0].nonfinite = 2;
    if (kind == QF_XF_X) ctrl[0].xseq = seq;
}

int qf_xchg_signal(qf_handle_s *h, int kind, bool gated, cudaStream_t st)
{
    const QfXchg *x = qf_xchg_desc(h);
    if (!x) { qf_set_error("qf_xchg_signal: the handle has no tile-exchange communicator"); return QF_ERR_INVALID; }
    k_xchg_signal<<<1, 32, 0, st>>>(*x, kind, h->ctrl, gated ? 1 : 0);
    h->launches++;
    QF_CUDA(cudaGetLastError());
    return QF_OK;
}

int qf_xchg_barrier(qf_handle_s *h, int kind, bool gated, cudaStream_t st)
{
    const QfXchg *x = qf_xchg_desc(h);
    if (!x) { qf_set_error("qf_xchg_barrier: the handle has no tile-exchange communicator"); return QF_ERR_INVALID; }
    k_xchg_barrier<<<1, 32, 0, st>>>(*x, kind, h->ctrl, gated ? 1 : 0);
    h->launches++;
    QF_CUDA(cudaGetLastError());
    return QF_OK;
}

int qf_xchg_wait(qf_handle_s *h, int kind, bool gated, cudaStream_t st)
{
    const QfXchg *x = qf_xchg_desc(h);
    if (!x) { qf_set_error("qf_xchg_wait: the handle has no tile-exchange communicator"); return QF_ERR_INVALID; }
    k_xchg_wait<<<1, 1, 0, st>>>(*x, kind, h->ctrl, gated ? 1 : 0);
    h->launches++;
    QF_CUDA(cudaGetLastError());
    return QF_OK;
}

// W~ exchange.  After the tail of an iteration (or the update) every rank holds the new W~ on the tile pairs it owns:
// the upper parts of its two row blocks, rows [b hb, (b+1) hb) x columns [b hb, N), and their mirrors, rows
// [(b+1) hb, N) x columns [b hb, (b+1) hb).  One copy kernel on all SMs stores them into every peer's W~: a warp moves a
// piece of 256 elements of one row, loaded once (eight 16-byte loads per lane in flight) and stored to each peer.
// When W is bit-for-bit skew-Hermitian (QfCtrl.skew_exact, checked at call start) W~ is too — the tail and the update
// write exact mirrors — so only the upper parts travel and every rank rebuilds the lower triangle locally
// (k_xchg_mirror_wh): half the NVLink volume.
__global__ void __launch_bounds__(256)
k_xchg_push_wh(const QfXchg x, int N, const QfCtrl *__restrict__ ctrl, int gated)
{
    if (gated && !ctrl[0].active) return;
    const bool upper_only = x.upper_only && ctrl[0].skew_exact;
    const double2 *__restrict__ mine = x.peerWh[x.rank];
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int hb = x.hb;
    for (int q = 0; q < 2; ++q) {
        const int b = q ? 2 * x.nranks - 1 - x.rank : x.rank;
        const int r0 = b * hb;
        // upper part: hb rows, pieces of 256 columns from column r0 on
        const int upieces = (N - r0 + 255) >> 8;
        // mirrored part: N - r0 - hb rows of hb columns, pieces of 256 columns
        const int mrows = upper_only ? 0 : N - r0 - hb, mpieces = (hb + 255) >> 8;
        const int total = hb * upieces + mrows * mpieces;
        for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total; w += warps) {
            int row, c0, cend;
            if (w < hb * upieces) {
                row = r0 + w / upieces;
                c0 = r0 + ((w % upieces) << 8);
                cend = N;
            } else {
                const int v = w - hb * upieces;
                row = r0 + hb + v / mpieces;
                c0 = r0 + ((v % mpieces) << 8);
                cend = r0 + hb;
            }
            const double2 *__restrict__ s = mine + (size_t)row * N;
            double2 v8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int c = c0 + lane + 32 * u;
                if (c < cend) v8[u] = s[c];
            }
            for (int p = 0; p < x.nranks; ++p) {
                if (p == x.rank) continue;
                double2 *__restrict__ d = x.peerWh[p] + (size_t)row * N;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int c = c0 + lane + 32 * u;
                    if (c < cend) d[c] = v8[u];
                }
            }
        }
    }
    // no fence: the kernel boundary orders these stores before the signal kernel, whose system-scope fence precedes the flag
}

int qf_xchg_push_wh(qf_handle_s *h, bool gated, cudaStream_t st)
{
    const QfXchg *x = qf_xchg_desc(h);
    if (!x) { qf_set_error("qf_xchg_push_wh: the handle has no tile-exchange communicator"); return QF_ERR_INVALID; }
    if (x->push_inline && !h->fuse_post) return QF_OK;      // the tail / update kernels have stored the tiles themselves
    if (h->xchg_ce) {
        // Copy-engine variant (QF_XCHG_PUSH=ce): the same rectangles as k_xchg_push_wh as 2-D device-to-device copies
        // into the peer mappings — memcpy nodes inside the step graph, no SM involved.  Not gated by the device flag: in
        // the graph they only run when the loop body runs; with eager launches a repeated copy of unchanged tiles is
        // harmless.  The upper-only decision needs the host copy of QfCtrl.skew_exact (qf_xchg_skew_check).
        QfP2P *p = reinterpret_cast<QfP2P *>(h->p2p);
        const QfArenaLayout l = arena_layout(h);
        const int N = h->N, hb = x->hb, G = x->nranks;
        const bool upper_only = x->upper_only && h->skew_host == 1;
        const size_t pitch = sizeof(double2) * (size_t)N;
        for (int q = 0; q < 2; ++q) {
            const int b = q ? 2 * G - 1 - x->rank : x->rank;
            const size_t r0 = (size_t)b * hb;
            for (int pr = 0; pr < G; ++pr) {
                if (pr == x->rank) continue;
                char *dst = p->peerArena[pr] + l.wh;
                const char *src = (const char *)h->Wh;
                const size_t offU = (r0 * N + r0) * sizeof(double2);
                QF_CUDA(cudaMemcpy2DAsync(dst + offU, pitch, src + offU, pitch, sizeof(double2) * (N - r0), hb, cudaMemcpyDeviceToDevice, st));
                if (!upper_only && r0 + hb < (size_t)N) {
                    const size_t offM = ((r0 + hb) * N + r0) * sizeof(double2);
                    QF_CUDA(cudaMemcpy2DAsync(dst + offM, pitch, src + offM, pitch, sizeof(double2) * hb, N - r0 - hb, cudaMemcpyDeviceToDevice, st));
                }
            }
        }
        return QF_OK;
    }
    k_xchg_push_wh<<<h->sm_count * 4, 256, 0, st>>>(*x, h->N, h->ctrl, gated ? 1 : 0);
    h->launches++;
    QF_CUDA(cudaGetLastError());
    return QF_OK;
}

// Lower triangle of W~ from the upper one, W~_ji = -conj(W~_ij), 32 x 32 tiles transposed through shared memory; a no-op
// unless the upper-only exchange is active (see k_xchg_push_wh).
__global__ void __launch_bounds__(256)
k_xchg_mirror_wh(const QfXchg x, double2 *__restrict__ Wh, int N, const QfCtrl *__restrict__ ctrl, int gated)
{
    if (gated && !ctrl[0].active) return;
    if (!(x.upper_only && ctrl[0].skew_exact)) return;
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bi > bj) return;
    __shared__ double2 T[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i = bi * 32 + ty + 8 * q, j = bj * 32 + tx;
        if (i < N && j < N) T[ty + 8 * q][tx] = __ldcg(Wh + (size_t)i * N + j);     // stored by a peer: not through L1
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int j = bj * 32 + ty + 8 * q, i = bi * 32 + tx;
        if (j < N && i < N && i < j) {
            const double2 v = T[tx][ty + 8 * q];
            Wh[(size_t)j * N + i] = make_double2(-v.x, v.y);
        }
    }
}

int qf_xchg_mirror_wh(qf_handle_s *h, bool gated, cudaStream_t st)
{
    const QfXchg *x = qf_xchg_desc(h);
    if (!x) { qf_set_error("qf_xchg_mirror_wh: the handle has no tile-exchange communicator"); return QF_ERR_INVALID; }
    if (!x->upper_only) return QF_OK;
    const int nb = (h->N + 31) / 32;
    k_xchg_mirror_wh<<<dim3(nb, nb), 256, 0, st>>>(*x, h->Wh, h->N, h->ctrl, gated ? 1 : 0);
    h->launches++;
    QF_CUDA(cudaGetLastError());
    return QF_OK;
}

// ctrl[0].skew_exact = 1 iff W_ji == -conj(W_ij) bit for bit for all i < j (the diagonal is free: the tail and the update
// only ever add purely imaginary numbers to it and mirror nothing onto it).
__global__ void k_xchg_skew_reset(QfCtrl *ctrl) { ctrl[0].skew_exact = 1; }
__global__ void __launch_bounds__(256)
k_xchg_skew_check(const double2 *__restrict__ W, int N, QfCtrl *ctrl)
{
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bi > bj) return;
    __shared__ double2 T[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int j = bj * 32 + ty + 8 * q, i = bi * 32 + tx;
        if (j < N && i < N) T[ty + 8 * q][tx] = W[(size_t)j * N + i];
    }
    __syncthreads();
    int bad = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i = bi * 32 + ty + 8 * q, j = bj * 32 + tx;
        if (i < N && j < N && i < j) {
            const double2 u = W[(size_t)i * N + j], l = T[tx][ty + 8 * q];
            // bit-for-bit: -0.0 and 0.0 count as equal (they add identically), NaNs never match
            if (!(l.x == -u.x && l.y == u.y)) bad = 1;
        }
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) ctrl[0].skew_exact = 0;
}

int qf_xchg_skew_check(qf_handle_s *h, const double2 *W, cudaStream_t st)
{
    const int nb = (h->N + 31) / 32;
    k_xchg_skew_reset<<<1, 1, 0, st>>>(h->ctrl);
    k_xchg_skew_check<<<dim3(nb, nb), 256, 0, st>>>(W, h->N, h->ctrl);
    h->launches += 2;
    QF_CUDA(cudaGetLastError());
    if (h->xchg_ce) {
        // the copy-engine variant enqueues different copies for the two cases: the host has to know
        QF_CUDA(cudaMemcpyAsync(h->ctrl_host, h->ctrl, sizeof(QfCtrl), cudaMemcpyDeviceToHost, st));
        QF_CUDA(cudaStreamSynchronize(st));
        const int skew = h->ctrl_host[0].skew_exact;
        if (skew != h->skew_host) qf_graph_destroy(h);      // the step graph bakes the copies in
        h->skew_host = skew;
    }
    return QF_OK;
}

// End of a call: every rank holds the up-to-date state on the tile pairs it owns; store them into every peer's Wst.
// One CTA per 32 x 32 tile pair (bi <= bj), as in k_update.
__global__ void __launch_bounds__(256)
k_xchg_push_state(const QfXchg x, int N)
{
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bi > bj || qf_owner_of_row(bi * 32, x.hb, x.nranks) != x.rank) return;
    const double2 *__restrict__ mine = x.peerWst[x.rank];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int half = 0; half < (bi == bj ? 1 : 2); ++half) {
        const int r0 = (half ? bj : bi) * 32, c0 = (half ? bi : bj) * 32;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int r = r0 + ty + 8 * q, c = c0 + tx;
            if (r < N && c < N) {
                const double2 v = mine[(size_t)r * N + c];
                for (int p = 0; p < x.nranks; ++p)
                    if (p != x.rank) x.peerWst[p][(size_t)r * N + c] = v;
            }
        }
    }
    __threadfence_system();
}

int qf_xchg_push_state(qf_handle_s *h, cudaStream_t st)
{
    const QfXchg *x = qf_xchg_desc(h);
    if (!x) { qf_set_error("qf_xchg_push_state: the handle has no tile-exchange communicator"); return QF_ERR_INVALID; }
    const int nb = (h->N + 31) / 32;
    k_xchg_push_state<<<dim3(nb, nb), 256, 0, st>>>(*x, h->N);
    h->launches++;
    QF_CUDA(cudaGetLastError());
    return QF_OK;
}

// Start of a row-sharded host call: this rank uploaded its two row blocks of the state; store them into every peer's Wst.
__global__ void __launch_bounds__(256)
k_xchg_push_rows(const QfXchg x, int N)
{
    const double2 *__restrict__ mine = x.peerWst[x.rank];
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < 2 * x.hb; w += warps) {
        const int blk = w < x.hb ? x.rank : 2 * x.nranks - 1 - x.rank;
        const int row = blk * x.hb + (w < x.hb ? w : w - x.hb);
        const double2 *__restrict__ s = mine + (size_t)row * N;
        for (int c = lane; c < N; c += 32 * 8) {
            double2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (c + 32 * u < N) v[u] = s[c + 32 * u];
            for (int p = 0; p < x.nranks; ++p) {
                if (p == x.rank) continue;
                double2 *__restrict__ d = x.peerWst[p] + (size_t)row * N;
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (c + 32 * u < N) d[c + 32 * u] = v[u];
            }
        }
    }
    __threadfence_system();
}

int qf_xchg_push_rows(qf_handle_s *h, cudaStream_t st)
{
    const QfXchg *x = qf_xchg_desc(h);
    if (!x) { qf_set_error("qf_xchg_push_rows: the handle has no tile-exchange communicator"); return QF_ERR_INVALID; }
    k_xchg_push_rows<<<h->sm_count * 2, 256, 0, st>>>(*x, h->N);
    h->launches++;
    QF_CUDA(cudaGetLastError());
    return QF_OK;
}

// Select the data path after qf_comm_p2p_import / qf_comm_attach_local: 0 pull all-gather, 1 tile exchange.
extern "C" int qf_comm_set_tile(qf_handle_t h, int enable)
{
    if (!h || !h->p2p || h->nranks < 2 || (h->comm_mode != 2 && h->comm_mode != 5)) {
        qf_set_error("qf_comm_set_tile: the handle has no peer-memory communicator");
        return QF_ERR_INVALID;
    }
    if (enable && !(h->N % (128 * h->nranks) == 0 && qf_gemm_can_fuse_post(h))) {
        qf_set_error("the tile-exchange path needs N divisible by 128*nranks (N=%d, nranks=%d) and the default GEMM kernel", h->N, h->nranks);
        return QF_ERR_UNSUPPORTED;
    }
    h->comm_mode = enable ? 5 : 2;
    qf_graph_destroy(h);   // the step graph bakes the data path in
    return QF_OK;
}

// 0: single GPU / emulated ranks, 1: NCCL all-gather, 2: pull all-gather, 5: tile exchange
extern "C" int qf_comm_mode(qf_handle_t h) { return h ? h->comm_mode : 0; }

// Single-GPU emulation of the row-sharded GEMM schedule of G ranks (tests): same tile lists, same permuted layout, no
// communication.
extern "C" int qf_set_emulated_ranks(qf_handle_t h, int nranks)
{
    if (!h || nranks < 1) { qf_set_error("qf_set_emulated_ranks: bad arguments"); return QF_ERR_INVALID; }
    if (h->comm_mode != 0) { qf_set_error("qf_set_emulated_ranks: handle already has a communicator"); return QF_ERR_INVALID; }
    if (h->batch != 1) { qf_set_error("row sharding needs batch == 1"); return QF_ERR_INVALID; }
    if (nranks > 1 && h->N % (2 * nranks) != 0) { qf_set_error("row sharding needs N divisible by 2*nranks"); return QF_ERR_INVALID; }
    h->rank = 0;
    h->nranks = nranks;
    qf_graph_destroy(h);
    return QF_OK;
}
