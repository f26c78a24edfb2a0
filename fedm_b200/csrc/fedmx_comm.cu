// Multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.
//
// Replaces the MPI layer hidden inside dolfin/PETSc (reference call sites: update_ghost_values() at
// fedm/functions.py:625-637, VecNorm/VecMDot allreduces and the MatMult ghost scatter behind
// fedm/functions.py:1047, df.norm at :1062-1064).  The mesh is partitioned by vertices on the host;
// every rank owns the first n_owned local vertices and keeps one ghost layer, so assembly is
// owner-computes with no communication.  Per Krylov iteration there is one halo exchange of the ghost
// entries of the input vector (grouped ncclSend/ncclRecv) and packed double allreduces.
#include <nccl.h>
#include <string.h>

#include "fedmx_internal.h"

#define FEDMX_CHECK_NCCL(ctx, expr)                                                   \
    do {                                                                              \
        ncclResult_t _r = (expr);                                                     \
        if (_r != ncclSuccess) {                                                      \
            (ctx)->err = std::string(#expr) + ": " + ncclGetErrorString(_r);          \
            return FEDMX_ENCCL;                                                       \
        }                                                                             \
    } while (0)

namespace {

__global__ void pack_kernel(const double *__restrict__ vec, const int32_t *__restrict__ nodes, int64_t n, int n_comp,
                            double *__restrict__ buf) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n * n_comp) return;
    const int64_t v = nodes[i / n_comp];
    buf[i] = vec[v * n_comp + (i % n_comp)];
}

__global__ void unpack_kernel(double *__restrict__ vec, const int32_t *__restrict__ nodes, int64_t n, int n_comp,
                              const double *__restrict__ buf) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n * n_comp) return;
    const int64_t v = nodes[i / n_comp];
    vec[v * n_comp + (i % n_comp)] = buf[i];
}

}  // namespace

extern "C" int fedmx_comm_unique_id(char id_out[128]) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return FEDMX_ENCCL;
    memcpy(id_out, &id, 128);
    return FEDMX_OK;
}

extern "C" int fedmx_comm_init(fedmx_ctx *ctx, int n_ranks, int rank, const char id[128]) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, n_ranks >= 1 && rank >= 0 && rank < n_ranks, "fedmx_comm_init: bad rank");
    FEDMX_REQUIRE(ctx, ctx->nccl_comm == nullptr, "fedmx_comm_init: already initialised");
    ncclUniqueId nid;
    memcpy(&nid, id, 128);
    ncclComm_t comm;
    FEDMX_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    FEDMX_CHECK_NCCL(ctx, ncclCommInitRank(&comm, n_ranks, nid, rank));
    ctx->nccl_comm = comm;
    ctx->n_ranks = n_ranks;
    ctx->rank = rank;
    return FEDMX_OK;
}

// one communicator per context: released with it (a context per Poisson_solver / project call would leak one each otherwise)
int fedmx_comm_destroy_internal(fedmx_ctx *ctx) {
    if (ctx->nccl_comm) {
        ncclCommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    return FEDMX_OK;
}

extern "C" int fedmx_halo_set(fedmx_ctx *ctx, int n_neighbors, const int32_t *h_neighbor_ranks, const int32_t *h_send_ptr,
                              const int32_t *h_send_nodes, const int32_t *h_recv_ptr, const int32_t *h_recv_nodes) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, ctx->n_vertices > 0, "fedmx_halo_set: set the mesh first");
    FEDMX_REQUIRE(ctx, n_neighbors >= 0, "fedmx_halo_set: bad neighbour count");
    HaloPlan &H = ctx->halo;
    H.n_neighbors = n_neighbors;
    H.ranks.assign(h_neighbor_ranks, h_neighbor_ranks + n_neighbors);
    H.send_ptr.assign(h_send_ptr, h_send_ptr + n_neighbors + 1);
    H.recv_ptr.assign(h_recv_ptr, h_recv_ptr + n_neighbors + 1);
    H.n_send = H.send_ptr[n_neighbors];
    H.n_recv = H.recv_ptr[n_neighbors];
    for (int64_t i = 0; i < H.n_send; ++i)
        FEDMX_REQUIRE(ctx, h_send_nodes[i] >= 0 && h_send_nodes[i] < ctx->n_owned, "fedmx_halo_set: send node not owned");
    for (int64_t i = 0; i < H.n_recv; ++i)
        FEDMX_REQUIRE(ctx, h_recv_nodes[i] >= ctx->n_owned && h_recv_nodes[i] < ctx->n_vertices, "fedmx_halo_set: recv node not a ghost");
    std::vector<int32_t> s(h_send_nodes, h_send_nodes + H.n_send), r(h_recv_nodes, h_recv_nodes + H.n_recv);
    int rc;
    if ((rc = fedmx_upload(ctx, &H.d_send_nodes, s))) return rc;
    if ((rc = fedmx_upload(ctx, &H.d_recv_nodes, r))) return rc;
    if ((rc = fedmx_dev_alloc(ctx, &H.d_send_buf, size_t(H.n_send) * FEDMX_MAX_EQ))) return rc;
    if ((rc = fedmx_dev_alloc(ctx, &H.d_recv_buf, size_t(H.n_recv) * FEDMX_MAX_EQ))) return rc;
    return FEDMX_OK;
}

int fedmx_halo_exchange_internal(fedmx_ctx *ctx, double *d_vec, int n_comp) {
    HaloPlan &H = ctx->halo;
    if (ctx->n_ranks <= 1 || H.n_neighbors == 0) return FEDMX_OK;
    ProfScope prof(ctx, FEDMX_PROF_COMM);
    if (ctx->peer.on && ctx->peer.halo_ready) return fedmx_peer_halo_exchange(ctx, d_vec, n_comp);   // NVLink peer stores
    FEDMX_REQUIRE(ctx, ctx->nccl_comm, "halo exchange: communicator not initialised");
    FEDMX_REQUIRE(ctx, n_comp >= 1 && n_comp <= FEDMX_MAX_EQ, "halo exchange: bad component count");
    const int threads = 256;
    if (H.n_send > 0) {
        pack_kernel<<<unsigned((H.n_send * n_comp + threads - 1) / threads), threads, 0, ctx->stream>>>(
            d_vec, H.d_send_nodes, H.n_send, n_comp, H.d_send_buf);
        ctx->launches++;
    }
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    FEDMX_CHECK_NCCL(ctx, ncclGroupStart());
    for (int i = 0; i < H.n_neighbors; ++i) {
        const int64_t ns = int64_t(H.send_ptr[i + 1] - H.send_ptr[i]) * n_comp;
        const int64_t nr = int64_t(H.recv_ptr[i + 1] - H.recv_ptr[i]) * n_comp;
        if (ns > 0) FEDMX_CHECK_NCCL(ctx, ncclSend(H.d_send_buf + int64_t(H.send_ptr[i]) * n_comp, ns, ncclDouble, H.ranks[i], comm, ctx->stream));
        if (nr > 0) FEDMX_CHECK_NCCL(ctx, ncclRecv(H.d_recv_buf + int64_t(H.recv_ptr[i]) * n_comp, nr, ncclDouble, H.ranks[i], comm, ctx->stream));
    }
    FEDMX_CHECK_NCCL(ctx, ncclGroupEnd());
    if (H.n_recv > 0) {
        unpack_kernel<<<unsigned((H.n_recv * n_comp + threads - 1) / threads), threads, 0, ctx->stream>>>(
            d_vec, H.d_recv_nodes, H.n_recv, n_comp, H.d_recv_buf);
        ctx->launches++;
    }
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    return FEDMX_OK;
}

int fedmx_global_sum(fedmx_ctx *ctx, double *d_buf, int n) {
    if (ctx->n_ranks <= 1) return FEDMX_OK;
    ProfScope prof(ctx, FEDMX_PROF_COMM);
    if (ctx->peer.on && (n <= FEDMX_PEER_AR_MAX || n <= ctx->peer.big_cap)) return fedmx_peer_allreduce(ctx, d_buf, n);
    FEDMX_REQUIRE(ctx, ctx->nccl_comm, "allreduce: communicator not initialised");
    FEDMX_CHECK_NCCL(ctx, ncclAllReduce(d_buf, d_buf, size_t(n), ncclDouble, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return FEDMX_OK;
}

extern "C" int fedmx_halo_exchange(fedmx_ctx *ctx, double *d_vec, int n_comp) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, d_vec, "fedmx_halo_exchange: null vector");
    return fedmx_halo_exchange_internal(ctx, d_vec, n_comp);
}

extern "C" int fedmx_allreduce_sum(fedmx_ctx *ctx, double *d_buf, int n) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, d_buf && n > 0, "fedmx_allreduce_sum: bad arguments");
    return fedmx_global_sum(ctx, d_buf, n);
}
