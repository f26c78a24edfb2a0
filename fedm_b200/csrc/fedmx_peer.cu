// Peer-memory collectives over NVLink/NVSwitch (one process per GPU, CUDA IPC): the Krylov loop's halo exchange and
// allreduces as plain kernels that store straight into the neighbours' HBM and hand over with release/acquire flags.
//
// Why: at 10 M DOF on 8 GPUs every kernel of the Newton-Krylov loop runs for 10-40 us, so the ~25 collectives per
// Newton iteration (reference: the MPI scatters/allreduces hidden in PETSc's MatMult, VecMDot, VecNorm and PCApply
// behind fedm/functions.py:1047) are latency-bound.  NCCL send/recv groups and small allreduces cost 20-40 us each;
// a peer store + flag costs an NVLink round trip (~2-4 us).  NCCL stays as the set-up / fallback transport.
//
// Symmetric heap: every rank cudaMalloc's one block with the SAME layout, exports it with cudaIpcGetMemHandle, the host
// all-gathers the handles (torch.distributed) and every rank maps all peers' blocks.  Layout (offsets identical on all
// ranks): flags | small-allreduce slots [2][R][128] | halo staging [2][halo_cap] | big-allreduce in / out.
//
// Protocols (all epochs are monotonically increasing 64-bit counters, identical on all ranks because every rank issues
// the same sequence of collectives):
//  * halo exchange e: push kernel packs the owned boundary values directly into the neighbour's staging buffer
//    [e & 1] (peer stores), fences, and its last block releases flag[me] = e on every neighbour; the wait kernel
//    acquires flag[nb] >= e for all neighbours and unpacks the staging buffer into the ghost entries of the vector.
//    Staging parity makes this safe for any vector: a sender can only reach exchange e+2 after it has seen the
//    receiver's push of e+1, which follows the receiver's unpack of e in stream order.
//  * small allreduce (n <= 128): every rank stores its partial vector into slot[me] of every peer, releases a flag,
//    acquires all flags and sums the slots in rank order -> bit-identical results on all ranks, one launch.
//  * big allreduce (restricted multigrid residual): barrier, then rank r pulls chunk r of every peer's input, sums in
//    rank order and pushes the chunk into every peer's output; flags; copy out.
// Every spin loop has a clock64() timeout that raises an error flag instead of hanging the GPU.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "fedmx_internal.h"

namespace {

constexpr int kArMax = FEDMX_PEER_AR_MAX;
// Spin time-out in SM clocks.  Ordinary rank skew (host-side multigrid set-up, JIT compilation, rank-0 output) can be many
// seconds, so the default is generous (FEDMX_PEER_TIMEOUT_S, default 120 s); a time-out is FATAL for the run -- the flag is
// sticky and every later library call on the context fails (ranks must not continue on diverging control flow).
__device__ long long g_spin_timeout = 240000000000LL;

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool spin_until(const unsigned long long *flag, unsigned long long epoch, int *err) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < epoch) {
        if (clock64() - t0 > g_spin_timeout) { *err = 1; return false; }
    }
    return true;
}

struct HaloArgs {
    int n_neighbors, n_comp;
    int send_ptr[FEDMX_MAX_RANKS + 1], recv_ptr[FEDMX_MAX_RANKS + 1];
    int nb_rank[FEDMX_MAX_RANKS];
    long long remote_off[FEDMX_MAX_RANKS];            // node offset of my segment in the neighbour's staging buffer
    double *remote_buf[FEDMX_MAX_RANKS];              // neighbour's staging buffer of this parity
    unsigned long long *remote_flag[FEDMX_MAX_RANKS]; // neighbour's halo flag of this rank
    const unsigned long long *my_flags;               // [R]
    const double *my_buf;                             // my staging buffer of this parity
    const int32_t *send_nodes, *recv_nodes;
    unsigned long long epoch;
    unsigned int *counter;
    int *err;
};

__global__ void __launch_bounds__(256) halo_push_kernel(const double *__restrict__ vec, HaloArgs a) {
    const int total = a.send_ptr[a.n_neighbors] * a.n_comp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int node = i / a.n_comp, c = i - node * a.n_comp;
        int nb = 0;
        while (node >= a.send_ptr[nb + 1]) ++nb;
        const long long pos = a.remote_off[nb] + (node - a.send_ptr[nb]);
        a.remote_buf[nb][pos * a.n_comp + c] = vec[(long long)a.send_nodes[node] * a.n_comp + c];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(a.counter, 1u);
        if (prev == gridDim.x - 1) {
            *a.counter = 0;
            __threadfence_system();
            for (int nb = 0; nb < a.n_neighbors; ++nb) st_release_sys(a.remote_flag[nb], a.epoch);
        }
    }
}

__global__ void __launch_bounds__(256) halo_wait_unpack_kernel(double *__restrict__ vec, HaloArgs a) {
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    if (threadIdx.x < a.n_neighbors && a.recv_ptr[threadIdx.x + 1] > a.recv_ptr[threadIdx.x]) {
        if (!spin_until(a.my_flags + a.nb_rank[threadIdx.x], a.epoch, a.err)) ok = 0;
    }
    __syncthreads();
    if (!ok) return;
    const int total = a.recv_ptr[a.n_neighbors] * a.n_comp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int node = i / a.n_comp, c = i - node * a.n_comp;
        vec[(long long)a.recv_nodes[node] * a.n_comp + c] = __ldcv(a.my_buf + i);
    }
}

struct ArArgs {
    int n, n_ranks, rank;
    double *slots[FEDMX_MAX_RANKS];                   // peer r's slot array of this parity: [R][kArMax]
    unsigned long long *flags[FEDMX_MAX_RANKS];       // peer r's flag of this rank
    const double *my_slots;
    const unsigned long long *my_flags;
    unsigned long long epoch;
    int *err;
};

__global__ void __launch_bounds__(256) allreduce_small_kernel(double *__restrict__ buf, ArArgs a) {
    __shared__ int ok;
    const int t = threadIdx.x;
    if (t == 0) ok = 1;
    if (t < a.n) {
        const double v = buf[t];
        for (int r = 0; r < a.n_ranks; ++r) a.slots[r][a.rank * kArMax + t] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (t < a.n_ranks) {
        st_release_sys(a.flags[t], a.epoch);
        if (!spin_until(a.my_flags + t, a.epoch, a.err)) ok = 0;
    }
    __syncthreads();
    if (!ok) return;
    if (t < a.n) {
        double s = 0.0;
        for (int r = 0; r < a.n_ranks; ++r) s += __ldcv(a.my_slots + r * kArMax + t);
        buf[t] = s;
    }
}

struct BigArgs {
    long long n;
    int n_ranks, rank;
    const double *in[FEDMX_MAX_RANKS];                // every peer's input buffer
    double *out[FEDMX_MAX_RANKS];                     // every peer's output buffer
    unsigned long long *flags[FEDMX_MAX_RANKS];       // peer r's big flag of this rank
    const unsigned long long *my_flags;
    const double *my_out;
    unsigned long long epoch_ready, epoch;            // "my input is complete" / "my chunk is in your output"
    unsigned int *counter;
    int *err;
};

__global__ void __launch_bounds__(256) big_reduce_push_kernel(BigArgs a) {
    // barrier: my input (written by the preceding kernel of this stream) is complete -> tell every peer, then wait until
    // every peer's input is complete.  All blocks of the (<= 148-block) grid are co-resident, so spinning is safe.
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    if (threadIdx.x < a.n_ranks) {
        if (blockIdx.x == 0) { __threadfence_system(); st_release_sys(a.flags[threadIdx.x], a.epoch_ready); }
        if (!spin_until(a.my_flags + threadIdx.x, a.epoch_ready, a.err)) ok = 0;
    }
    __syncthreads();
    if (!ok) return;
    const long long chunk = (a.n + a.n_ranks - 1) / a.n_ranks;
    const long long lo = chunk * a.rank, hi = min(a.n, lo + chunk);
    for (long long i = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hi; i += (long long)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < a.n_ranks; ++r) s += __ldcv(a.in[r] + i);
        for (int r = 0; r < a.n_ranks; ++r) a.out[r][i] = s;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(a.counter, 1u);
        if (prev == gridDim.x - 1) {
            *a.counter = 0;
            __threadfence_system();
            for (int r = 0; r < a.n_ranks; ++r) st_release_sys(a.flags[r], a.epoch);
        }
    }
}

__global__ void __launch_bounds__(256) big_wait_copy_kernel(double *__restrict__ buf, BigArgs a) {
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    if (threadIdx.x < a.n_ranks) {
        if (!spin_until(a.my_flags + threadIdx.x, a.epoch, a.err)) ok = 0;
    }
    __syncthreads();
    if (!ok) return;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x)
        buf[i] = __ldcv(a.my_out + i);
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

// ---- set-up ---------------------------------------------------------------------------------------------------------
extern "C" int fedmx_peer_create(fedmx_ctx *ctx, int64_t halo_capacity, int64_t big_capacity, char handle_out[64]) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, ctx->n_ranks > 1 && ctx->n_ranks <= FEDMX_MAX_RANKS, "fedmx_peer_create: call fedmx_comm_init first (2..8 ranks)");
    FEDMX_REQUIRE(ctx, !ctx->peer.base, "fedmx_peer_create: already created");
    FEDMX_REQUIRE(ctx, halo_capacity >= 0 && big_capacity >= 0 && handle_out, "fedmx_peer_create: bad arguments");
    PeerState &P = ctx->peer;
    const int R = FEDMX_MAX_RANKS;
    size_t off = 0;
    P.off_flags = off;      off += align_up(size_t(4) * R * sizeof(unsigned long long), 256);     // halo | ar | big | probe
    P.off_ar_slots = off;   off += align_up(size_t(2) * R * kArMax * sizeof(double), 256);
    P.off_halo = off;       off += align_up(size_t(2) * size_t(halo_capacity) * sizeof(double), 256);
    P.off_big_in = off;     off += align_up(size_t(big_capacity) * sizeof(double), 256);
    P.off_big_out = off;    off += align_up(size_t(big_capacity) * sizeof(double), 256);
    // a large block is always its own allocation, so the IPC handle maps exactly this base address on the peers
    P.bytes = std::max<size_t>(off, size_t(64) << 20);
    P.halo_cap = halo_capacity;
    P.big_cap = big_capacity;
    FEDMX_CHECK_CUDA(ctx, cudaMalloc((void **)&P.base, P.bytes));
    FEDMX_CHECK_CUDA(ctx, cudaMemset(P.base, 0, P.bytes));
    // probe word: my rank + 1, checked by the peers after mapping
    const unsigned long long probe = (unsigned long long)ctx->rank + 1;
    FEDMX_CHECK_CUDA(ctx, cudaMemcpy(P.base + P.off_flags + size_t(3) * R * sizeof(unsigned long long), &probe, sizeof(probe), cudaMemcpyHostToDevice));
    FEDMX_CHECK_CUDA(ctx, cudaMalloc((void **)&P.d_counter, 2 * sizeof(unsigned int)));
    FEDMX_CHECK_CUDA(ctx, cudaMemset(P.d_counter, 0, 2 * sizeof(unsigned int)));
    {
        const char *env = getenv("FEDMX_PEER_TIMEOUT_S");
        const double secs = env ? atof(env) : 120.0;
        const long long clocks = (long long)(std::max(1.0, secs) * 2.0e9);
        FEDMX_CHECK_CUDA(ctx, cudaMemcpyToSymbol(g_spin_timeout, &clocks, sizeof(clocks)));
    }
    FEDMX_CHECK_CUDA(ctx, cudaHostAlloc((void **)&P.h_err, sizeof(int), cudaHostAllocMapped));
    *P.h_err = 0;
    FEDMX_CHECK_CUDA(ctx, cudaHostGetDevicePointer((void **)&P.d_err, P.h_err, 0));
    cudaIpcMemHandle_t h;
    FEDMX_CHECK_CUDA(ctx, cudaIpcGetMemHandle(&h, P.base));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
    memcpy(handle_out, &h, 64);
    FEDMX_CHECK_CUDA(ctx, cudaDeviceSynchronize());
    return FEDMX_OK;
}

extern "C" int fedmx_peer_open(fedmx_ctx *ctx, const char *handles /* [n_ranks][64] */) {
    if (!ctx) return FEDMX_EINVAL;
    PeerState &P = ctx->peer;
    FEDMX_REQUIRE(ctx, P.base && handles, "fedmx_peer_open: create the heap first");
    const int R = FEDMX_MAX_RANKS;
    for (int r = 0; r < ctx->n_ranks; ++r) {
        if (r == ctx->rank) { P.bases[r] = P.base; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + size_t(r) * 64, 64);
        void *p = nullptr;
        FEDMX_CHECK_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        P.bases[r] = (char *)p;
        unsigned long long probe = 0;
        FEDMX_CHECK_CUDA(ctx, cudaMemcpy(&probe, P.bases[r] + P.off_flags + size_t(3) * R * sizeof(unsigned long long), sizeof(probe), cudaMemcpyDeviceToHost));
        FEDMX_REQUIRE(ctx, probe == (unsigned long long)r + 1, "fedmx_peer_open: mapped peer heap does not carry the peer's probe word");
    }
    P.opened = true;
    P.on = true;
    return FEDMX_OK;
}

extern "C" int fedmx_peer_halo_set(fedmx_ctx *ctx, int n_neighbors, const int32_t *h_remote_offsets) {
    if (!ctx) return FEDMX_EINVAL;
    PeerState &P = ctx->peer;
    FEDMX_REQUIRE(ctx, P.on, "fedmx_peer_halo_set: peer heap not open");
    FEDMX_REQUIRE(ctx, n_neighbors == ctx->halo.n_neighbors && n_neighbors <= FEDMX_MAX_RANKS && (n_neighbors == 0 || h_remote_offsets),
                  "fedmx_peer_halo_set: neighbour count differs from fedmx_halo_set");
    FEDMX_REQUIRE(ctx, ctx->halo.n_recv * FEDMX_MAX_EQ <= P.halo_cap, "fedmx_peer_halo_set: staging buffer too small");
    P.remote_off.assign(h_remote_offsets, h_remote_offsets + n_neighbors);
    P.halo_ready = true;
    return FEDMX_OK;
}

extern "C" int fedmx_peer_enable(fedmx_ctx *ctx, int on) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, !on || ctx->peer.opened, "fedmx_peer_enable: no peer heap");
    ctx->peer.on = on != 0;
    return FEDMX_OK;
}

extern "C" int fedmx_peer_error(fedmx_ctx *ctx) {
    if (!ctx || !ctx->peer.h_err) return 0;
    return *ctx->peer.h_err;
}

int fedmx_peer_destroy(fedmx_ctx *ctx) {
    PeerState &P = ctx->peer;
    if (!P.base) return FEDMX_OK;
    for (int r = 0; r < ctx->n_ranks; ++r)
        if (r != ctx->rank && P.bases[r]) cudaIpcCloseMemHandle(P.bases[r]);
    cudaFree(P.base);
    cudaFree(P.d_counter);
    cudaFreeHost(P.h_err);
    P = PeerState{};
    return FEDMX_OK;
}

// ---- collectives ----------------------------------------------------------------------------------------------------
static inline unsigned long long *flag_ptr(const PeerState &P, int owner_rank, int kind, int index) {
    return (unsigned long long *)(P.bases[owner_rank] + P.off_flags) + size_t(kind) * FEDMX_MAX_RANKS + index;
}

int fedmx_peer_halo_exchange(fedmx_ctx *ctx, double *d_vec, int n_comp) {
    PeerState &P = ctx->peer;
    HaloPlan &H = ctx->halo;
    FEDMX_REQUIRE(ctx, P.on && P.halo_ready, "peer halo exchange: not set up");
    FEDMX_REQUIRE(ctx, n_comp >= 1 && n_comp <= FEDMX_MAX_EQ, "halo exchange: bad component count");
    const unsigned long long epoch = ++P.halo_epoch;
    const int par = int(epoch & 1);
    HaloArgs a{};
    a.n_neighbors = H.n_neighbors;
    a.n_comp = n_comp;
    for (int i = 0; i <= H.n_neighbors; ++i) { a.send_ptr[i] = H.send_ptr[i]; a.recv_ptr[i] = H.recv_ptr[i]; }
    for (int i = 0; i < H.n_neighbors; ++i) {
        const int q = H.ranks[i];
        a.nb_rank[i] = q;
        a.remote_off[i] = P.remote_off[i];
        a.remote_buf[i] = (double *)(P.bases[q] + P.off_halo) + size_t(par) * P.halo_cap;
        a.remote_flag[i] = flag_ptr(P, q, 0, ctx->rank);
    }
    a.my_flags = flag_ptr(P, ctx->rank, 0, 0);
    a.my_buf = (const double *)(P.base + P.off_halo) + size_t(par) * P.halo_cap;
    a.send_nodes = H.d_send_nodes;
    a.recv_nodes = H.d_recv_nodes;
    a.epoch = epoch;
    a.counter = P.d_counter;
    a.err = P.d_err;
    const int ns = int(H.n_send) * n_comp, nr = int(H.n_recv) * n_comp;
    const unsigned bs = unsigned(std::min(16, std::max(1, (ns + 255) / 256)));
    const unsigned br = unsigned(std::min(16, std::max(1, (nr + 255) / 256)));
    halo_push_kernel<<<bs, 256, 0, ctx->stream>>>(d_vec, a);
    halo_wait_unpack_kernel<<<br, 256, 0, ctx->stream>>>(d_vec, a);
    ctx->launches += 2;
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    return FEDMX_OK;
}

int fedmx_peer_allreduce(fedmx_ctx *ctx, double *d_buf, int64_t n) {
    PeerState &P = ctx->peer;
    FEDMX_REQUIRE(ctx, P.on, "peer allreduce: not set up");
    if (n <= kArMax) {
        const unsigned long long epoch = ++P.ar_epoch;
        const int par = int(epoch & 1);
        ArArgs a{};
        a.n = int(n);
        a.n_ranks = ctx->n_ranks;
        a.rank = ctx->rank;
        for (int r = 0; r < ctx->n_ranks; ++r) {
            a.slots[r] = (double *)(P.bases[r] + P.off_ar_slots) + size_t(par) * FEDMX_MAX_RANKS * kArMax;
            a.flags[r] = flag_ptr(P, r, 1, ctx->rank);
        }
        a.my_slots = (const double *)(P.base + P.off_ar_slots) + size_t(par) * FEDMX_MAX_RANKS * kArMax;
        a.my_flags = flag_ptr(P, ctx->rank, 1, 0);
        a.epoch = epoch;
        a.err = P.d_err;
        allreduce_small_kernel<<<1, 256, 0, ctx->stream>>>(d_buf, a);
        ctx->launches++;
        FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
        return FEDMX_OK;
    }
    FEDMX_REQUIRE(ctx, n <= P.big_cap, "peer allreduce: vector longer than the big-allreduce buffers");
    BigArgs a{};
    a.n = n;
    a.n_ranks = ctx->n_ranks;
    a.rank = ctx->rank;
    for (int r = 0; r < ctx->n_ranks; ++r) {
        a.in[r] = (const double *)(P.bases[r] + P.off_big_in);
        a.out[r] = (double *)(P.bases[r] + P.off_big_out);
        a.flags[r] = flag_ptr(P, r, 2, ctx->rank);
    }
    a.my_flags = flag_ptr(P, ctx->rank, 2, 0);
    a.my_out = (const double *)(P.base + P.off_big_out);
    a.counter = P.d_counter + 1;
    a.err = P.d_err;
    if ((const char *)d_buf != P.base + P.off_big_in)      // callers that produce their vector in the heap skip the copy
        FEDMX_CHECK_CUDA(ctx, cudaMemcpyAsync(P.base + P.off_big_in, d_buf, size_t(n) * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    a.epoch_ready = ++P.big_epoch;
    a.epoch = ++P.big_epoch;
    const long long chunk = (n + ctx->n_ranks - 1) / ctx->n_ranks;
    const unsigned b1 = unsigned(std::min<long long>(148, std::max<long long>(1, (chunk + 255) / 256)));
    big_reduce_push_kernel<<<b1, 256, 0, ctx->stream>>>(a);
    const unsigned b2 = unsigned(std::min<long long>(296, std::max<long long>(1, (n + 255) / 256)));
    big_wait_copy_kernel<<<b2, 256, 0, ctx->stream>>>(d_buf, a);
    ctx->launches += 2;
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    return FEDMX_OK;
}

// Input buffer of the big allreduce: a vector produced directly here is reduced without a staging copy.
double *fedmx_peer_big_buffer(fedmx_ctx *ctx, int64_t n) {
    PeerState &P = ctx->peer;
    if (!P.opened || n > P.big_cap) return nullptr;
    return (double *)(P.base + P.off_big_in);
}
