// Two-stage preconditioner: algebraic multigrid V-cycle on the potential block + node-block Jacobi (sm_100a).
//
// The hierarchy (smoothed aggregation, fedm_b200/amg.py) is built once on the host -- the Phi-Phi block
// 2 pi r grad(u).grad(v) (reference: fedm/functions.py:401) depends on the mesh only -- and uploaded level by
// level; every operator (A_l, P_l, R_l) is stored in scalar SELL-32 so that all V-cycle kernels are coalesced,
// HBM-bound streaming kernels (one thread per row).  Preconditioner application for r = (r_T, r_Phi):
//     z     = Dinv r                     node-block Jacobi (all components; captures the node-local coupling)
//     z_Phi = Vcycle(r_Phi)              multigrid on the elliptic component
// i.e. a block-diagonal two-stage (CPR-like) preconditioner.  PETSc's role here is PCApply inside KSPGMRES
// (reference call site fedm/functions.py:1047); results do not depend on the preconditioner, only iteration counts.
#include <algorithm>
#include <stdlib.h>
#include <string>

#include <cooperative_groups.h>
#include <cuda_fp16.h>

#include "fedmx_internal.h"

namespace cg = cooperative_groups;

int fedmx_halo_exchange_internal(fedmx_ctx *ctx, double *d_vec, int n_comp);
int fedmx_global_sum(fedmx_ctx *ctx, double *d_buf, int n);

namespace {

enum { MODE_MULT = 0, MODE_RESID = 1, MODE_JACOBI = 2, MODE_ADD = 3, MODE_JACOBI_OUT = 4, MODE_FIRST_RESID = 5 };

// MODE_JACOBI_OUT: the last sweep of the cycle writes its result straight into component `yo` of the node-interleaved
// preconditioned vector (y[row*ys + yo]) instead of the level's iterate.
template <int MODE, bool COMPACT = false>
__global__ void __launch_bounds__(128) ssell_kernel(ScalarSell A, const double *__restrict__ x, const double *__restrict__ b,
                                                    const double *__restrict__ dinv, double omega, double *__restrict__ y,
                                                    int ys = 1, int yo = 0) {
    const int64_t row = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t s = row / FEDMX_SLICE;
    const int lane = int(row % FEDMX_SLICE);
    if (s >= A.n_slices) return;
    const int width = A.slice_width[s];
    const int64_t base = A.slice_ptr[s] + lane;
    // entry k of this row: (value, column); COMPACT: FP16 value scaled by the row's diagonal, 16-bit column offset
    auto entry = [&](int64_t e, double &v, int64_t &c) {
        if (COMPACT) {
            v = double(__half2float(__ushort_as_half(A.val16[e])));
            c = row + A.col16[e];
        } else {
            v = double(A.val[e]);
            c = A.col[e];
        }
    };
    double t = 0.0;
    if (MODE == MODE_FIRST_RESID) {
        // zero-guess sweep fused with the residual: x = omega Dinv b is formed on the fly at the neighbours (b and dinv are
        // gathered instead of x), x itself is written for the later prolongation:  y = b - A (omega Dinv b), ys/yo unused
        for (int k = 0; k < width; ++k) {
            double v;
            int64_t c;
            entry(base + int64_t(k) * FEDMX_SLICE, v, c);
            t = fma(v, omega * __ldg(dinv + c) * __ldg(b + c), t);
        }
        if (row >= A.n_rows) return;
        if (COMPACT) t *= A.rdiag[row];
        const_cast<double *>(x)[row] = omega * dinv[row] * b[row];
        y[row] = b[row] - t;
        return;
    }
    for (int k = 0; k < width; ++k) {
        double v;
        int64_t c;
        entry(base + int64_t(k) * FEDMX_SLICE, v, c);
        t = fma(v, __ldg(x + c), t);
    }
    if (row >= A.n_rows) return;
    if (COMPACT) t *= A.rdiag[row];
    if (MODE == MODE_MULT) y[row] = t;
    else if (MODE == MODE_RESID) y[row] = b[row] - t;
    else if (MODE == MODE_JACOBI) y[row] = x[row] + omega * dinv[row] * (b[row] - t);
    else if (MODE == MODE_JACOBI_OUT) y[row * ys + yo] = x[row] + omega * dinv[row] * (b[row] - t);
    else y[row] += t;
}

// y = M x with LPR lanes per row (restriction rows are 40-60 entries long: one thread per row is latency-bound)
template <int LPR>
__global__ void __launch_bounds__(128) ssell_mult_split_kernel(ScalarSell A, const double *__restrict__ x, double *__restrict__ y) {
    const int64_t w = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t row = w / LPR;
    const int part = int(w % LPR);
    double t = 0.0;
    if (row < A.n_rows) {
        const int64_t s = row / FEDMX_SLICE;
        const int lane = int(row % FEDMX_SLICE);
        const int width = A.slice_width[s];
        const int64_t base = A.slice_ptr[s] + lane;
        for (int k = part; k < width; k += LPR) {
            const int64_t e = base + int64_t(k) * FEDMX_SLICE;
            t = fma(double(A.val[e]), __ldg(x + A.col[e]), t);
        }
    }
#pragma unroll
    for (int o = LPR >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (part == 0 && row < A.n_rows) y[row] = t;
}

__global__ void first_sweep_kernel(const double *__restrict__ b, const double *__restrict__ dinv, double omega,
                                   double *__restrict__ x, int64_t n) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) x[i] = omega * dinv[i] * b[i];
}

// y = M b for a small dense M (n <= 4096): one warp per row
__global__ void dense_matvec_kernel(const double *__restrict__ M, const double *__restrict__ b, double *__restrict__ y, int n) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    double acc = 0.0;
    for (int j = lane; j < n; j += 32) acc = fma(M[int64_t(row) * n + j], b[j], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[row] = acc;
}

// z = Dinv r (node-block Jacobi) and, in the same pass, rhs_c[i] = r[NEQ*i + comp] (right-hand side of the AMG stage)
template <int NEQ>
__global__ void __launch_bounds__(128) dinv_extract_kernel(SellView S, const double *__restrict__ dinv, const double *__restrict__ r,
                                                           int comp, double *__restrict__ z, double *__restrict__ rhs_c) {
    constexpr int NB = NEQ * NEQ;
    const int64_t row = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (row >= S.n_rows) return;
    const int64_t s = row / FEDMX_SLICE;
    const int lane = int(row % FEDMX_SLICE);
    const double *d = dinv + s * FEDMX_SLICE * NB + lane;
    const int64_t node = S.row_node[row];
    double in[NEQ];
#pragma unroll
    for (int l = 0; l < NEQ; ++l) in[l] = r[int64_t(NEQ) * node + l];
#pragma unroll
    for (int i = 0; i < NEQ; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int l = 0; l < NEQ; ++l) acc = fma(d[(i * NEQ + l) * FEDMX_SLICE], in[l], acc);
        z[int64_t(NEQ) * node + i] = acc;
        if (i == comp) rhs_c[node] = in[i];
    }
}

// ---- fused coarse levels: one CTA walks levels [first, last] of the V-cycle with __syncthreads() between the phases
struct CoarseLevelDev {
    int64_t n, n_coarse;
    double omega;
    ScalarSell A, P, R;
    const double *dinv;
    double *b, *x, *x2, *r;
};

// Partial row product of a scalar SELL-32 matrix: lane `part` of the LPR lanes that share the row takes the stored
// entries k = part, part+LPR, ...; the caller sums the parts with shuffles.  Coarse levels are latency-bound (one thread
// walking a 40-60 entry restriction row costs ~15 us regardless of the level size), so rows are split across lanes there.
__device__ __forceinline__ double ssell_row_part(const ScalarSell &M, int64_t row, int part, int lpr, const double *x) {
    const int64_t s = row / FEDMX_SLICE;
    const int lane = int(row % FEDMX_SLICE);
    const int width = M.slice_width[s];
    const int64_t base = M.slice_ptr[s] + lane;
    double t = 0.0;
    if (M.val16) {                                               // compact finest operator (see ScalarSell)
        for (int k = part; k < width; k += lpr) {
            const int64_t e = base + int64_t(k) * FEDMX_SLICE;
            t = fma(double(__half2float(__ushort_as_half(M.val16[e]))), x[row + M.col16[e]], t);
        }
        return t * M.rdiag[row];
    }
    for (int k = part; k < width; k += lpr) {
        const int64_t e = base + int64_t(k) * FEDMX_SLICE;
        t = fma(double(M.val[e]), x[M.col[e]], t);
    }
    return t;
}

__device__ __forceinline__ double group_sum(double t, int lpr) {
    for (int o = lpr >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
}

__device__ __forceinline__ int pick_lpr(int64_t n, int nt, int cap) {
    int l = 1;
    while (l < cap && int64_t(2 * l) * n <= nt) l <<= 1;
    return l;
}

// One thread-block CLUSTER walks levels [first, last] of the V-cycle; the phases are separated by hardware cluster
// barriers (barrier.cluster arrive.release / wait.acquire, ~0.3 us) instead of kernel boundaries (~3-5 us each, and
// 5 launches per level).  tid/nt run over the whole cluster.  `op(row, sum)` consumes the finished row product.
template <class Op>
__device__ __forceinline__ void for_rows(const ScalarSell &M, int64_t n, const double *x, int tid, int nt, int lpr_cap, Op op) {
    const int lpr = pick_lpr(n, nt, lpr_cap);
    const int64_t items = ((n * lpr + 31) / 32) * 32;            // whole warps iterate together (shuffles)
    for (int64_t w = tid; w < items; w += nt) {
        const int64_t row = w / lpr;
        const int part = int(w % lpr);
        double t = row < n ? ssell_row_part(M, row, part, lpr, x) : 0.0;
        t = group_sum(t, lpr);
        if (part == 0 && row < n) op(row, t);
    }
}

// Body shared by the two fused variants: SYNC is a cluster barrier (one 8-CTA cluster, small levels) or a cooperative grid
// barrier (whole GPU, mid levels).  Phases per level: residual | restriction (+ the zero-guess sweep of the next level) on the
// way down, prolongation | Jacobi sweep on the way up; iterates ping-pong between x and x2 instead of being copied
// (only the finest fused level, which the caller reads, is copied back into x).
template <class Sync>
__device__ __forceinline__ void coarse_vcycle_body(Sync sync, const int tid, const int nt, const CoarseLevelDev *L, int nl, int nu,
                                                   const double *__restrict__ cinv) {
    double *xa[12], *xb[12];                                  // current iterate / scratch per level
    for (int l = 0; l < nl; ++l) { xa[l] = L[l].x; xb[l] = L[l].x2; }
    {   // zero-guess sweep of the finest fused level (its right-hand side was produced by the preceding kernel)
        const CoarseLevelDev &lv = L[0];
        if (nl > 1) {
            for (int64_t i = tid; i < lv.n; i += nt) xa[0][i] = lv.omega * lv.dinv[i] * lv.b[i];
            sync();
        }
    }
    for (int l = 0; l < nl - 1; ++l) {
        const CoarseLevelDev &lv = L[l];
        for (int sweep = 1; sweep < nu; ++sweep) {
            const double *xc = xa[l];
            double *xn = xb[l];
            for_rows(lv.A, lv.n, xc, tid, nt, 8, [&](int64_t i, double t) { xn[i] = xc[i] + lv.omega * lv.dinv[i] * (lv.b[i] - t); });
            sync();
            xb[l] = xa[l];
            xa[l] = xn;
        }
        for_rows(lv.A, lv.n, xa[l], tid, nt, 8, [&](int64_t i, double t) { lv.r[i] = lv.b[i] - t; });
        sync();
        const CoarseLevelDev &lc = L[l + 1];
        const bool sweep_next = l + 1 < nl - 1;
        double *xnext = xa[l + 1];
        for_rows(lv.R, lv.n_coarse, lv.r, tid, nt, 32, [&](int64_t i, double t) {
            lc.b[i] = t;
            if (sweep_next) xnext[i] = lc.omega * lc.dinv[i] * t;
        });
        sync();
    }
    {   // coarsest level: dense inverse, one warp per row
        const CoarseLevelDev &lc = L[nl - 1];
        const int n = int(lc.n);
        const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
        for (int row = warp; row < n; row += nw) {
            double acc = 0.0;
            for (int j = lane; j < n; j += 32) acc = fma(cinv[int64_t(row) * n + j], lc.b[j], acc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) xa[nl - 1][row] = acc;
        }
        sync();
    }
    for (int l = nl - 2; l >= 0; --l) {
        const CoarseLevelDev &lv = L[l];
        double *xc = xa[l];
        for_rows(lv.P, lv.n, xa[l + 1], tid, nt, 4, [&](int64_t i, double t) { xc[i] += t; });
        sync();
        for (int sweep = 0; sweep < nu; ++sweep) {
            const double *xo = xa[l];
            double *xn = xb[l];
            for_rows(lv.A, lv.n, xo, tid, nt, 8, [&](int64_t i, double t) { xn[i] = xo[i] + lv.omega * lv.dinv[i] * (lv.b[i] - t); });
            sync();
            xb[l] = xa[l];
            xa[l] = xn;
        }
    }
    if (xa[0] != L[0].x) {                                     // the caller reads the result from L[0].x
        const double *src = xa[0];
        double *dst = L[0].x;
        for (int64_t i = tid; i < L[0].n; i += nt) dst[i] = src[i];
    }
}

__global__ void __launch_bounds__(1024) coarse_vcycle_kernel(const CoarseLevelDev *L, int nl, int nu, const double *__restrict__ cinv) {
    cg::cluster_group cluster = cg::this_cluster();
    const int nt = int(cluster.num_blocks()) * blockDim.x;
    const int tid = int(cluster.block_rank()) * blockDim.x + threadIdx.x;
    coarse_vcycle_body([&]() { cluster.sync(); }, tid, nt, L, nl, nu, cinv);
}

// Whole-GPU variant (cooperative launch, all CTAs co-resident): the mid levels (62 k - 555 k rows) need every SM's load/store
// bandwidth, and a grid barrier (~1.5 us) is still cheaper than a kernel boundary plus launch (~5-8 us for these sizes).
// Grid barrier: one arrive counter + a generation word (release/acquire at gpu scope; the acquire also drops stale L1 lines).
__device__ __forceinline__ void grid_barrier(unsigned int *state) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int gen;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(state + 1) : "memory");
        __threadfence();
        if (atomicAdd(state, 1u) == gridDim.x - 1) {
            state[0] = 0;
            __threadfence();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(state + 1), "r"(gen + 1) : "memory");
        } else {
            unsigned int now;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(now) : "l"(state + 1) : "memory");
            } while (now == gen);
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(512) coarse_vcycle_grid_kernel(const CoarseLevelDev *L, int nl, int nu, const double *__restrict__ cinv,
                                                                 unsigned int *barrier_state) {
    const int nt = int(gridDim.x) * blockDim.x;
    const int tid = int(blockIdx.x) * blockDim.x + threadIdx.x;
    coarse_vcycle_body([&]() { grid_barrier(barrier_state); }, tid, nt, L, nl, nu, cinv);
}

__global__ void comp_scatter_kernel(const double *__restrict__ e, double *__restrict__ z, int neq, int comp, int64_t n) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) z[int64_t(neq) * i + comp] = e[i];
}

int upload_ssell(fedmx_ctx *ctx, ScalarSell &M, int64_t n_rows, int64_t n_cols, const int32_t *rowptr, const int32_t *col,
                 const double *val, bool try_compact = false) {
    M.n_rows = n_rows;
    M.n_cols = n_cols;
    M.n_slices = (n_rows + FEDMX_SLICE - 1) / FEDMX_SLICE;
    std::vector<int64_t> sp(M.n_slices + 1, 0);
    std::vector<int32_t> sw(M.n_slices, 0);
    for (int64_t s = 0; s < M.n_slices; ++s) {
        int w = 0;
        for (int64_t r = s * FEDMX_SLICE; r < std::min<int64_t>(n_rows, (s + 1) * FEDMX_SLICE); ++r)
            w = std::max(w, rowptr[r + 1] - rowptr[r]);
        sw[s] = w;
        sp[s + 1] = sp[s] + int64_t(w) * FEDMX_SLICE;
    }
    int rc;
    int64_t *dsp; int32_t *dsw, *dc; float *dv;
    if ((rc = fedmx_upload(ctx, &dsp, sp))) return rc;
    if ((rc = fedmx_upload(ctx, &dsw, sw))) return rc;
    M.val16 = nullptr; M.col16 = nullptr; M.rdiag = nullptr;
    if (try_compact && n_rows == n_cols) {
        // 4 bytes per nonzero: possible when every column offset fits 16 bits and every row has a non-zero diagonal
        bool ok = true;
        std::vector<double> diag(n_rows, 0.0);
        for (int64_t r = 0; r < n_rows && ok; ++r)
            for (int32_t q = rowptr[r]; q < rowptr[r + 1]; ++q) {
                const int64_t d = int64_t(col[q]) - r;
                if (d < -32768 || d > 32767) { ok = false; break; }
                if (d == 0) diag[r] = val[q];
            }
        for (int64_t r = 0; r < n_rows && ok; ++r) ok = diag[r] != 0.0;
        if (ok) {
            std::vector<unsigned short> v16(sp[M.n_slices], 0);
            std::vector<short> c16(sp[M.n_slices], 0);
            for (int64_t r = 0; r < n_rows; ++r) {
                const int64_t s = r / FEDMX_SLICE, l = r % FEDMX_SLICE;
                for (int32_t q = rowptr[r]; q < rowptr[r + 1]; ++q) {
                    const int64_t e = sp[s] + int64_t(q - rowptr[r]) * FEDMX_SLICE + l;
                    c16[e] = short(int64_t(col[q]) - r);
                    v16[e] = __half_as_ushort(__float2half_rn(float(val[q] / diag[r])));
                }
            }
            unsigned short *dv16; short *dc16; double *dd;
            if ((rc = fedmx_upload(ctx, &dv16, v16))) return rc;
            if ((rc = fedmx_upload(ctx, &dc16, c16))) return rc;
            if ((rc = fedmx_upload(ctx, &dd, diag))) return rc;
            M.slice_ptr = dsp; M.slice_width = dsw; M.col = nullptr; M.val = nullptr;
            M.val16 = dv16; M.col16 = dc16; M.rdiag = dd;
            return FEDMX_OK;
        }
    }
    std::vector<int32_t> c(sp[M.n_slices], 0);
    std::vector<float> v(sp[M.n_slices], 0.0f);
    for (int64_t r = 0; r < n_rows; ++r) {
        const int64_t s = r / FEDMX_SLICE, l = r % FEDMX_SLICE;
        for (int32_t q = rowptr[r]; q < rowptr[r + 1]; ++q) {
            const int64_t e = sp[s] + int64_t(q - rowptr[r]) * FEDMX_SLICE + l;
            c[e] = col[q];
            v[e] = float(val[q]);
        }
    }
    if ((rc = fedmx_upload(ctx, &dc, c))) return rc;
    if ((rc = fedmx_upload(ctx, &dv, v))) return rc;
    M.slice_ptr = dsp; M.slice_width = dsw; M.col = dc; M.val = dv;
    return FEDMX_OK;
}

template <int MODE>
void launch_ssell(fedmx_ctx *ctx, const ScalarSell &A, const double *x, const double *b, const double *dinv, double omega, double *y,
                  int ys = 1, int yo = 0) {
    const int threads = 128;
    const unsigned blocks = unsigned((A.n_slices * FEDMX_SLICE + threads - 1) / threads);
    if (A.val16) ssell_kernel<MODE, true><<<blocks, threads, 0, ctx->stream>>>(A, x, b, dinv, omega, y, ys, yo);
    else ssell_kernel<MODE><<<blocks, threads, 0, ctx->stream>>>(A, x, b, dinv, omega, y, ys, yo);
    ctx->launches++;
}

}  // namespace

// Levels with at most kFuseRows rows are walked by the single-CTA kernel (they are launch-latency bound otherwise).
static const int64_t kFuseRows = getenv("FEDMX_FUSE_ROWS") ? atoll(getenv("FEDMX_FUSE_ROWS")) : 3000;  // one cluster of 8 CTAs; measured on B200: 3000, 8000 and 80000 rows give 26.8 / 26.9 / 29.9 ms per step (8 SMs cannot feed larger levels)
constexpr int kClusterSize = 8;

// FEDMX_MG_FUSE=cluster (default): only the levels of up to kFuseRows rows are fused (one cluster), the others are separate
// launches.  FEDMX_MG_FUSE=grid: levels of up to kGridFuseRows rows run inside ONE cooperative whole-GPU kernel.  Measured on
// B200 at 10 M DOF: the grid variant needs 182-195 us for the levels below the finest one against 149 us (mid 121 + cluster 27)
// with separate launches -- a persistent grid has 4x fewer threads than rows on the 555 k-row level, so its gathers run with a
// quarter of the memory-level parallelism; the barrier flavour (cooperative_groups or the hand-written one) makes no difference.
static const bool kGridFuse = getenv("FEDMX_MG_FUSE") && std::string(getenv("FEDMX_MG_FUSE")) == "grid";
static const int64_t kGridFuseRows = getenv("FEDMX_GRID_FUSE_ROWS") ? atoll(getenv("FEDMX_GRID_FUSE_ROWS")) : 1000000;

static int setup_coarse_fusion(fedmx_ctx *ctx) {
    PoolScope pool(ctx, &ctx->amg_device);
    auto &L = ctx->amg_levels;
    const int nl = int(L.size());
    int first = nl - 1;
    int coop = 0;
    FEDMX_CHECK_CUDA(ctx, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
    ctx->coarse_grid_blocks = 0;
    if (kGridFuse && coop && nl <= 12) {
        int per_sm = 0, sms = 0;
        FEDMX_CHECK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, coarse_vcycle_grid_kernel, 512, 0));
        FEDMX_CHECK_CUDA(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
        ctx->coarse_grid_blocks = std::min(per_sm, 2) * sms;
        if (!ctx->d_grid_barrier) {
            int rc = fedmx_dev_alloc(ctx, &ctx->d_grid_barrier, 2);
            if (rc) return rc;
            FEDMX_CHECK_CUDA(ctx, cudaMemset(ctx->d_grid_barrier, 0, 2 * sizeof(unsigned int)));
        }
    }
    const int64_t limit = ctx->coarse_grid_blocks > 0 ? kGridFuseRows : kFuseRows;
    while (first > 0 && L[first - 1].n <= limit) --first;
    if (ctx->n_ranks > 1 && first == 0) first = 1;          // the finest level of a partitioned run needs halo exchanges
    if (first > nl - 1) first = nl - 1;
    std::vector<CoarseLevelDev> h(nl - first);
    for (int l = first; l < nl; ++l) {
        CoarseLevelDev &d = h[l - first];
        d.n = L[l].n; d.n_coarse = L[l].n_coarse; d.omega = L[l].omega;
        d.A = L[l].A; d.P = L[l].P; d.R = L[l].R;
        d.dinv = L[l].dinv; d.b = L[l].b; d.x = L[l].x; d.x2 = L[l].x2; d.r = L[l].r;
    }
    CoarseLevelDev *dev;
    int rc = fedmx_dev_alloc(ctx, &dev, h.size());
    if (rc) return rc;
    FEDMX_CHECK_CUDA(ctx, cudaMemcpy(dev, h.data(), h.size() * sizeof(CoarseLevelDev), cudaMemcpyHostToDevice));
    ctx->d_coarse_levels = dev;
    ctx->coarse_first = first;
    ctx->coarse_count = nl - first;
    return FEDMX_OK;
}

// One V-cycle with zero initial guess: x_0 = Vcycle(b_0); b_0 = levels[0].b.  The result goes to component z_off of the
// node-interleaved vector z_out (stride z_stride): written by the last smoothing sweep itself when the finest level runs
// as separate launches, by a small scatter kernel otherwise.
int fedmx_amg_vcycle(fedmx_ctx *ctx, double *z_out, int z_stride, int z_off) {
    auto &L = ctx->amg_levels;
    const int nl = int(L.size());
    FEDMX_REQUIRE(ctx, nl > 0 && ctx->amg_coarse_inv, "AMG hierarchy not set");
    const int nu = ctx->amg_nu;
    int rc;
    if (ctx->coarse_first < 0 && (rc = setup_coarse_fusion(ctx))) return rc;
    const int cf = ctx->coarse_first;
    // V(0,1) on the finest level (FEDMX_MG_V01=1; default for partitioned runs whose finest prolongation covers the ghost rows):
    // no pre-smoothing there, the right-hand side itself is restricted and the prolongated correction is smoothed once.
    // On a partitioned level this removes BOTH halo exchanges of the cycle (x = P e is evaluated for the ghost rows too, the
    // coarse correction e being replicated), at the price of a slightly weaker preconditioner.
    const bool v01 = ctx->amg_v01 && cf > 0 && nu == 1;
    // ---- the levels below the finest are replicated on every rank and need no communication: their ~15 small kernels (6-9 us
    // each, launch-latency bound) are captured once into a CUDA graph and replayed.  The smoothers ping-pong x/x2, so the
    // kernel arguments alternate with period two: one graph per parity.
    static const bool env_graph = !(getenv("FEDMX_MG_GRAPH") && atoi(getenv("FEDMX_MG_GRAPH")) == 0);
    auto down_level = [&](int l) -> int {
        ProfScope lprof(ctx, l == 0 ? FEDMX_PROF_MG_FINE : FEDMX_PROF_MG_MID);
        AmgLevel &lv = L[l];
        if (v01 && l == 0) {
            constexpr int LPR = 4;
            const unsigned rb = unsigned((lv.R.n_rows * LPR + 127) / 128);
            ssell_mult_split_kernel<LPR><<<rb, 128, 0, ctx->stream>>>(lv.R, lv.b, L[1].b);
            ctx->launches++;
            if (ctx->n_ranks > 1 && (rc = fedmx_global_sum(ctx, L[1].b, int(L[1].n)))) return rc;
            return FEDMX_OK;
        }
        const unsigned vb = unsigned((lv.n + 255) / 256);
        const bool fuse_first = nu == 1 && !(l == 0 && ctx->n_ranks > 1);   // distributed level: ghosts of x come by halo exchange
        if (fuse_first) {
            launch_ssell<MODE_FIRST_RESID>(ctx, lv.A, lv.x, lv.b, lv.dinv, lv.omega, lv.r);
        } else {
            first_sweep_kernel<<<vb, 256, 0, ctx->stream>>>(lv.b, lv.dinv, lv.omega, lv.x, lv.n);
            ctx->launches++;
        }
        for (int sweep = 1; sweep < nu; ++sweep) {
            if (l == 0 && ctx->n_ranks > 1 && (rc = fedmx_halo_exchange_internal(ctx, lv.x, 1))) return rc;
            launch_ssell<MODE_JACOBI>(ctx, lv.A, lv.x, lv.b, lv.dinv, lv.omega, lv.x2);
            std::swap(lv.x, lv.x2);
        }
        const bool dist = (l == 0 && ctx->n_ranks > 1);
        if (dist && (rc = fedmx_halo_exchange_internal(ctx, lv.x, 1))) return rc;      // ghosts of x for A x
        if (!fuse_first) launch_ssell<MODE_RESID>(ctx, lv.A, lv.x, lv.b, nullptr, 0.0, lv.r);
        {
            constexpr int LPR = 4;
            const unsigned rb = unsigned((lv.R.n_rows * LPR + 127) / 128);
            ssell_mult_split_kernel<LPR><<<rb, 128, 0, ctx->stream>>>(lv.R, lv.r, L[l + 1].b);
            ctx->launches++;
        }
        // distributed fine level: R holds only this rank's fine columns; the coarse right-hand side is the sum over ranks
        if (dist && (rc = fedmx_global_sum(ctx, L[l + 1].b, int(L[l + 1].n)))) return rc;
        return FEDMX_OK;
    };
    auto coarse_levels = [&]() -> int {
        ProfScope cprof(ctx, FEDMX_PROF_MG_COARSE);
        if (ctx->coarse_grid_blocks > 0) {
            const CoarseLevelDev *levels = (const CoarseLevelDev *)ctx->d_coarse_levels;
            int count = ctx->coarse_count, sweeps = nu;
            const double *cinv = ctx->amg_coarse_inv;
            unsigned int *bstate = ctx->d_grid_barrier;
            void *args[] = {(void *)&levels, (void *)&count, (void *)&sweeps, (void *)&cinv, (void *)&bstate};
            FEDMX_CHECK_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)coarse_vcycle_grid_kernel, dim3(ctx->coarse_grid_blocks),
                                                              dim3(512), args, 0, ctx->stream));
        } else {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(kClusterSize);
            cfg.blockDim = dim3(1024);
            cfg.stream = ctx->stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = kClusterSize;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            FEDMX_CHECK_CUDA(ctx, cudaLaunchKernelEx(&cfg, coarse_vcycle_kernel, (const CoarseLevelDev *)ctx->d_coarse_levels,
                                                     ctx->coarse_count, nu, (const double *)ctx->amg_coarse_inv));
        }
        ctx->launches++;
        return FEDMX_OK;
    };
    auto up_level = [&](int l) -> int {
        ProfScope lprof(ctx, l == 0 ? FEDMX_PROF_MG_FINE : FEDMX_PROF_MG_MID);
        AmgLevel &lv = L[l];
        if (v01 && l == 0) launch_ssell<MODE_MULT>(ctx, lv.P, L[l + 1].x, nullptr, nullptr, 0.0, lv.x);     // x = P e (all local rows)
        else launch_ssell<MODE_ADD>(ctx, lv.P, L[l + 1].x, nullptr, nullptr, 0.0, lv.x);
        for (int sweep = 0; sweep < nu; ++sweep) {
            const bool ghosts_valid = v01 && l == 0 && lv.P.n_rows >= ctx->n_vertices;       // P covers the ghost rows
            if (l == 0 && ctx->n_ranks > 1 && !ghosts_valid && (rc = fedmx_halo_exchange_internal(ctx, lv.x, 1))) return rc;
            if (l == 0 && sweep == nu - 1) {
                launch_ssell<MODE_JACOBI_OUT>(ctx, lv.A, lv.x, lv.b, lv.dinv, lv.omega, z_out, z_stride, z_off);
                break;
            }
            launch_ssell<MODE_JACOBI>(ctx, lv.A, lv.x, lv.b, lv.dinv, lv.omega, lv.x2);
            std::swap(lv.x, lv.x2);
        }
        return FEDMX_OK;
    };
    auto replicated_section = [&]() -> int {               // levels 1 .. cf-1 down, fused coarse levels, levels cf-1 .. 1 up
        for (int l = 1; l < cf; ++l)
            if ((rc = down_level(l))) return rc;
        if ((rc = coarse_levels())) return rc;
        for (int l = cf - 1; l >= 1; --l)
            if ((rc = up_level(l))) return rc;
        return FEDMX_OK;
    };
    if (cf >= 1 && (rc = down_level(0))) return rc;
    const bool use_graph = env_graph && cf >= 2 && nu == 1 && ctx->prof_on < 2 && ctx->coarse_grid_blocks == 0 && ctx->stream != 0 && !ctx->mg_graph_off;
    if (cf == 0) {
        if ((rc = coarse_levels())) return rc;
    } else if (!use_graph) {
        if ((rc = replicated_section())) return rc;
    } else {
        cudaGraphExec_t &exec = ctx->mg_graph[ctx->mg_parity];
        bool launched = false;
        if (!exec) {
            // capture this cycle's section; if the driver refuses (capture unsupported in this context), run it directly and
            // stop trying
            const int64_t before = ctx->launches;
            cudaGraph_t graph = nullptr;
            if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
                cudaGetLastError();
                ctx->mg_graph_off = true;
                if ((rc = replicated_section())) return rc;
                launched = true;
            } else {
                rc = replicated_section();                 // recorded, not executed; performs this cycle's x/x2 swaps
                cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
                if (rc) return rc;
                if (ce == cudaSuccess) ce = cudaGraphInstantiate(&exec, graph, 0);
                if (graph) cudaGraphDestroy(graph);
                FEDMX_CHECK_CUDA(ctx, ce);
                ctx->mg_graph_nodes = int(ctx->launches - before);
            }
        } else {
            for (int l = 1; l < cf; ++l) std::swap(L[l].x, L[l].x2);      // what the section does on the host
            ctx->launches += ctx->mg_graph_nodes;
        }
        if (!launched) FEDMX_CHECK_CUDA(ctx, cudaGraphLaunch(exec, ctx->stream));
        ctx->mg_parity ^= 1;
    }
    if (cf >= 1 && (rc = up_level(0))) return rc;
    if (cf == 0) {                                          // the whole hierarchy ran inside the fused kernel
        const int64_t n = L[0].n;
        comp_scatter_kernel<<<unsigned((n + 255) / 256), 256, 0, ctx->stream>>>(L[0].x, z_out, z_stride, z_off, n);
        ctx->launches++;
    }
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    return FEDMX_OK;
}

int fedmx_apply_dinv(fedmx_ctx *ctx, const double *d_r, double *d_z);

// z = M^-1 r  (r and z may not alias)
int fedmx_precond_apply(fedmx_ctx *ctx, const double *d_r, double *d_z) {
    int rc;
    if (ctx->precond_kind == 0) return fedmx_apply_dinv(ctx, d_r, d_z);
    if (ctx->precond_kind == 2) return fedmx_ilu0_solve_internal(ctx, d_r, d_z);
    ProfScope prof(ctx, FEDMX_PROF_PRECOND);
    AmgLevel &l0 = ctx->amg_levels[0];
    const int comp = ctx->amg_comp;
    const int64_t n = ctx->n_owned;
    const int threads = 128;
    // block-diagonal two-stage preconditioner: z = Dinv r on every node block, then the potential component is replaced
    // by the multigrid solve of its own residual.  (The block-triangular variants with an extra A_{Phi,T} pass give the
    // same Krylov counts on this problem -- the node-local coupling is already inside Dinv -- and cost a partial SpMV.)
    if (ctx->n_eq == 1) {
        FEDMX_CHECK_CUDA(ctx, cudaMemcpyAsync(l0.b, d_r, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        const unsigned nb = unsigned((n + threads - 1) / threads);
        switch (ctx->n_eq) {
            case 2: dinv_extract_kernel<2><<<nb, threads, 0, ctx->stream>>>(ctx->sell, ctx->d_dinv, d_r, comp, d_z, l0.b); break;
            case 3: dinv_extract_kernel<3><<<nb, threads, 0, ctx->stream>>>(ctx->sell, ctx->d_dinv, d_r, comp, d_z, l0.b); break;
            case 4: dinv_extract_kernel<4><<<nb, threads, 0, ctx->stream>>>(ctx->sell, ctx->d_dinv, d_r, comp, d_z, l0.b); break;
            case 5: dinv_extract_kernel<5><<<nb, threads, 0, ctx->stream>>>(ctx->sell, ctx->d_dinv, d_r, comp, d_z, l0.b); break;
            default: ctx->err = "two-stage preconditioner: unsupported n_eq"; return FEDMX_EUNSUPPORTED;
        }
        ctx->launches++;
    }
    return fedmx_amg_vcycle(ctx, d_z, ctx->n_eq, comp);
}

// Krylov fast path of the two-stage preconditioner: z = M^-1 A v with the node-block Jacobi stage and the extraction of the
// multigrid right-hand side fused into the SpMV epilogue, and the multigrid result written by its last smoothing sweep.
int fedmx_two_stage_after_spmv(fedmx_ctx *ctx, const double *d_v, double *d_z) {
    int rc;
    if ((rc = fedmx_spmv_precond(ctx, d_v, d_z, true, ctx->amg_levels[0].b, ctx->amg_comp))) return rc;
    ProfScope prof(ctx, FEDMX_PROF_PRECOND);
    return fedmx_amg_vcycle(ctx, d_z, ctx->n_eq, ctx->amg_comp);
}

extern "C" int fedmx_pc_apply(fedmx_ctx *ctx, const double *d_r, double *d_z) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, d_r && d_z && d_r != d_z, "fedmx_pc_apply: null or aliased vectors");
    FEDMX_REQUIRE(ctx, ctx->d_dinv, "fedmx_pc_apply: no matrix assembled");
    return fedmx_precond_apply(ctx, d_r, d_z);
}

extern "C" int fedmx_amg_reset(fedmx_ctx *ctx) {
    if (!ctx) return FEDMX_EINVAL;
    cudaStreamSynchronize(ctx->stream);
    for (void *p : ctx->amg_device) cudaFree(p);      // a second set-up (new mesh marks, new block) must not leak the first
    ctx->amg_device.clear();
    ctx->amg_levels.clear();
    for (auto &g : ctx->mg_graph) { if (g) cudaGraphExecDestroy(g); g = nullptr; }
    ctx->mg_parity = 0;
    ctx->amg_coarse_inv = nullptr;
    ctx->d_coarse_levels = nullptr;
    ctx->d_grid_barrier = nullptr;
    ctx->precond_kind = 0;
    ctx->coarse_first = -1;
    ctx->coarse_count = 0;
    return FEDMX_OK;
}

extern "C" int fedmx_amg_add_level(fedmx_ctx *ctx, int64_t n, const double *h_dinv, double omega, const int32_t *a_rowptr,
                                   const int32_t *a_col, const double *a_val, int64_t n_coarse, const int32_t *p_rowptr,
                                   const int32_t *p_col, const double *p_val, const int32_t *r_rowptr, const int32_t *r_col,
                                   const double *r_val) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, n > 0 && h_dinv, "fedmx_amg_add_level: bad level");
    FEDMX_REQUIRE(ctx, !ctx->amg_levels.empty() || n == ctx->n_owned, "fedmx_amg_add_level: finest level must have n_owned rows");
    FEDMX_REQUIRE(ctx, ctx->amg_levels.empty() || ctx->amg_levels.back().n_coarse == n, "fedmx_amg_add_level: level sizes do not chain");
    PoolScope pool(ctx, &ctx->amg_device);
    AmgLevel lv{};
    lv.n = n;
    lv.n_coarse = n_coarse;
    lv.omega = omega;
    int rc;
    std::vector<double> d(h_dinv, h_dinv + n);
    if ((rc = fedmx_upload(ctx, &lv.dinv, d))) return rc;
    // the finest level of a partitioned run carries ghost entries (columns of A beyond the owned rows)
    const int64_t nvec = ctx->amg_levels.empty() ? std::max<int64_t>(n, ctx->n_vertices) : n;
    // level 1 of a partitioned run: the restricted residual is summed over the ranks -- produce it directly in the peer
    // heap's allreduce buffer (no staging copy)
    double *heap_b = (ctx->n_ranks > 1 && ctx->amg_levels.size() == 1) ? fedmx_peer_big_buffer(ctx, n) : nullptr;
    if (heap_b) lv.b = heap_b;
    else if ((rc = fedmx_dev_alloc(ctx, &lv.b, nvec))) return rc;
    if ((rc = fedmx_dev_alloc(ctx, &lv.x, nvec))) return rc;
    if ((rc = fedmx_dev_alloc(ctx, &lv.x2, nvec))) return rc;
    if ((rc = fedmx_dev_alloc(ctx, &lv.r, nvec))) return rc;
    FEDMX_CHECK_CUDA(ctx, cudaMemset(lv.b, 0, nvec * sizeof(double)));
    FEDMX_CHECK_CUDA(ctx, cudaMemset(lv.x, 0, nvec * sizeof(double)));
    FEDMX_CHECK_CUDA(ctx, cudaMemset(lv.x2, 0, nvec * sizeof(double)));
    FEDMX_CHECK_CUDA(ctx, cudaMemset(lv.r, 0, nvec * sizeof(double)));
    if (n_coarse > 0) {
        FEDMX_REQUIRE(ctx, a_rowptr && p_rowptr && r_rowptr, "fedmx_amg_add_level: missing operators");
        // finest operator: compact 4-byte entries unless FEDMX_AMG_COMPACT=0 (it is the only multigrid array streamed from HBM)
        static const bool env_compact = !(getenv("FEDMX_AMG_COMPACT") && atoi(getenv("FEDMX_AMG_COMPACT")) == 0);
        const bool finest = ctx->amg_levels.empty();
        if ((rc = upload_ssell(ctx, lv.A, n, finest ? nvec : n, a_rowptr, a_col, a_val, finest && env_compact && nvec == n))) return rc;
        if ((rc = upload_ssell(ctx, lv.P, n, n_coarse, p_rowptr, p_col, p_val))) return rc;
        if ((rc = upload_ssell(ctx, lv.R, n_coarse, n, r_rowptr, r_col, r_val))) return rc;
    }
    ctx->amg_levels.push_back(lv);
    return FEDMX_OK;
}

// Partitioned runs: replace the finest level's prolongation by one that also has the rows of the ghost vertices (the coarse
// correction is replicated, so x = P e can then be formed for the ghosts without a halo exchange) and switch the finest level
// to the V(0,1) cycle.  n_rows = local vertices incl. ghosts.
extern "C" int fedmx_amg_set_fine_prolongation(fedmx_ctx *ctx, int64_t n_rows, const int32_t *p_rowptr, const int32_t *p_col,
                                               const double *p_val, int v01) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, !ctx->amg_levels.empty() && ctx->amg_levels[0].n_coarse > 0, "fedmx_amg_set_fine_prolongation: no hierarchy");
    ctx->amg_v01 = v01 != 0;
    if (!p_rowptr) return FEDMX_OK;
    FEDMX_REQUIRE(ctx, n_rows >= ctx->amg_levels[0].n && n_rows <= ctx->n_vertices, "fedmx_amg_set_fine_prolongation: bad row count");
    PoolScope pool(ctx, &ctx->amg_device);
    return upload_ssell(ctx, ctx->amg_levels[0].P, n_rows, ctx->amg_levels[0].n_coarse, p_rowptr, p_col, p_val);
}

extern "C" int fedmx_amg_set_coarse(fedmx_ctx *ctx, int n, const double *h_inverse) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, !ctx->amg_levels.empty() && ctx->amg_levels.back().n == n && ctx->amg_levels.back().n_coarse == 0,
                  "fedmx_amg_set_coarse: the last level must be the coarsest one");
    FEDMX_REQUIRE(ctx, n <= 8192, "fedmx_amg_set_coarse: coarsest level too large");
    std::vector<double> m(h_inverse, h_inverse + size_t(n) * n);
    PoolScope pool(ctx, &ctx->amg_device);
    return fedmx_upload(ctx, &ctx->amg_coarse_inv, m);
}

extern "C" int fedmx_precond_set(fedmx_ctx *ctx, int kind, int comp, int nu) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, kind == 0 || kind == 1 || kind == 2, "fedmx_precond_set: kind must be 0 (block-Jacobi), 1 (two-stage AMG) or 2 (ILU(0))");
    if (kind == 1) {
        FEDMX_REQUIRE(ctx, !ctx->amg_levels.empty() && ctx->amg_coarse_inv, "fedmx_precond_set: upload the AMG hierarchy first");
        FEDMX_REQUIRE(ctx, comp >= 0 && comp < ctx->n_eq && nu >= 1 && nu <= 4, "fedmx_precond_set: bad component / sweeps");
    }
    ctx->precond_kind = kind;
    ctx->amg_comp = comp;
    ctx->amg_nu = nu;
    ctx->coarse_first = -1;
    for (auto &g : ctx->mg_graph) { if (g) cudaGraphExecDestroy(g); g = nullptr; }
    ctx->mg_parity = 0;
    if (kind == 2 && ctx->d_val) return fedmx_ilu0_factor_internal(ctx);     // factor the matrix that is assembled right now
    return FEDMX_OK;
}

// Entries (i,j) of every stored block, in block-CSR order (host buffer of nnz_blocks doubles).
extern "C" int fedmx_matrix_get_component(fedmx_ctx *ctx, int i, int j, double *h_values) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, ctx->d_val && h_values && i >= 0 && j >= 0 && i < ctx->n_eq && j < ctx->n_eq, "fedmx_matrix_get_component: bad arguments");
    const int NB = ctx->n_eq * ctx->n_eq;
    const int e = i * ctx->n_eq + j;
    const int64_t groups = ctx->sell.n_slots / FEDMX_SLICE;
    std::vector<double> plane(size_t(ctx->sell.n_slots));
    FEDMX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    FEDMX_CHECK_CUDA(ctx, cudaMemcpy2D(plane.data(), FEDMX_SLICE * sizeof(double), ctx->d_val + int64_t(e) * FEDMX_SLICE,
                                       size_t(NB) * FEDMX_SLICE * sizeof(double), FEDMX_SLICE * sizeof(double), groups,
                                       cudaMemcpyDeviceToHost));
    const auto &rp = ctx->h_rowptr;
    for (int64_t r = 0; r < ctx->n_owned; ++r) {
        const int64_t sr = ctx->h_node_row[r];
        const int64_t s = sr / FEDMX_SLICE, l = sr % FEDMX_SLICE;
        for (int32_t q = rp[r]; q < rp[r + 1]; ++q)
            h_values[q] = plane[int64_t(ctx->h_slice_ptr[s]) + int64_t(q - rp[r]) * FEDMX_SLICE + l];
    }
    return FEDMX_OK;
}
