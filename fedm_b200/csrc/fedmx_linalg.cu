// FP64 sparse/dense vector kernels of the Newton-Krylov loop (sm_100a): SELL-32 block SpMV with the
// block-Jacobi preconditioner fused into its epilogue, deterministic two-stage reductions (warp shuffles,
// fixed summation order, no atomics), the fused multi-dot / multi-axpy of classical Gram-Schmidt, the
// BDF2 error estimator, table interpolation and strided copies.
//
// Replaces PETSc MatMult / VecMDot / VecMAXPY / VecNorm / PCApply(bjacobi) under KSPGMRES
// (reference call site: fedm/functions.py:1047 via linear_solver='gmres', fedm_streamer.py:32), numpy's
// np.interp on dof vectors (fedm/functions.py:627-636, 739-747), df.norm (fedm/functions.py:1062-1064) and
// FunctionAssigner.assign (fedm/functions.py:1050).  All of it is HBM-bandwidth-bound.
#include <math.h>

#include "fedmx_internal.h"

namespace {

constexpr int kRedThreads = 256;
constexpr int kMaxRedBlocks = 2368;   // 148 SMs x 16

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum in a fixed order; result valid in thread 0.
__device__ __forceinline__ double block_sum(double v, double *smem) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) smem[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = (l < (blockDim.x >> 5)) ? smem[l] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

// ---- SpMV: y = A x  or  y = Dinv (A x), one thread per block row, SELL-32 entry-major storage ----------
// EXTRACT: additionally rhs_c[row] = (A x)[NEQ*row + comp], the raw right-hand side of the multigrid stage of the
// two-stage preconditioner (saves a separate pass over the residual and the block inverses).
// 32 registers (launch bound of 16 CTAs/SM): full occupancy; 39 registers at the default bound cost 3.4 % (0.334 -> 0.323 ms)
template <int NEQ, bool PRECOND, bool EXTRACT>
__global__ void __launch_bounds__(128, 16) spmv_kernel(SellView S, const double *__restrict__ val,
                                                   const double *__restrict__ dinv, const double *__restrict__ x,
                                                   double *__restrict__ y, int comp, double *__restrict__ rhs_c,
                                                   unsigned zmask) {
    constexpr int NB = NEQ * NEQ;
    const int64_t row = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t s = row / FEDMX_SLICE;
    const int lane = int(row % FEDMX_SLICE);
    if (s >= S.n_slices) return;
    const int width = S.slice_width[s];
    const int64_t gbase = S.slice_ptr[s];
    double acc[NEQ];
#pragma unroll
    for (int k = 0; k < NEQ; ++k) acc[k] = 0.0;
    for (int k = 0; k < width; ++k) {
        const int64_t slot = gbase + int64_t(k) * FEDMX_SLICE;
        const int32_t c = S.col[slot + lane];
        const double *v = val + slot * NB + lane;
        double xc[NEQ];
#pragma unroll
        for (int l = 0; l < NEQ; ++l) xc[l] = __ldg(x + int64_t(NEQ) * c + l);
#pragma unroll
        for (int i = 0; i < NEQ; ++i)
#pragma unroll
            for (int l = 0; l < NEQ; ++l)
                if (!((zmask >> (i * NEQ + l)) & 1u)) acc[i] = fma(__ldcs(v + (i * NEQ + l) * FEDMX_SLICE), xc[l], acc[i]);
    }
    if (row >= S.n_rows) return;
    const int64_t node = S.row_node[row];                // vectors live in node order, the matrix in tiled storage order
    if (EXTRACT) {
#pragma unroll
        for (int i = 0; i < NEQ; ++i)
            if (i == comp) rhs_c[node] = acc[i];
    }
    if (PRECOND) {
        const double *d = dinv + s * FEDMX_SLICE * NB + lane;
        double z[NEQ];
#pragma unroll
        for (int i = 0; i < NEQ; ++i) {
            z[i] = 0.0;
#pragma unroll
            for (int l = 0; l < NEQ; ++l) z[i] = fma(d[(i * NEQ + l) * FEDMX_SLICE], acc[l], z[i]);
        }
#pragma unroll
        for (int i = 0; i < NEQ; ++i) y[int64_t(NEQ) * node + i] = z[i];
    } else {
#pragma unroll
        for (int i = 0; i < NEQ; ++i) y[int64_t(NEQ) * node + i] = acc[i];
    }
}

// ---- block-Jacobi: invert the diagonal blocks (Gauss-Jordan with partial pivoting in registers) ----------
template <int NEQ>
__global__ void block_jacobi_kernel(SellView S, const double *__restrict__ val, double *__restrict__ dinv) {
    constexpr int NB = NEQ * NEQ;
    const int64_t row = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t s = row / FEDMX_SLICE;
    const int lane = int(row % FEDMX_SLICE);
    if (s >= S.n_slices || row >= S.n_rows) return;
    const int diag_slot = int((S.row_pack[row] >> 8) & 0xffu);
    const double *v = val + (int64_t(S.slice_ptr[s]) + int64_t(diag_slot) * FEDMX_SLICE) * NB + lane;
    double a[NEQ][NEQ], b[NEQ][NEQ];
#pragma unroll
    for (int i = 0; i < NEQ; ++i)
#pragma unroll
        for (int j = 0; j < NEQ; ++j) {
            a[i][j] = v[(i * NEQ + j) * FEDMX_SLICE];
            b[i][j] = (i == j) ? 1.0 : 0.0;
        }
#pragma unroll
    for (int c = 0; c < NEQ; ++c) {
        // partial pivoting without dynamic register indexing: swap rows pairwise
#pragma unroll
        for (int r = c + 1; r < NEQ; ++r) {
            if (fabs(a[r][c]) > fabs(a[c][c])) {
#pragma unroll
                for (int j = 0; j < NEQ; ++j) {
                    double t = a[c][j]; a[c][j] = a[r][j]; a[r][j] = t;
                    t = b[c][j]; b[c][j] = b[r][j]; b[r][j] = t;
                }
            }
        }
        const double ip = 1.0 / a[c][c];
#pragma unroll
        for (int j = 0; j < NEQ; ++j) { a[c][j] *= ip; b[c][j] *= ip; }
#pragma unroll
        for (int r = 0; r < NEQ; ++r) {
            if (r == c) continue;
            const double f = a[r][c];
#pragma unroll
            for (int j = 0; j < NEQ; ++j) { a[r][j] -= f * a[c][j]; b[r][j] -= f * b[c][j]; }
        }
    }
    double *d = dinv + s * FEDMX_SLICE * NB + lane;
#pragma unroll
    for (int i = 0; i < NEQ; ++i)
#pragma unroll
        for (int j = 0; j < NEQ; ++j) d[(i * NEQ + j) * FEDMX_SLICE] = b[i][j];
}

// z = Dinv r  (PCApply on a vector)
template <int NEQ>
__global__ void apply_dinv_kernel(SellView S, const double *__restrict__ dinv, const double *__restrict__ r,
                                  double *__restrict__ z) {
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
    }
}

// ---- reductions --------------------------------------------------------------------------------------------
// Fused multi-dot of classical Gram-Schmidt: partial[(j0+j)*gridDim.x + blockIdx.x] = chunk of V_{j0+j} . w for j < J
// (J compile-time, so that the J+1 load streams of an iteration are independent and stay in flight together), and,
// if WW, partial[ww_slot*gridDim.x + blockIdx.x] = chunk of w . w.  Four elements per thread per iteration.
template <int J, bool WW>
__global__ void __launch_bounds__(kRedThreads) multidot_kernel(const double *__restrict__ V, int64_t ld, int j0, int ww_slot,
                                                               const double *__restrict__ w, int64_t n,
                                                               double *__restrict__ partial) {
    __shared__ double smem[32];
    double acc[J > 0 ? J : 1], ww = 0.0;
#pragma unroll
    for (int j = 0; j < J; ++j) acc[j] = 0.0;
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        double wv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) wv[u] = __ldcs(w + i + u * stride);
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const double *vj = V + int64_t(j0 + j) * ld + i;
            double vv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) vv[u] = __ldcs(vj + u * stride);
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[j] = fma(vv[u], wv[u], acc[j]);
        }
        if (WW) {
#pragma unroll
            for (int u = 0; u < 4; ++u) ww = fma(wv[u], wv[u], ww);
        }
    }
    for (; i < n; i += stride) {
        const double wi = w[i];
#pragma unroll
        for (int j = 0; j < J; ++j) acc[j] = fma(V[int64_t(j0 + j) * ld + i], wi, acc[j]);
        if (WW) ww = fma(wi, wi, ww);
    }
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const double r = block_sum(acc[j], smem);
        if (threadIdx.x == 0) partial[int64_t(j0 + j) * gridDim.x + blockIdx.x] = r;
    }
    if (WW) {
        const double r = block_sum(ww, smem);
        if (threadIdx.x == 0) partial[int64_t(ww_slot) * gridDim.x + blockIdx.x] = r;
    }
}

// out[j] = sum_b partial[j*nblocks + b]  in fixed order; one block per j
__global__ void __launch_bounds__(kRedThreads) finish_reduce_kernel(const double *__restrict__ partial, int nblocks,
                                                                    double *__restrict__ out) {
    __shared__ double smem[32];
    const int j = blockIdx.x;
    double acc = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) acc += partial[int64_t(j) * nblocks + b];
    const double r = block_sum(acc, smem);
    if (threadIdx.x == 0) out[j] = r;
}

// w -= sum_{j<J} h[j0+j] V_{j0+j} ; if NORM: partial[blockIdx] = chunk of ||w_new||^2 ;
// if OUT: the result is not written back to w but, scaled, to `out` (the next Krylov basis vector, whose norm the host
// already knows from ||w||^2 - sum h^2) -- saves the separate normalisation pass and the norm reduction.
template <int J, bool NORM, bool OUT>
__global__ void __launch_bounds__(kRedThreads) multiaxpy_kernel(const double *__restrict__ V, int64_t ld, int j0,
                                                                const double *__restrict__ h, double *__restrict__ w,
                                                                int64_t n, double *__restrict__ partial,
                                                                double *__restrict__ out, double scale) {
    __shared__ double smem[32];
    double hs[J];
#pragma unroll
    for (int j = 0; j < J; ++j) hs[j] = h[j0 + j];
    double ww = 0.0;
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    for (; i + stride < n; i += 2 * stride) {
        double w0 = w[i], w1 = w[i + stride];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const double *vj = V + int64_t(j0 + j) * ld + i;
            w0 = fma(-hs[j], __ldcs(vj), w0);
            w1 = fma(-hs[j], __ldcs(vj + stride), w1);
        }
        if (OUT) {
            out[i] = scale * w0;
            out[i + stride] = scale * w1;
        } else {
            w[i] = w0;
            w[i + stride] = w1;
        }
        if (NORM) ww = fma(w0, w0, fma(w1, w1, ww));
    }
    for (; i < n; i += stride) {
        double wi = w[i];
#pragma unroll
        for (int j = 0; j < J; ++j) wi = fma(-hs[j], V[int64_t(j0 + j) * ld + i], wi);
        if (OUT) out[i] = scale * wi;
        else w[i] = wi;
        if (NORM) ww = fma(wi, wi, ww);
    }
    if (NORM) {
        const double r = block_sum(ww, smem);
        if (threadIdx.x == 0) partial[blockIdx.x] = r;
    }
}

// y = alpha * x  (alpha read from device memory as 1/ *inv_of if inv_of != null)
__global__ void scale_copy_kernel(const double *__restrict__ x, double *__restrict__ y, int64_t n, double alpha,
                                  const double *__restrict__ inv_of) {
    const double a = inv_of ? alpha / *inv_of : alpha;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
        y[i] = a * x[i];
}

// x += sum_j c[j] V_j
__global__ void combine_kernel(const double *__restrict__ V, int64_t ld, int nvec, const double *__restrict__ c,
                               double *__restrict__ x, int64_t n, double sign) {
    __shared__ double cs[64];
    if (threadIdx.x < nvec) cs[threadIdx.x] = c[threadIdx.x];
    __syncthreads();
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        double xi = x[i];
        for (int j = 0; j < nvec; ++j) xi = fma(sign * cs[j], V[int64_t(j) * ld + i], xi);
        x[i] = xi;
    }
}

__global__ void axpy_kernel(double alpha, const double *__restrict__ x, double *__restrict__ y, int64_t n) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
        y[i] = fma(alpha, x[i], y[i]);
}

// r = b - r   (r holds A x on entry)
__global__ void residual_kernel(const double *__restrict__ b, double *__restrict__ r, int64_t n) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
        r[i] = b[i] - r[i];
}

// error estimator partial sums: [0] = sum (a-b+eps)^2, [1] = sum (b+eps)^2 over component comp
__global__ void __launch_bounds__(kRedThreads) error_norm_kernel(const double *__restrict__ a, const double *__restrict__ b,
                                                                 int64_t n_nodes, int n_eq, int comp, double eps,
                                                                 double *__restrict__ partial) {
    __shared__ double smem[32];
    double s0 = 0.0, s1 = 0.0;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_nodes; i += int64_t(gridDim.x) * blockDim.x) {
        const double bv = b[i * n_eq + comp];
        const double d = a[i * n_eq + comp] - bv + eps;
        const double e = bv + eps;
        s0 = fma(d, d, s0);
        s1 = fma(e, e, s1);
    }
    const double r0 = block_sum(s0, smem);
    const double r1 = block_sum(s1, smem);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = r0;
        partial[gridDim.x + blockIdx.x] = r1;
    }
}

// np.interp: clamped piecewise-linear look-up, binary search per node
__global__ void interp_kernel(const double *__restrict__ x, int64_t x_stride, int64_t n, const double *__restrict__ kx,
                              const double *__restrict__ ky, int nt, double divisor, double *__restrict__ out,
                              int64_t out_stride) {
    extern __shared__ double tab[];
    double *sx = tab, *sy = tab + nt;
    for (int i = threadIdx.x; i < nt; i += blockDim.x) { sx[i] = kx[i]; sy[i] = ky[i]; }
    __syncthreads();
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const double xv = x[i * x_stride];
        double r;
        if (xv != xv) r = xv;                      // NaN propagates, as in numpy
        else if (xv <= sx[0]) r = sy[0];
        else if (xv >= sx[nt - 1]) r = sy[nt - 1];
        else {
            int lo = 0, hi = nt - 1;               // invariant: sx[lo] <= xv < sx[hi]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (sx[mid] <= xv) lo = mid; else hi = mid;
            }
            const double slope = (sy[lo + 1] - sy[lo]) / (sx[lo + 1] - sx[lo]);
            // numpy's formula (compiled arr_interp); explicit rn ops so that nvcc does not contract it into an FMA
            r = __dadd_rn(__dmul_rn(slope, xv - sx[lo]), sy[lo]);
        }
        out[i * out_stride] = r / divisor;
    }
}

__global__ void strided_copy_kernel(const double *__restrict__ src, int64_t ss, int64_t so, double *__restrict__ dst,
                                    int64_t ds, int64_t dof, int64_t n) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
        dst[i * ds + dof] = src[i * ss + so];
}

// dst = src * mul / div (strided); the reference's ESR rule D = kB*Tgas*mu/e (fedm/functions.py:633)
__global__ void strided_scale_kernel(const double *__restrict__ src, int64_t ss, int64_t so, double *__restrict__ dst,
                                     int64_t ds, int64_t dof, int64_t n, double mul, double div) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
        dst[i * ds + dof] = __ddiv_rn(__dmul_rn(mul, src[i * ss + so]), div);
}

// dst = exp(x[.,ca] - x[.,cb]) : mean energy from ln w_e - ln n_e (fedm_gd.py:452)
__global__ void exp_diff_kernel(const double *__restrict__ x, int n_eq, int ca, int cb, double *__restrict__ dst,
                                int64_t ds, int64_t dof, int64_t n) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
        dst[i * ds + dof] = exp(x[i * n_eq + ca] - x[i * n_eq + cb]);
}

inline unsigned grid_for(int64_t n, int threads, int cap = kMaxRedBlocks) {
    int64_t b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > cap) b = cap;
    return unsigned(b);
}

}  // namespace

// ---- internal helpers used by the Krylov / Newton drivers ---------------------------------------------------
int fedmx_ensure_workspace(fedmx_ctx *ctx, int restart) {
    const int64_t n = int64_t(ctx->n_eq) * ctx->n_vertices;
    int rc;
    if (!ctx->d_res) {
        if ((rc = fedmx_dev_alloc(ctx, &ctx->d_res, n))) return rc;
        if ((rc = fedmx_dev_alloc(ctx, &ctx->d_dx, n))) return rc;
        if ((rc = fedmx_dev_alloc(ctx, &ctx->d_w, n))) return rc;
        if ((rc = fedmx_dev_alloc(ctx, &ctx->d_tmp, n))) return rc;
        if ((rc = fedmx_dev_alloc(ctx, &ctx->d_red, size_t(kMaxRedBlocks) * 72))) return rc;
        if ((rc = fedmx_dev_alloc(ctx, &ctx->d_hess, 4096))) return rc;
        FEDMX_CHECK_CUDA(ctx, cudaMemset(ctx->d_res, 0, n * sizeof(double)));
        FEDMX_CHECK_CUDA(ctx, cudaMemset(ctx->d_dx, 0, n * sizeof(double)));
        FEDMX_CHECK_CUDA(ctx, cudaMemset(ctx->d_w, 0, n * sizeof(double)));
        FEDMX_CHECK_CUDA(ctx, cudaMemset(ctx->d_tmp, 0, n * sizeof(double)));
        FEDMX_CHECK_CUDA(ctx, cudaMallocHost((void **)&ctx->h_pinned, 4096 * sizeof(double)));
    }
    if (restart > ctx->krylov_restart_alloc) {
        FEDMX_REQUIRE(ctx, restart <= 60, "GMRES restart > 60 unsupported");
        // basis vectors are never freed individually; a larger restart allocates a new block
        if ((rc = fedmx_dev_alloc(ctx, &ctx->d_V, size_t(restart + 1) * n))) return rc;
        FEDMX_CHECK_CUDA(ctx, cudaMemset(ctx->d_V, 0, size_t(restart + 1) * n * sizeof(double)));
        ctx->krylov_restart_alloc = restart;
    }
    return FEDMX_OK;
}

int fedmx_update_block_jacobi(fedmx_ctx *ctx) {
    const int threads = 128;
    const unsigned blocks = unsigned((ctx->sell.n_slices * FEDMX_SLICE + threads - 1) / threads);
    switch (ctx->n_eq) {
        case 1: block_jacobi_kernel<1><<<blocks, threads, 0, ctx->stream>>>(ctx->sell, ctx->d_val, ctx->d_dinv); break;
        case 2: block_jacobi_kernel<2><<<blocks, threads, 0, ctx->stream>>>(ctx->sell, ctx->d_val, ctx->d_dinv); break;
        case 3: block_jacobi_kernel<3><<<blocks, threads, 0, ctx->stream>>>(ctx->sell, ctx->d_val, ctx->d_dinv); break;
        case 4: block_jacobi_kernel<4><<<blocks, threads, 0, ctx->stream>>>(ctx->sell, ctx->d_val, ctx->d_dinv); break;
        case 5: block_jacobi_kernel<5><<<blocks, threads, 0, ctx->stream>>>(ctx->sell, ctx->d_val, ctx->d_dinv); break;
        default: ctx->err = "block Jacobi: unsupported n_eq"; return FEDMX_EUNSUPPORTED;
    }
    ctx->launches++;
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    if (ctx->precond_kind == 2) return fedmx_ilu0_factor_internal(ctx);      // new matrix values: new ILU(0) factors
    return FEDMX_OK;
}

int fedmx_spmv_precond(fedmx_ctx *ctx, const double *d_x, double *d_y, bool precond, double *d_rhs_c, int comp) {
    const int threads = 128;
    const unsigned blocks = unsigned((ctx->sell.n_slices * FEDMX_SLICE + threads - 1) / threads);
    ProfScope prof(ctx, FEDMX_PROF_SPMV);
    FEDMX_REQUIRE(ctx, !d_rhs_c || precond, "spmv: component extraction is only built for the preconditioned variant");
#define SPMV_CASE(N)                                                                                              \
    case N:                                                                                                       \
        if (d_rhs_c) spmv_kernel<N, true, true><<<blocks, threads, 0, ctx->stream>>>(ctx->sell, ctx->d_val, ctx->d_dinv, d_x, d_y, comp, d_rhs_c, ctx->zero_mask); \
        else if (precond) spmv_kernel<N, true, false><<<blocks, threads, 0, ctx->stream>>>(ctx->sell, ctx->d_val, ctx->d_dinv, d_x, d_y, 0, nullptr, ctx->zero_mask); \
        else spmv_kernel<N, false, false><<<blocks, threads, 0, ctx->stream>>>(ctx->sell, ctx->d_val, ctx->d_dinv, d_x, d_y, 0, nullptr, ctx->zero_mask);        \
        break;
    switch (ctx->n_eq) {
        SPMV_CASE(1)
        SPMV_CASE(2)
        SPMV_CASE(3)
        SPMV_CASE(4)
        SPMV_CASE(5)
        default: ctx->err = "spmv: unsupported n_eq"; return FEDMX_EUNSUPPORTED;
    }
#undef SPMV_CASE
    ctx->launches++;
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    return FEDMX_OK;
}

int fedmx_apply_dinv(fedmx_ctx *ctx, const double *d_r, double *d_z) {
    const int threads = 128;
    const unsigned blocks = unsigned((ctx->n_owned + threads - 1) / threads);
    switch (ctx->n_eq) {
        case 1: apply_dinv_kernel<1><<<blocks, threads, 0, ctx->stream>>>(ctx->sell, ctx->d_dinv, d_r, d_z); break;
        case 2: apply_dinv_kernel<2><<<blocks, threads, 0, ctx->stream>>>(ctx->sell, ctx->d_dinv, d_r, d_z); break;
        case 3: apply_dinv_kernel<3><<<blocks, threads, 0, ctx->stream>>>(ctx->sell, ctx->d_dinv, d_r, d_z); break;
        case 4: apply_dinv_kernel<4><<<blocks, threads, 0, ctx->stream>>>(ctx->sell, ctx->d_dinv, d_r, d_z); break;
        case 5: apply_dinv_kernel<5><<<blocks, threads, 0, ctx->stream>>>(ctx->sell, ctx->d_dinv, d_r, d_z); break;
        default: ctx->err = "apply_dinv: unsupported n_eq"; return FEDMX_EUNSUPPORTED;
    }
    ctx->launches++;
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    return FEDMX_OK;
}

template <bool WW>
static void launch_multidot(fedmx_ctx *ctx, int J, unsigned blocks, const double *V, int64_t ld, int j0, int ww_slot,
                            const double *w, int64_t n) {
#define MD_CASE(JJ) case JJ: multidot_kernel<JJ, WW><<<blocks, kRedThreads, 0, ctx->stream>>>(V, ld, j0, ww_slot, w, n, ctx->d_red); break;
    switch (J) {
        MD_CASE(0) MD_CASE(1) MD_CASE(2) MD_CASE(3) MD_CASE(4) MD_CASE(5) MD_CASE(6) MD_CASE(7) MD_CASE(8)
    }
#undef MD_CASE
    ctx->launches++;
}

// out_dev[0..nvec) = V_j . w (owned entries), out_dev[nvec] = w.w ; result stays on the device.
int fedmx_multidot(fedmx_ctx *ctx, const double *V, int64_t ld, int nvec, const double *w, int64_t n, double *out_dev,
                   bool with_ww) {
    const unsigned blocks = grid_for(n, kRedThreads * 4);
    int j0 = 0;
    bool ww_done = !with_ww;
    do {
        const int J = nvec - j0 > 8 ? 8 : nvec - j0;
        if (!ww_done) launch_multidot<true>(ctx, J, blocks, V, ld, j0, nvec, w, n);
        else launch_multidot<false>(ctx, J, blocks, V, ld, j0, nvec, w, n);
        ww_done = true;
        j0 += J;
    } while (j0 < nvec);
    finish_reduce_kernel<<<nvec + (with_ww ? 1 : 0), kRedThreads, 0, ctx->stream>>>(ctx->d_red, int(blocks), out_dev);
    ctx->launches++;
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    return FEDMX_OK;
}

template <bool NORM, bool OUT>
static void launch_multiaxpy(fedmx_ctx *ctx, int J, unsigned blocks, const double *V, int64_t ld, int j0, const double *h,
                             double *w, int64_t n, double *out = nullptr, double scale = 1.0) {
#define MA_CASE(JJ) case JJ: multiaxpy_kernel<JJ, NORM, OUT><<<blocks, kRedThreads, 0, ctx->stream>>>(V, ld, j0, h, w, n, ctx->d_red, out, scale); break;
    switch (J) {
        MA_CASE(1) MA_CASE(2) MA_CASE(3) MA_CASE(4) MA_CASE(5) MA_CASE(6) MA_CASE(7) MA_CASE(8)
    }
#undef MA_CASE
    ctx->launches++;
}

int fedmx_multiaxpy_norm(fedmx_ctx *ctx, const double *V, int64_t ld, int nvec, const double *h_dev, double *w, int64_t n,
                         double *norm2_dev) {
    const unsigned blocks = grid_for(n, kRedThreads * 2);
    int j0 = 0;
    while (j0 < nvec) {
        const int J = nvec - j0 > 8 ? 8 : nvec - j0;
        const bool last = j0 + J >= nvec;
        if (last) launch_multiaxpy<true, false>(ctx, J, blocks, V, ld, j0, h_dev, w, n);
        else launch_multiaxpy<false, false>(ctx, J, blocks, V, ld, j0, h_dev, w, n);
        j0 += J;
    }
    finish_reduce_kernel<<<1, kRedThreads, 0, ctx->stream>>>(ctx->d_red, int(blocks), norm2_dev);
    ctx->launches++;
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    return FEDMX_OK;
}

// out = scale * (w - sum_j h_j V_j): Gram-Schmidt update and normalisation in one pass (w is left untouched when
// nvec <= 8, otherwise it holds the partially updated vector)
int fedmx_multiaxpy_scaled(fedmx_ctx *ctx, const double *V, int64_t ld, int nvec, const double *h_dev, double *w, int64_t n,
                           double *out, double scale) {
    const unsigned blocks = grid_for(n, kRedThreads * 2);
    int j0 = 0;
    while (j0 < nvec) {
        const int J = nvec - j0 > 8 ? 8 : nvec - j0;
        const bool last = j0 + J >= nvec;
        if (last) launch_multiaxpy<false, true>(ctx, J, blocks, V, ld, j0, h_dev, w, n, out, scale);
        else launch_multiaxpy<false, false>(ctx, J, blocks, V, ld, j0, h_dev, w, n);
        j0 += J;
    }
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    return FEDMX_OK;
}

int fedmx_scale_copy(fedmx_ctx *ctx, const double *x, double *y, int64_t n, double alpha, const double *inv_of_dev) {
    scale_copy_kernel<<<grid_for(n, 256, 4736), 256, 0, ctx->stream>>>(x, y, n, alpha, inv_of_dev);
    ctx->launches++;
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    return FEDMX_OK;
}

int fedmx_combine(fedmx_ctx *ctx, const double *V, int64_t ld, int nvec, const double *c_dev, double *x, int64_t n, double sign) {
    combine_kernel<<<grid_for(n, 256, 4736), 256, 0, ctx->stream>>>(V, ld, nvec, c_dev, x, n, sign);
    ctx->launches++;
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    return FEDMX_OK;
}

int fedmx_axpy(fedmx_ctx *ctx, double alpha, const double *x, double *y, int64_t n) {
    axpy_kernel<<<grid_for(n, 256, 4736), 256, 0, ctx->stream>>>(alpha, x, y, n);
    ctx->launches++;
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    return FEDMX_OK;
}

int fedmx_residual_from_ax(fedmx_ctx *ctx, const double *b, double *r, int64_t n) {
    residual_kernel<<<grid_for(n, 256, 4736), 256, 0, ctx->stream>>>(b, r, n);
    ctx->launches++;
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    return FEDMX_OK;
}

int fedmx_global_sum(fedmx_ctx *ctx, double *d_buf, int n);   // fedmx_comm.cu

// sum of squares of the first n entries; global = allreduce over ranks. Synchronises the stream.
int fedmx_norm2_local(fedmx_ctx *ctx, const double *d_x, int64_t n, double *out, bool global) {
    int rc = fedmx_multidot(ctx, nullptr, 0, 0, d_x, n, ctx->d_hess + 2048, true);
    if (rc) return rc;
    if (global && ctx->n_ranks > 1 && (rc = fedmx_global_sum(ctx, ctx->d_hess + 2048, 1))) return rc;
    FEDMX_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 2048, ctx->d_hess + 2048, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FEDMX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    FEDMX_REQUIRE(ctx, !(ctx->peer.h_err && *ctx->peer.h_err), "peer-memory collective timed out waiting for a neighbour's flag");
    *out = sqrt(ctx->h_pinned[2048]);
    return FEDMX_OK;
}

// ---- C ABI --------------------------------------------------------------------------------------------------
extern "C" int fedmx_spmv(fedmx_ctx *ctx, const double *d_x, double *d_y) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, ctx->d_val && d_x && d_y, "fedmx_spmv: no matrix / null vectors");
    return fedmx_spmv_precond(ctx, d_x, d_y, false);
}

extern "C" int fedmx_norm2(fedmx_ctx *ctx, const double *d_x, int64_t n, double *out) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, d_x && out && n >= 0, "fedmx_norm2: bad arguments");
    int rc;
    if (!ctx->d_red && (rc = fedmx_ensure_workspace(ctx, 0))) return rc;       // contexts without a triangle mesh (1-D P2 path)
    return fedmx_norm2_local(ctx, d_x, n, out, true);
}

extern "C" int fedmx_dot(fedmx_ctx *ctx, const double *d_x, const double *d_y, int64_t n, double *out) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, ctx->d_red && d_x && d_y && out && n >= 0, "fedmx_dot: bad arguments");
    int rc = fedmx_multidot(ctx, d_x, n, 1, d_y, n, ctx->d_hess + 2048, false);
    if (rc) return rc;
    if (ctx->n_ranks > 1 && (rc = fedmx_global_sum(ctx, ctx->d_hess + 2048, 1))) return rc;
    FEDMX_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 2048, ctx->d_hess + 2048, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FEDMX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = ctx->h_pinned[2048];
    return FEDMX_OK;
}

extern "C" int fedmx_error_norm(fedmx_ctx *ctx, const double *d_new, const double *d_old, int comp, double eps, double *err) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, ctx->d_red && d_new && d_old && err, "fedmx_error_norm: bad arguments");
    FEDMX_REQUIRE(ctx, comp >= 0 && comp < ctx->n_eq, "fedmx_error_norm: bad component");
    const unsigned blocks = grid_for(ctx->n_owned, kRedThreads);
    error_norm_kernel<<<blocks, kRedThreads, 0, ctx->stream>>>(d_new, d_old, ctx->n_owned, ctx->n_eq, comp, eps, ctx->d_red);
    finish_reduce_kernel<<<2, kRedThreads, 0, ctx->stream>>>(ctx->d_red, int(blocks), ctx->d_hess + 2048);
    ctx->launches += 2;
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    int rc;
    if (ctx->n_ranks > 1 && (rc = fedmx_global_sum(ctx, ctx->d_hess + 2048, 2))) return rc;
    FEDMX_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 2048, ctx->d_hess + 2048, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FEDMX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *err = sqrt(ctx->h_pinned[2048]) / sqrt(ctx->h_pinned[2049]);
    return FEDMX_OK;
}

extern "C" int fedmx_error_norm_strided(fedmx_ctx *ctx, const double *d_new, const double *d_old, int64_t stride, int64_t offset,
                                        int64_t n, double eps, double *err) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, ctx->d_red && d_new && d_old && err && stride >= 1 && offset >= 0 && offset < stride && n >= 0,
                  "fedmx_error_norm_strided: bad arguments");
    const unsigned blocks = grid_for(n, kRedThreads);
    error_norm_kernel<<<blocks, kRedThreads, 0, ctx->stream>>>(d_new, d_old, n, int(stride), int(offset), eps, ctx->d_red);
    finish_reduce_kernel<<<2, kRedThreads, 0, ctx->stream>>>(ctx->d_red, int(blocks), ctx->d_hess + 2048);
    ctx->launches += 2;
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    int rc;
    if (ctx->n_ranks > 1 && (rc = fedmx_global_sum(ctx, ctx->d_hess + 2048, 2))) return rc;
    FEDMX_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 2048, ctx->d_hess + 2048, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FEDMX_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *err = sqrt(ctx->h_pinned[2048]) / sqrt(ctx->h_pinned[2049]);
    return FEDMX_OK;
}

extern "C" int fedmx_interp_table(fedmx_ctx *ctx, const double *d_x, int64_t x_stride, int64_t n, const double *d_kx,
                                  const double *d_ky, int n_table, double divisor, double *d_out, int64_t out_stride) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, d_x && d_kx && d_ky && d_out && n >= 0, "fedmx_interp_table: null argument");
    FEDMX_REQUIRE(ctx, n_table >= 1 && n_table <= 2048, "fedmx_interp_table: table size must be in 1..2048");
    if (n == 0) return FEDMX_OK;
    interp_kernel<<<grid_for(n, 256, 4736), 256, 2 * n_table * sizeof(double), ctx->stream>>>(
        d_x, x_stride, n, d_kx, d_ky, n_table, divisor, d_out, out_stride);
    ctx->launches++;
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    return FEDMX_OK;
}

extern "C" int fedmx_strided_copy(fedmx_ctx *ctx, const double *d_src, int64_t src_stride, int64_t src_off, double *d_dst,
                                  int64_t dst_stride, int64_t dst_off, int64_t n) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, d_src && d_dst && n >= 0, "fedmx_strided_copy: null argument");
    if (n == 0) return FEDMX_OK;
    strided_copy_kernel<<<grid_for(n, 256, 4736), 256, 0, ctx->stream>>>(d_src, src_stride, src_off, d_dst, dst_stride, dst_off, n);
    ctx->launches++;
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    return FEDMX_OK;
}

extern "C" int fedmx_strided_scale(fedmx_ctx *ctx, const double *d_src, int64_t src_stride, int64_t src_off, double *d_dst,
                                   int64_t dst_stride, int64_t dst_off, int64_t n, double mul, double div) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, d_src && d_dst && n >= 0, "fedmx_strided_scale: null argument");
    if (n == 0) return FEDMX_OK;
    strided_scale_kernel<<<grid_for(n, 256, 4736), 256, 0, ctx->stream>>>(d_src, src_stride, src_off, d_dst, dst_stride, dst_off, n, mul, div);
    ctx->launches++;
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    return FEDMX_OK;
}

extern "C" int fedmx_exp_diff(fedmx_ctx *ctx, const double *d_x, int n_eq, int comp_a, int comp_b, double *d_dst,
                              int64_t dst_stride, int64_t dst_off, int64_t n) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, d_x && d_dst && n >= 0 && comp_a >= 0 && comp_b >= 0 && comp_a < n_eq && comp_b < n_eq, "fedmx_exp_diff: bad arguments");
    if (n == 0) return FEDMX_OK;
    exp_diff_kernel<<<grid_for(n, 256, 4736), 256, 0, ctx->stream>>>(d_x, n_eq, comp_a, comp_b, d_dst, dst_stride, dst_off, n);
    ctx->launches++;
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    return FEDMX_OK;
}
