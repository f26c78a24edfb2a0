// ILU(0) preconditioner, per-GPU block of block Jacobi (sm_100a).
//
// Replaces PETSc's default preconditioner of the reference TEST's KSP: `linear_solver = "gmres"` with no PC given
// (tests/integrated_tests/streamer_discharge/fedm_streamer.py:32, 302-305) = ILU(0) in serial and block Jacobi with ILU(0) on
// every rank's diagonal block under mpirun.  Here: one block per GPU (ghost columns are dropped), node blocks of n_eq x n_eq.
// dolfin's pattern is block-dense, so point ILU(0) in the node-interleaved dof order equals block ILU(0) with un-pivoted LU
// of the diagonal blocks -- which is what is computed.
//
// GPU formulation: MULTICOLOUR ordering.  The mesh nodes are greedily coloured (no two neighbours share a colour); with the
// rows ordered colour by colour, every row of a colour depends only on rows of EARLIER colours, both in the factorisation and
// in the triangular solves, so each colour is one fully parallel kernel (one thread per block row): C launches to factor,
// 2 C launches per application (C = 3..7 on triangulations).  The factorisation is the exact ILU(0) of the permuted matrix
// (up-looking IKJ with precomputed update lists); it is NOT the natural-ordering ILU(0) PETSc computes -- colour orderings are
// weaker, so Krylov counts are an upper bound of the reference's.  Everything is deterministic (no atomics).
#include <algorithm>
#include <vector>

#include "fedmx_internal.h"

struct IluPlan {
    int n_colors = 0;
    std::vector<int32_t> color_ptr;        // host: rows of colour c are perm positions [color_ptr[c], color_ptr[c+1])
    int64_t n = 0, n_entries = 0, n_updates = 0;
    int32_t *d_row_node = nullptr;         // [n] node of permuted row p
    int32_t *d_iptr = nullptr;             // [n+1] entries of row p (L part first, in elimination order, then U part)
    int32_t *d_ldiv = nullptr;             // [n]   first U entry
    int32_t *d_col = nullptr;              // [n_entries] column NODE
    int32_t *d_colpos = nullptr;           // [n_entries] permuted position of the column (index into dval/dinv)
    int64_t *d_src = nullptr;              // [n_entries] index of entry 0 of the block in the SELL value array
    int64_t *d_dsrc = nullptr;             // [n] same for the diagonal block
    int32_t *d_uptr = nullptr;             // [n_entries+1] update triples of L entry e: uptr[e]..uptr[e+1]  (U entries: empty)
    int32_t *d_usrc = nullptr;             // [n_updates] entry index f of row k (U_kj)
    int32_t *d_udst = nullptr;             // [n_updates] entry index g of row p, or -1 for the diagonal block
    double *d_val = nullptr;               // [n_entries][NB] row-major blocks: L_ik (unit lower) and U_ij
    double *d_diag = nullptr;              // [n][NB] U_ii
    double *d_dinv = nullptr;              // [n][NB] U_ii^-1
    double *d_y = nullptr;                 // [n][NEQ] work vector in permuted order
};

namespace {

template <int NEQ>
__device__ __forceinline__ void matmul(const double *a, const double *b, double *c) {        // c = a b
#pragma unroll
    for (int i = 0; i < NEQ; ++i)
#pragma unroll
        for (int j = 0; j < NEQ; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < NEQ; ++k) s = fma(a[i * NEQ + k], b[k * NEQ + j], s);
            c[i * NEQ + j] = s;
        }
}

template <int NEQ>
__device__ __forceinline__ void invert(const double *m, double *out) {                      // Gauss-Jordan without pivoting (ILU semantics)
    double a[NEQ][NEQ], b[NEQ][NEQ];
#pragma unroll
    for (int i = 0; i < NEQ; ++i)
#pragma unroll
        for (int j = 0; j < NEQ; ++j) { a[i][j] = m[i * NEQ + j]; b[i][j] = i == j ? 1.0 : 0.0; }
#pragma unroll
    for (int c = 0; c < NEQ; ++c) {
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
#pragma unroll
    for (int i = 0; i < NEQ; ++i)
#pragma unroll
        for (int j = 0; j < NEQ; ++j) out[i * NEQ + j] = b[i][j];
}

template <int NEQ>
__global__ void ilu_gather_kernel(int64_t n_entries, int64_t n, const int64_t *__restrict__ src, const int64_t *__restrict__ dsrc,
                                  const double *__restrict__ aval, double *__restrict__ val, double *__restrict__ diag) {
    constexpr int NB = NEQ * NEQ;
    const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e < n_entries) {
        const double *s = aval + src[e];
#pragma unroll
        for (int q = 0; q < NB; ++q) val[e * NB + q] = s[q * FEDMX_SLICE];
    } else if (e < n_entries + n) {
        const int64_t p = e - n_entries;
        const double *s = aval + dsrc[p];
#pragma unroll
        for (int q = 0; q < NB; ++q) diag[p * NB + q] = s[q * FEDMX_SLICE];
    }
}

// rows [p0, p1) of one colour: eliminate with the rows of earlier colours (already final), then invert the pivot block
template <int NEQ>
__global__ void ilu_factor_kernel(int p0, int p1, const int32_t *__restrict__ iptr, const int32_t *__restrict__ ldiv,
                                  const int32_t *__restrict__ colpos, const int32_t *__restrict__ uptr,
                                  const int32_t *__restrict__ usrc, const int32_t *__restrict__ udst, double *__restrict__ val,
                                  double *__restrict__ diag, double *__restrict__ dinv) {
    constexpr int NB = NEQ * NEQ;
    const int p = p0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= p1) return;
    for (int e = iptr[p]; e < ldiv[p]; ++e) {
        double a[NB], lik[NB];
#pragma unroll
        for (int q = 0; q < NB; ++q) a[q] = val[int64_t(e) * NB + q];
        matmul<NEQ>(a, dinv + int64_t(colpos[e]) * NB, lik);                                  // L_ik = A_ik U_kk^-1
#pragma unroll
        for (int q = 0; q < NB; ++q) val[int64_t(e) * NB + q] = lik[q];
        for (int u = uptr[e]; u < uptr[e + 1]; ++u) {
            double prod[NB];
            matmul<NEQ>(lik, val + int64_t(usrc[u]) * NB, prod);                              // L_ik U_kj
            double *dst = udst[u] < 0 ? diag + int64_t(p) * NB : val + int64_t(udst[u]) * NB;
#pragma unroll
            for (int q = 0; q < NB; ++q) dst[q] -= prod[q];
        }
    }
    invert<NEQ>(diag + int64_t(p) * NB, dinv + int64_t(p) * NB);
}

// forward substitution (unit lower): y_p = r_node(p) - sum_k L_pk y_k
template <int NEQ>
__global__ void ilu_lower_kernel(int p0, int p1, const int32_t *__restrict__ row_node, const int32_t *__restrict__ iptr,
                                 const int32_t *__restrict__ ldiv, const int32_t *__restrict__ colpos, const double *__restrict__ val,
                                 const double *__restrict__ r, double *__restrict__ y) {
    constexpr int NB = NEQ * NEQ;
    const int p = p0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= p1) return;
    double acc[NEQ];
    const int64_t node = row_node[p];
#pragma unroll
    for (int i = 0; i < NEQ; ++i) acc[i] = r[NEQ * node + i];
    for (int e = iptr[p]; e < ldiv[p]; ++e) {
        const double *yk = y + int64_t(colpos[e]) * NEQ;
        const double *l = val + int64_t(e) * NB;
#pragma unroll
        for (int i = 0; i < NEQ; ++i)
#pragma unroll
            for (int j = 0; j < NEQ; ++j) acc[i] = fma(-l[i * NEQ + j], yk[j], acc[i]);
    }
#pragma unroll
    for (int i = 0; i < NEQ; ++i) y[int64_t(p) * NEQ + i] = acc[i];
}

// backward substitution: x_p = U_pp^-1 (y_p - sum_j U_pj x_j); x overwrites y in permuted order and is scattered to z (node order)
template <int NEQ>
__global__ void ilu_upper_kernel(int p0, int p1, const int32_t *__restrict__ row_node, const int32_t *__restrict__ iptr,
                                 const int32_t *__restrict__ ldiv, const int32_t *__restrict__ colpos, const double *__restrict__ val,
                                 const double *__restrict__ dinv, double *__restrict__ y, double *__restrict__ z) {
    constexpr int NB = NEQ * NEQ;
    const int p = p0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= p1) return;
    double acc[NEQ], x[NEQ];
#pragma unroll
    for (int i = 0; i < NEQ; ++i) acc[i] = y[int64_t(p) * NEQ + i];
    for (int e = ldiv[p]; e < iptr[p + 1]; ++e) {
        const double *xj = y + int64_t(colpos[e]) * NEQ;
        const double *u = val + int64_t(e) * NB;
#pragma unroll
        for (int i = 0; i < NEQ; ++i)
#pragma unroll
            for (int j = 0; j < NEQ; ++j) acc[i] = fma(-u[i * NEQ + j], xj[j], acc[i]);
    }
    const double *d = dinv + int64_t(p) * NB;
    const int64_t node = row_node[p];
#pragma unroll
    for (int i = 0; i < NEQ; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NEQ; ++j) s = fma(d[i * NEQ + j], acc[j], s);
        x[i] = s;
    }
#pragma unroll
    for (int i = 0; i < NEQ; ++i) { y[int64_t(p) * NEQ + i] = x[i]; z[NEQ * node + i] = x[i]; }
}

int ilu_symbolic(fedmx_ctx *ctx) {
    IluPlan *P = new IluPlan();
    ctx->ilu = P;
    const int64_t n = ctx->n_owned;
    const auto &rp = ctx->h_rowptr;
    const auto &ci = ctx->h_colidx;
    // greedy colouring over the owned-owned graph, nodes visited in index order, smallest free colour
    std::vector<int32_t> color(n, -1);
    int nc = 0;
    {
        std::vector<int> mark;
        for (int64_t v = 0; v < n; ++v) {
            mark.assign(nc + 1, 0);
            for (int32_t q = rp[v]; q < rp[v + 1]; ++q) {
                const int32_t w = ci[q];
                if (w < n && w != v && color[w] >= 0) mark[color[w]] = 1;
            }
            int c = 0;
            while (c < nc && mark[c]) ++c;
            color[v] = c;
            nc = std::max(nc, c + 1);
        }
    }
    P->n_colors = nc;
    P->n = n;
    std::vector<int32_t> pos(n), row_node(n);
    P->color_ptr.assign(nc + 1, 0);
    for (int64_t v = 0; v < n; ++v) P->color_ptr[color[v] + 1]++;
    for (int c = 0; c < nc; ++c) P->color_ptr[c + 1] += P->color_ptr[c];
    {
        std::vector<int32_t> fill(P->color_ptr.begin(), P->color_ptr.end() - 1);
        // rows of a colour keep the STORAGE order of the matrix (tiles): neighbouring threads touch neighbouring data
        for (int64_t r = 0; r < n; ++r) {
            const int32_t v = ctx->h_row_node[r];
            pos[v] = fill[color[v]]++;
        }
    }
    for (int64_t v = 0; v < n; ++v) row_node[pos[v]] = int32_t(v);
    const int NB = ctx->n_eq * ctx->n_eq;
    std::vector<int32_t> iptr(n + 1, 0), ldiv(n), col, colpos, uptr(1, 0), usrc, udst;
    std::vector<int64_t> src, dsrc(n);
    std::vector<std::pair<int32_t, int32_t>> ent;          // (pos of column, k-th stored column of the node row)
    for (int64_t p = 0; p < n; ++p) {
        const int32_t v = row_node[p];
        const int64_t sr = ctx->h_node_row[v];
        const int64_t s = sr / FEDMX_SLICE, lane = sr % FEDMX_SLICE;
        ent.clear();
        for (int32_t q = rp[v]; q < rp[v + 1]; ++q) {
            const int32_t w = ci[q];
            const int64_t slot = (int64_t(ctx->h_slice_ptr[s]) + int64_t(q - rp[v]) * FEDMX_SLICE) * NB + lane;
            if (w == v) { dsrc[p] = slot; continue; }
            if (w >= n) continue;                           // ghost column: outside this rank's diagonal block
            ent.emplace_back(pos[w], q - rp[v]);
        }
        std::sort(ent.begin(), ent.end());                  // elimination order = ascending permuted position
        int nl = 0;
        for (auto &e : ent) {
            if (e.first < p) ++nl;
            col.push_back(row_node[e.first]);
            colpos.push_back(e.first);
            src.push_back((int64_t(ctx->h_slice_ptr[s]) + int64_t(e.second) * FEDMX_SLICE) * NB + lane);
        }
        ldiv[p] = iptr[p] + nl;
        iptr[p + 1] = iptr[p] + int32_t(ent.size());
    }
    // update lists: L entry e = (p,k) meets U entry f = (k,j) at g = (p,j) or at the diagonal (j == p)
    for (int64_t p = 0; p < n; ++p) {
        for (int32_t e = iptr[p]; e < iptr[p + 1]; ++e) {
            if (e < ldiv[p]) {
                const int32_t k = colpos[e];
                for (int32_t f = ldiv[k]; f < iptr[k + 1]; ++f) {
                    const int32_t j = colpos[f];
                    if (j == p) { usrc.push_back(f); udst.push_back(-1); continue; }
                    const int32_t *b = colpos.data() + iptr[p], *en = colpos.data() + iptr[p + 1];
                    const int32_t *it = std::lower_bound(b, en, j);
                    if (it != en && *it == j) { usrc.push_back(f); udst.push_back(int32_t(it - colpos.data())); }
                }
            }
            uptr.push_back(int32_t(usrc.size()));
        }
    }
    P->n_entries = int64_t(col.size());
    P->n_updates = int64_t(usrc.size());
    int rc;
    if ((rc = fedmx_upload(ctx, &P->d_row_node, row_node))) return rc;
    if ((rc = fedmx_upload(ctx, &P->d_iptr, iptr))) return rc;
    if ((rc = fedmx_upload(ctx, &P->d_ldiv, ldiv))) return rc;
    if ((rc = fedmx_upload(ctx, &P->d_col, col))) return rc;
    if ((rc = fedmx_upload(ctx, &P->d_colpos, colpos))) return rc;
    if ((rc = fedmx_upload(ctx, &P->d_src, src))) return rc;
    if ((rc = fedmx_upload(ctx, &P->d_dsrc, dsrc))) return rc;
    if ((rc = fedmx_upload(ctx, &P->d_uptr, uptr))) return rc;
    if ((rc = fedmx_upload(ctx, &P->d_usrc, usrc))) return rc;
    if ((rc = fedmx_upload(ctx, &P->d_udst, udst))) return rc;
    if ((rc = fedmx_dev_alloc(ctx, &P->d_val, size_t(P->n_entries) * NB))) return rc;
    if ((rc = fedmx_dev_alloc(ctx, &P->d_diag, size_t(n) * NB))) return rc;
    if ((rc = fedmx_dev_alloc(ctx, &P->d_dinv, size_t(n) * NB))) return rc;
    if ((rc = fedmx_dev_alloc(ctx, &P->d_y, size_t(n) * ctx->n_eq))) return rc;
    return FEDMX_OK;
}

template <int NEQ>
int ilu_factor_t(fedmx_ctx *ctx) {
    IluPlan *P = ctx->ilu;
    const int threads = 128;
    const int64_t tot = P->n_entries + P->n;
    ilu_gather_kernel<NEQ><<<unsigned((tot + threads - 1) / threads), threads, 0, ctx->stream>>>(P->n_entries, P->n, P->d_src, P->d_dsrc,
                                                                                              ctx->d_val, P->d_val, P->d_diag);
    ctx->launches++;
    for (int c = 0; c < P->n_colors; ++c) {
        const int p0 = P->color_ptr[c], p1 = P->color_ptr[c + 1];
        if (p1 == p0) continue;
        ilu_factor_kernel<NEQ><<<unsigned((p1 - p0 + threads - 1) / threads), threads, 0, ctx->stream>>>(
            p0, p1, P->d_iptr, P->d_ldiv, P->d_colpos, P->d_uptr, P->d_usrc, P->d_udst, P->d_val, P->d_diag, P->d_dinv);
        ctx->launches++;
    }
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    return FEDMX_OK;
}

template <int NEQ>
int ilu_solve_t(fedmx_ctx *ctx, const double *d_r, double *d_z) {
    IluPlan *P = ctx->ilu;
    const int threads = 128;
    for (int c = 0; c < P->n_colors; ++c) {
        const int p0 = P->color_ptr[c], p1 = P->color_ptr[c + 1];
        if (p1 == p0) continue;
        ilu_lower_kernel<NEQ><<<unsigned((p1 - p0 + threads - 1) / threads), threads, 0, ctx->stream>>>(
            p0, p1, P->d_row_node, P->d_iptr, P->d_ldiv, P->d_colpos, P->d_val, d_r, P->d_y);
        ctx->launches++;
    }
    for (int c = P->n_colors - 1; c >= 0; --c) {
        const int p0 = P->color_ptr[c], p1 = P->color_ptr[c + 1];
        if (p1 == p0) continue;
        ilu_upper_kernel<NEQ><<<unsigned((p1 - p0 + threads - 1) / threads), threads, 0, ctx->stream>>>(
            p0, p1, P->d_row_node, P->d_iptr, P->d_ldiv, P->d_colpos, P->d_val, P->d_dinv, P->d_y, d_z);
        ctx->launches++;
    }
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    return FEDMX_OK;
}

}  // namespace

#define ILU_DISPATCH(fn, ...)                                                              \
    switch (ctx->n_eq) {                                                                   \
        case 1: return fn<1>(__VA_ARGS__);                                                 \
        case 2: return fn<2>(__VA_ARGS__);                                                 \
        case 3: return fn<3>(__VA_ARGS__);                                                 \
        case 4: return fn<4>(__VA_ARGS__);                                                 \
        case 5: return fn<5>(__VA_ARGS__);                                                 \
        default: ctx->err = "ILU(0): unsupported n_eq"; return FEDMX_EUNSUPPORTED;         \
    }

int fedmx_ilu0_factor_internal(fedmx_ctx *ctx) {
    if (!ctx->ilu) {
        int rc = ilu_symbolic(ctx);
        if (rc) return rc;
    }
    ProfScope prof(ctx, FEDMX_PROF_PRECOND);
    ILU_DISPATCH(ilu_factor_t, ctx)
}

int fedmx_ilu0_solve_internal(fedmx_ctx *ctx, const double *d_r, double *d_z) {
    ProfScope prof(ctx, FEDMX_PROF_PRECOND);
    ILU_DISPATCH(ilu_solve_t, ctx, d_r, d_z)
}

void fedmx_ilu0_free(fedmx_ctx *ctx) {
    delete ctx->ilu;                 // device arrays belong to the context's allocation list
    ctx->ilu = nullptr;
}

extern "C" int fedmx_ilu0_factor(fedmx_ctx *ctx) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, ctx->d_val && ctx->n_owned > 0, "fedmx_ilu0_factor: no matrix");
    return fedmx_ilu0_factor_internal(ctx);
}

extern "C" int fedmx_ilu0_solve(fedmx_ctx *ctx, const double *d_r, double *d_z) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, ctx->ilu && d_r && d_z && d_r != d_z, "fedmx_ilu0_solve: factor first; r and z must differ");
    return fedmx_ilu0_solve_internal(ctx, d_r, d_z);
}

// h_order (optional, n_owned entries): the elimination order, h_order[p] = node of permuted row p
extern "C" int fedmx_ilu0_info(fedmx_ctx *ctx, int *n_colors, int64_t *n_blocks, int32_t *h_order) {
    if (!ctx || !ctx->ilu) return FEDMX_EINVAL;
    if (n_colors) *n_colors = ctx->ilu->n_colors;
    if (n_blocks) *n_blocks = ctx->ilu->n_entries + ctx->ilu->n;
    if (h_order) FEDMX_CHECK_CUDA(ctx, cudaMemcpy(h_order, ctx->ilu->d_row_node, size_t(ctx->ilu->n) * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return FEDMX_OK;
}
