// 1-D P2 (quadratic Lagrange on intervals) log-density balance equation: BASELINE.json configs[0], the reference's
// CPU-runnable example examples/time_of_flight_1D/fedm-tof_1d.py (IntervalMesh(4000, 0, 1e-3), FunctionSpace(mesh, 'P', 2),
// Gamma = -grad(D e^u) + w e^u, weak_form_balance_equation_log_representation in Cartesian form: fedm/functions.py:350-368).
//
// 8 001 unknowns: the case is latency-bound, so the design is the simplest deterministic one.
//   assemble  one thread per ROW (dof) evaluates the element integrals of the one (midpoint) or two (vertex) cells it belongs to
//             with the 5-point Gauss-Legendre rule UFL's degree estimate (8) selects, and writes its residual entry and its
//             five band entries -- no scatter, no atomics;
//   solve     the Jacobian is pentadiagonal: banded LU without pivoting, rows streamed through registers by one warp lane
//             (the reference uses a direct solver, linear_solver = 'mumps', fedm-tof_1d.py:21).
// dof numbering: coordinate order, dof 2c / 2c+1 / 2c+2 = left vertex / midpoint / right vertex of cell c.
#include <math.h>

#include "fedmx_internal.h"

namespace {

__device__ __constant__ double kGL5_x[5] = {0.04691007703066800360, 0.23076534494715845448, 0.5, 0.76923465505284154552,
                                           0.95308992296933199640};
__device__ __constant__ double kGL5_w[5] = {0.11846344252809454376, 0.23931433524968323402, 0.28444444444444444444,
                                           0.23931433524968323402, 0.11846344252809454376};

struct P2Params {
    double h, ct, D, w;
};

// contribution of cell c to row `a` (local index 0..2): residual R and the three Jacobian entries J[b]
__device__ __forceinline__ void p2_cell_row(const P2Params &P, const double *__restrict__ u, const double *__restrict__ hist,
                                            const double *__restrict__ f, int64_t c, int a, double &R, double J[3]) {
    const double u0 = u[2 * c], u1 = u[2 * c + 1], u2 = u[2 * c + 2];
    const double h0 = hist[2 * c], h1 = hist[2 * c + 1], h2 = hist[2 * c + 2];
    const double f0 = f[2 * c], f1 = f[2 * c + 1], f2 = f[2 * c + 2];
    const double ih = 1.0 / P.h;
    R = 0.0;
    J[0] = J[1] = J[2] = 0.0;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const double xi = kGL5_x[q], W = kGL5_w[q] * P.h;
        const double phi[3] = {(1.0 - xi) * (1.0 - 2.0 * xi), 4.0 * xi * (1.0 - xi), xi * (2.0 * xi - 1.0)};
        const double dphi[3] = {(4.0 * xi - 3.0) * ih, (4.0 - 8.0 * xi) * ih, (4.0 * xi - 1.0) * ih};
        const double uq = u0 * phi[0] + u1 * phi[1] + u2 * phi[2];
        const double duq = u0 * dphi[0] + u1 * dphi[1] + u2 * dphi[2];
        const double hq = h0 * phi[0] + h1 * phi[1] + h2 * phi[2];
        const double fq = f0 * phi[0] + f1 * phi[1] + f2 * phi[2];
        const double n = exp(uq);
        const double beta = P.ct * uq + hq;
        const double gamma = -P.D * n * duq + P.w * n;
        R += ((n * beta - fq) * phi[a] - gamma * dphi[a]) * W;
        const double c0 = n * (beta + P.ct) * W, c1 = gamma * W, c2 = -P.D * n * W;
#pragma unroll
        for (int b = 0; b < 3; ++b) J[b] += c0 * phi[b] * phi[a] - c1 * dphi[a] * phi[b] - c2 * dphi[a] * dphi[b];
    }
}

// band[i][k], k = 0..4 <-> column i - 2 + k
__global__ void p2line_assemble_kernel(int64_t n_cells, P2Params P, const double *__restrict__ u, const double *__restrict__ hist,
                                       const double *__restrict__ f, double *__restrict__ res, double *__restrict__ band) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t ndof = 2 * n_cells + 1;
    if (i >= ndof) return;
    double b[5] = {0, 0, 0, 0, 0}, R = 0.0, r, J[3];
    if (i & 1) {                                        // midpoint of cell c: columns i-1, i, i+1
        p2_cell_row(P, u, hist, f, (i - 1) / 2, 1, r, J);
        R = r; b[1] = J[0]; b[2] = J[1]; b[3] = J[2];
    } else {
        const int64_t c = i / 2;
        if (c > 0) {                                    // right vertex of cell c-1: columns i-2, i-1, i
            p2_cell_row(P, u, hist, f, c - 1, 2, r, J);
            R += r; b[0] += J[0]; b[1] += J[1]; b[2] += J[2];
        }
        if (c < n_cells) {                              // left vertex of cell c: columns i, i+1, i+2
            p2_cell_row(P, u, hist, f, c, 0, r, J);
            R += r; b[2] += J[0]; b[3] += J[1]; b[4] += J[2];
        }
    }
    res[i] = R;
    if (band) {
#pragma unroll
        for (int k = 0; k < 5; ++k) band[5 * i + k] = b[k];
    }
}

// Pentadiagonal LU without pivoting + forward/backward substitution, in place (band and rhs are destroyed; x may alias rhs).
// One thread: the recurrence is strictly sequential; rows i+1, i+2 are updated by row i.
__global__ void p2line_solve_kernel(int64_t n, double *__restrict__ band, double *__restrict__ rhs, double *__restrict__ x) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (int64_t i = 0; i < n; ++i) {
        const double piv = 1.0 / band[5 * i + 2];
        const double a1 = band[5 * i + 3], a2 = band[5 * i + 4], bi = rhs[i];
        if (i + 1 < n) {
            const double l = band[5 * (i + 1) + 1] * piv;            // column i of row i+1
            band[5 * (i + 1) + 1] = l;
            band[5 * (i + 1) + 2] -= l * a1;
            band[5 * (i + 1) + 3] -= l * a2;
            rhs[i + 1] -= l * bi;
        }
        if (i + 2 < n) {
            const double l = band[5 * (i + 2) + 0] * piv;            // column i of row i+2
            band[5 * (i + 2) + 0] = l;
            band[5 * (i + 2) + 1] -= l * a1;
            band[5 * (i + 2) + 2] -= l * a2;
            rhs[i + 2] -= l * bi;
        }
    }
    for (int64_t i = n - 1; i >= 0; --i) {
        double s = rhs[i];
        if (i + 1 < n) s -= band[5 * i + 3] * x[i + 1];
        if (i + 2 < n) s -= band[5 * i + 4] * x[i + 2];
        x[i] = s / band[5 * i + 2];
    }
}

}  // namespace

extern "C" int fedmx_p2line_assemble(fedmx_ctx *ctx, int64_t n_cells, double h, double dt, double dt_old, double D, double w,
                                     const double *d_u, const double *d_u_old, const double *d_u_old1, const double *d_f,
                                     double *d_hist, double *d_res, double *d_band) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, n_cells > 0 && d_u && d_u_old && d_u_old1 && d_f && d_hist && d_res, "fedmx_p2line_assemble: bad arguments");
    const int64_t ndof = 2 * n_cells + 1;
    const double tr = dt / dt_old;
    P2Params P{h, (1.0 + 2.0 * tr) / ((1.0 + tr) * dt), D, w};
    // BDF2 history (-(1+tr)^2 u_old + tr^2 u_old1) / ((1+tr) dt): the same kernel as the P1 path, one "equation" per dof
    int rc = fedmx_bdf2_history_raw(ctx, ndof, d_u_old, d_u_old1, dt, dt_old, d_hist);
    if (rc) return rc;
    const int threads = 128;
    p2line_assemble_kernel<<<unsigned((ndof + threads - 1) / threads), threads, 0, ctx->stream>>>(n_cells, P, d_u, d_hist, d_f, d_res, d_band);
    ctx->launches++;
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    return FEDMX_OK;
}

extern "C" int fedmx_p2line_solve(fedmx_ctx *ctx, int64_t ndof, double *d_band, double *d_rhs, double *d_x) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, ndof > 0 && d_band && d_rhs && d_x, "fedmx_p2line_solve: bad arguments");
    p2line_solve_kernel<<<1, 32, 0, ctx->stream>>>(ndof, d_band, d_rhs, d_x);
    ctx->launches++;
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    return FEDMX_OK;
}
