// CPU restatement of the reference's streamer hot path in C++/OpenMP (oracle -- TEST INFRASTRUCTURE / REPORTED CPU BASELINE
// ONLY; never part of the product path, never linked into libfedmx).
//
// What it restates (reference = INP-PM/FEDM over FEniCS 2019.1.0 + PETSc, none of which can run in this image):
//   * the forms of tests/integrated_tests/streamer_discharge/fedm_streamer.py:243-279 over fedm/functions.py:219-237 (Flux),
//     :350-368 (log-BDF2 balance equation), :401 (Poisson), :523-524 ('Neumann' drift flux), quadrature degree 2
//     (fedm_streamer.py:29), evaluated per cell and quadrature point the way an FFC-generated tabulate_tensor does; the
//     Jacobian is the exact derivative (forward-mode AD over the 9 cell unknowns = UFL's derivative(F, u_new, u), :296-297);
//   * dolfin's assembly + DirichletBC.apply (fedm/functions.py:188-202) into a scalar CSR matrix with dolfin's pattern;
//   * the reference TEST's solver (fedm_streamer.py:32, 302-305): SNES newtonls/basic, KSP GMRES(30), classical Gram-Schmidt,
//     left preconditioning, rtol 1e-5 on the preconditioned residual, PC = block Jacobi with ILU(0) on every block -- PETSc's
//     default under `mpirun -np P`.  The P "ranks" are P OpenMP threads, each owning a contiguous strip of mesh nodes
//     (owner-computes assembly with redundantly evaluated ghost cells, like dolfin's MPI decomposition).
// Parity: tests/test_oracle_cport.py compares residual / Jacobian entries with the numpy oracle (<= 1e-12 of the row scale),
// ILU(0) with oracle/newton.py:ilu0 and Newton counts / states of whole steps.
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace {

constexpr double kTwoPi = 6.283185307179586476925286766559;
constexpr double kKappa = 1.6021766208e-19 / 8.854187817e-12;      // e / eps0 (fedm/physical_constants.py:5,7)

// ---- forward-mode AD over the 9 unknowns of a cell ---------------------------------------------------------------------
template <int N>
struct Dual {
    double v;
    double d[N];
};
template <int N> inline Dual<N> cst(double a) { Dual<N> r; r.v = a; for (int i = 0; i < N; ++i) r.d[i] = 0.0; return r; }
template <int N> inline Dual<N> operator+(const Dual<N> &a, const Dual<N> &b) { Dual<N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N> &a, const Dual<N> &b) { Dual<N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N> &a) { Dual<N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <int N> inline Dual<N> operator*(const Dual<N> &a, const Dual<N> &b) { Dual<N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int N> inline Dual<N> operator*(double a, const Dual<N> &b) { Dual<N> r; r.v = a * b.v; for (int i = 0; i < N; ++i) r.d[i] = a * b.d[i]; return r; }
template <int N> inline Dual<N> operator*(const Dual<N> &b, double a) { return a * b; }
template <int N> inline Dual<N> operator+(const Dual<N> &a, double b) { Dual<N> r = a; r.v += b; return r; }
template <int N> inline Dual<N> operator-(const Dual<N> &a, double b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> inline Dual<N> chain(const Dual<N> &a, double f, double df) { Dual<N> r; r.v = f; for (int i = 0; i < N; ++i) r.d[i] = df * a.d[i]; return r; }
template <int N> inline Dual<N> dexp(const Dual<N> &a) { const double e = exp(a.v); return chain(a, e, e); }
template <int N> inline Dual<N> dsqrt(const Dual<N> &a) { const double s = sqrt(a.v); return chain(a, s, 0.5 / s); }
template <int N> inline Dual<N> dpow(const Dual<N> &a, double p) { const double f = pow(a.v, p); return chain(a, f, p * f / a.v); }
template <int N> inline Dual<N> dinv(const Dual<N> &a) { const double f = 1.0 / a.v; return chain(a, f, -f * f); }
// plain doubles (residual-only instantiation)
inline double cst0(double a) { return a; }
inline double dexp(double a) { return exp(a); }
inline double dsqrt(double a) { return sqrt(a); }
inline double dpow(double a, double p) { return pow(a, p); }
inline double dinv(double a) { return 1.0 / a; }
template <class T> struct Val { static double of(const T &a) { return a.v; } };
template <> struct Val<double> { static double of(const double &a) { return a; } };

struct Csr {
    int64_t n = 0;
    std::vector<int64_t> ptr;
    std::vector<int32_t> col;
    std::vector<double> val;
};

struct Ref {
    int64_t nv = 0, nc = 0, ndof = 0;
    int P = 1;
    std::vector<double> coords;
    std::vector<int32_t> cells;
    // node adjacency (sorted, incl. self) and the scalar matrix with dolfin's pattern: row 3*i+k, columns 3*j+l
    std::vector<int64_t> nptr;
    std::vector<int32_t> ncol;
    Csr A;
    std::vector<int32_t> cellpos;                 // [nc][3][3]: index of block (node_a, node_b) inside node row a
    // "ranks": thread t owns nodes [lo[t], lo[t+1]) and works on the cells listed in tcells[t] (every cell touching a node)
    std::vector<int64_t> lo;
    std::vector<std::vector<int32_t>> tcells;
    // marked exterior facets carrying the electron drift flux: (a, b, opposite vertex)
    std::vector<int32_t> facets;
    std::vector<std::vector<int32_t>> tfacets;
    // Dirichlet on component 2
    std::vector<char> is_bc;
    std::vector<double> bc_val;
    // state
    std::vector<double> uold, uold1;
    double dt = 5e-12, dt_old = 1e30;
    // ILU(0) of the P diagonal blocks (local CSR copies)
    std::vector<Csr> L;
    std::vector<std::vector<int64_t>> Ldiag;
    std::vector<std::vector<int64_t>> Lsrc;       // position in A.val of every entry of the block
    // Krylov work
    std::vector<double> V, w, tmp;
    int restart_alloc = 0;
    double t_asm_jac = 0, t_asm_res = 0, t_ilu = 0, t_krylov = 0;
};

template <class T>
inline void transport(const T &Em, T &mu, T &D, T &alpha) {
    // e_Nb.dat / e_ND.dat / fedm_streamer.py:247 (strings eval'ed by the reference)
    mu = 2.3987 * dpow(Em, -0.26);
    D = 4.3628e-3 * dpow(Em, 0.22);
    alpha = (4.3666e26 * dpow(Em, -3.0) + 1.1944e6) * dexp(-2.73e7 * dinv(Em)) - 340.75;
}

// Element residual (9 entries, [a][k]) of one cell, generic in the scalar type
template <class T>
inline void cell_residual(const Ref &R, int64_t c, const T U[3][3], T out[3][3]) {
    const int32_t *cv = &R.cells[3 * c];
    double x[3][2], uo[3][2], uo1[3][2];
    for (int a = 0; a < 3; ++a) {
        x[a][0] = R.coords[2 * cv[a]];
        x[a][1] = R.coords[2 * cv[a] + 1];
        for (int k = 0; k < 2; ++k) {
            uo[a][k] = R.uold[3 * int64_t(cv[a]) + k];
            uo1[a][k] = R.uold1[3 * int64_t(cv[a]) + k];
        }
    }
    const double j00 = x[1][0] - x[0][0], j01 = x[2][0] - x[0][0], j10 = x[1][1] - x[0][1], j11 = x[2][1] - x[0][1];
    const double det = j00 * j11 - j01 * j10, inv = 1.0 / det;
    // physical gradients of the P1 basis functions
    double g[3][2];
    g[1][0] = j11 * inv; g[1][1] = -j01 * inv;
    g[2][0] = -j10 * inv; g[2][1] = j00 * inv;
    g[0][0] = -g[1][0] - g[2][0]; g[0][1] = -g[1][1] - g[2][1];
    const double adet = fabs(det);
    T Hx = U[0][1] * g[0][0] + U[1][1] * g[1][0] + U[2][1] * g[2][0];
    T Hy = U[0][1] * g[0][1] + U[1][1] * g[1][1] + U[2][1] * g[2][1];
    T Gx = U[0][2] * g[0][0] + U[1][2] * g[1][0] + U[2][2] * g[2][0];
    T Gy = U[0][2] * g[0][1] + U[1][2] * g[1][1] + U[2][2] * g[2][1];
    T Em = dsqrt(Gx * Gx + Gy * Gy);
    T mu, D, alpha;
    transport(Em, mu, D, alpha);
    const double dt = R.dt, tr = R.dt / R.dt_old, trp1 = 1.0 + tr, tr2p1 = 1.0 + 2.0 * tr;
    for (int a = 0; a < 3; ++a)
        for (int k = 0; k < 3; ++k) out[a][k] = U[0][0] * 0.0;
    // FIAT degree-2 rule: points (1/6,1/6), (1/6,2/3), (2/3,1/6) in reference coordinates, weights 1/6
    static const double lam[3][3] = {{2.0 / 3.0, 1.0 / 6.0, 1.0 / 6.0}, {1.0 / 6.0, 1.0 / 6.0, 2.0 / 3.0}, {1.0 / 6.0, 2.0 / 3.0, 1.0 / 6.0}};
    for (int q = 0; q < 3; ++q) {
        const double *l = lam[q];
        const double r = l[0] * x[0][0] + l[1] * x[1][0] + l[2] * x[2][0];
        const double w = adet / 6.0;
        const double two_pi_r = kTwoPi * r;
        T ui = U[0][0] * l[0] + U[1][0] * l[1] + U[2][0] * l[2];
        T ue = U[0][1] * l[0] + U[1][1] * l[1] + U[2][1] * l[2];
        T ni = dexp(ui), ne = dexp(ue);
        T src = alpha * mu * Em * ne;
        double oi = 0, o1i = 0, oe = 0, o1e = 0;
        for (int a = 0; a < 3; ++a) { oi += l[a] * uo[a][0]; o1i += l[a] * uo1[a][0]; oe += l[a] * uo[a][1]; o1e += l[a] * uo1[a][1]; }
        T upi = (ui * tr2p1 - (trp1 * trp1 * oi) + (tr * tr * o1i)) * (1.0 / trp1);
        T upe = (ue * tr2p1 - (trp1 * trp1 * oe) + (tr * tr * o1e)) * (1.0 / trp1);
        T si = two_pi_r * (ni * upi * (1.0 / dt) - src);
        T se = two_pi_r * (ne * upe * (1.0 / dt) - src);
        T gamx = (mu * Gx - D * Hx) * ne, gamy = (mu * Gy - D * Hy) * ne;    // Gamma = -D grad(e^u) - mu E e^u, E = -grad(Phi)
        T f2 = (ni - ne) * kKappa;
        for (int a = 0; a < 3; ++a) {
            out[a][0] = out[a][0] + (w * l[a]) * si;
            out[a][1] = out[a][1] + w * (l[a] * se - two_pi_r * (gamx * g[a][0] + gamy * g[a][1]));
            out[a][2] = out[a][2] + w * (two_pi_r * (Gx * g[a][0] + Gy * g[a][1]) - (two_pi_r * l[a]) * f2);
        }
    }
}

// Exterior-facet residual on the facet (a, b) of a cell with opposite vertex c: 2 pi r mu (G.n) e^{u_e} v ds, 2-point Gauss
template <class T>
inline void facet_residual(const Ref &R, const int32_t fv[3], const T U[3][3], T out[3]) {
    double x[3][2];
    for (int a = 0; a < 3; ++a) { x[a][0] = R.coords[2 * fv[a]]; x[a][1] = R.coords[2 * fv[a] + 1]; }
    const double j00 = x[1][0] - x[0][0], j01 = x[2][0] - x[0][0], j10 = x[1][1] - x[0][1], j11 = x[2][1] - x[0][1];
    const double det = j00 * j11 - j01 * j10, inv = 1.0 / det;
    double g[3][2];
    g[1][0] = j11 * inv; g[1][1] = -j01 * inv;
    g[2][0] = -j10 * inv; g[2][1] = j00 * inv;
    g[0][0] = -g[1][0] - g[2][0]; g[0][1] = -g[1][1] - g[2][1];
    T Gx = U[0][2] * g[0][0] + U[1][2] * g[1][0] + U[2][2] * g[2][0];
    T Gy = U[0][2] * g[0][1] + U[1][2] * g[1][1] + U[2][2] * g[2][1];
    T Em = dsqrt(Gx * Gx + Gy * Gy);
    T mu, D, alpha;
    transport(Em, mu, D, alpha);
    const double tx = x[1][0] - x[0][0], ty = x[1][1] - x[0][1];
    const double len = sqrt(tx * tx + ty * ty);
    double nx = ty / len, ny = -tx / len;
    if (nx * (x[2][0] - x[0][0]) + ny * (x[2][1] - x[0][1]) > 0.0) { nx = -nx; ny = -ny; }      // outward
    T flux = mu * (Gx * nx + Gy * ny);
    static const double gt[2] = {0.21132486540518711775, 0.78867513459481288225};
    for (int a = 0; a < 3; ++a) out[a] = U[0][0] * 0.0;
    for (int q = 0; q < 2; ++q) {
        const double t = gt[q], la = 1.0 - t;
        const double r = la * x[0][0] + t * x[1][0];
        T ne = dexp(U[0][1] * la + U[1][1] * t);
        T f = (0.5 * len * kTwoPi * r) * flux * ne;
        out[0] = out[0] + la * f;
        out[1] = out[1] + t * f;
    }
}

inline int thread_of(const Ref &R, int64_t node) {
    int t = int(std::upper_bound(R.lo.begin(), R.lo.end(), node) - R.lo.begin()) - 1;
    return t;
}

void build_pattern(Ref &R) {
    const int64_t nv = R.nv;
    std::vector<std::vector<int32_t>> adj(nv);
    for (int64_t c = 0; c < R.nc; ++c)
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) adj[R.cells[3 * c + a]].push_back(R.cells[3 * c + b]);
    R.nptr.assign(nv + 1, 0);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nv; ++i) {
        std::sort(adj[i].begin(), adj[i].end());
        adj[i].erase(std::unique(adj[i].begin(), adj[i].end()), adj[i].end());
    }
    for (int64_t i = 0; i < nv; ++i) R.nptr[i + 1] = R.nptr[i] + int64_t(adj[i].size());
    R.ncol.resize(R.nptr[nv]);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nv; ++i) std::copy(adj[i].begin(), adj[i].end(), R.ncol.begin() + R.nptr[i]);
    R.cellpos.resize(size_t(R.nc) * 9);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < R.nc; ++c)
        for (int a = 0; a < 3; ++a) {
            const int32_t i = R.cells[3 * c + a];
            const int32_t *b0 = &R.ncol[R.nptr[i]], *b1 = &R.ncol[R.nptr[i + 1]];
            for (int b = 0; b < 3; ++b) R.cellpos[9 * c + 3 * a + b] = int32_t(std::lower_bound(b0, b1, R.cells[3 * c + b]) - b0);
        }
    Csr &A = R.A;
    A.n = R.ndof;
    A.ptr.assign(R.ndof + 1, 0);
    for (int64_t i = 0; i < nv; ++i)
        for (int k = 0; k < 3; ++k) A.ptr[3 * i + k + 1] = 3 * (R.nptr[i + 1] - R.nptr[i]);
    for (int64_t r = 0; r < R.ndof; ++r) A.ptr[r + 1] += A.ptr[r];
    A.col.resize(A.ptr[R.ndof]);
    A.val.assign(A.ptr[R.ndof], 0.0);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nv; ++i)
        for (int k = 0; k < 3; ++k) {
            int64_t p = A.ptr[3 * i + k];
            for (int64_t q = R.nptr[i]; q < R.nptr[i + 1]; ++q)
                for (int l = 0; l < 3; ++l) A.col[p++] = 3 * R.ncol[q] + l;
        }
}

void partition(Ref &R) {
    const int P = R.P;
    R.lo.resize(P + 1);
    for (int t = 0; t <= P; ++t) R.lo[t] = R.nv * t / P;
    R.tcells.assign(P, {});
    for (int64_t c = 0; c < R.nc; ++c) {
        int seen[3];
        int ns = 0;
        for (int a = 0; a < 3; ++a) {
            const int t = thread_of(R, R.cells[3 * c + a]);
            bool dup = false;
            for (int s = 0; s < ns; ++s) dup = dup || seen[s] == t;
            if (!dup) { seen[ns++] = t; R.tcells[t].push_back(int32_t(c)); }
        }
    }
}

template <bool JAC>
void assemble(Ref &R, const double *x, double *b) {
    const double t0 = omp_get_wtime();
#pragma omp parallel num_threads(R.P)
    {
        const int t = omp_get_thread_num();
        const int64_t lo = R.lo[t], hi = R.lo[t + 1];
        if (b) memset(b + 3 * lo, 0, size_t(3 * (hi - lo)) * sizeof(double));
        if (JAC) memset(&R.A.val[R.A.ptr[3 * lo]], 0, size_t(R.A.ptr[3 * hi] - R.A.ptr[3 * lo]) * sizeof(double));
        for (int32_t c : R.tcells[t]) {
            const int32_t *cv = &R.cells[3 * int64_t(c)];
            if (JAC) {
                Dual<9> U[3][3], out[3][3];
                for (int a = 0; a < 3; ++a)
                    for (int k = 0; k < 3; ++k) {
                        U[a][k] = cst<9>(x[3 * int64_t(cv[a]) + k]);
                        U[a][k].d[3 * a + k] = 1.0;
                    }
                cell_residual(R, c, U, out);
                for (int a = 0; a < 3; ++a) {
                    const int64_t i = cv[a];
                    if (i < lo || i >= hi) continue;
                    for (int k = 0; k < 3; ++k) {
                        if (b) b[3 * i + k] += out[a][k].v;
                        double *row = &R.A.val[R.A.ptr[3 * i + k]];
                        for (int bb = 0; bb < 3; ++bb) {
                            const int pos = 3 * R.cellpos[9 * int64_t(c) + 3 * a + bb];
                            for (int l = 0; l < 3; ++l) row[pos + l] += out[a][k].d[3 * bb + l];
                        }
                    }
                }
            } else {
                double U[3][3], out[3][3];
                for (int a = 0; a < 3; ++a)
                    for (int k = 0; k < 3; ++k) U[a][k] = x[3 * int64_t(cv[a]) + k];
                cell_residual(R, c, U, out);
                for (int a = 0; a < 3; ++a) {
                    const int64_t i = cv[a];
                    if (i < lo || i >= hi) continue;
                    for (int k = 0; k < 3; ++k) b[3 * i + k] += out[a][k];
                }
            }
        }
        for (int32_t f : R.tfacets[t]) {
            const int32_t *fv = &R.facets[3 * int64_t(f)];
            if (JAC) {
                Dual<9> U[3][3], out[3];
                for (int a = 0; a < 3; ++a)
                    for (int k = 0; k < 3; ++k) {
                        U[a][k] = cst<9>(x[3 * int64_t(fv[a]) + k]);
                        U[a][k].d[3 * a + k] = 1.0;
                    }
                facet_residual(R, fv, U, out);
                for (int a = 0; a < 2; ++a) {
                    const int64_t i = fv[a];
                    if (i < lo || i >= hi) continue;
                    if (b) b[3 * i + 1] += out[a].v;
                    double *row = &R.A.val[R.A.ptr[3 * i + 1]];
                    const int32_t *b0 = &R.ncol[R.nptr[i]], *b1 = &R.ncol[R.nptr[i + 1]];
                    for (int bb = 0; bb < 3; ++bb) {
                        const int pos = 3 * int(std::lower_bound(b0, b1, fv[bb]) - b0);
                        for (int l = 0; l < 3; ++l) row[pos + l] += out[a].d[3 * bb + l];
                    }
                }
            } else {
                double U[3][3], out[3];
                for (int a = 0; a < 3; ++a)
                    for (int k = 0; k < 3; ++k) U[a][k] = x[3 * int64_t(fv[a]) + k];
                facet_residual(R, fv, U, out);
                for (int a = 0; a < 2; ++a) {
                    const int64_t i = fv[a];
                    if (i >= lo && i < hi) b[3 * i + 1] += out[a];
                }
            }
        }
        // DirichletBC.apply: b_i = x_i - g ; matrix row -> unit row (pattern kept)
        for (int64_t i = lo; i < hi; ++i) {
            if (!R.is_bc[i]) continue;
            const int64_t r = 3 * i + 2;
            if (b) b[r] = x[r] - R.bc_val[i];
            if (JAC)
                for (int64_t p = R.A.ptr[r]; p < R.A.ptr[r + 1]; ++p) R.A.val[p] = (R.A.col[p] == r) ? 1.0 : 0.0;
        }
    }
    (JAC ? R.t_asm_jac : R.t_asm_res) += omp_get_wtime() - t0;
}

// ---- block-Jacobi / ILU(0) ----------------------------------------------------------------------------------------------
void ilu_symbolic(Ref &R) {
    R.L.assign(R.P, Csr());
    R.Ldiag.assign(R.P, {});
    R.Lsrc.assign(R.P, {});
#pragma omp parallel num_threads(R.P)
    {
        const int t = omp_get_thread_num();
        const int64_t r0 = 3 * R.lo[t], r1 = 3 * R.lo[t + 1], nb = r1 - r0;
        Csr &B = R.L[t];
        B.n = nb;
        B.ptr.assign(nb + 1, 0);
        for (int64_t r = r0; r < r1; ++r) {
            int64_t cnt = 0;
            for (int64_t p = R.A.ptr[r]; p < R.A.ptr[r + 1]; ++p) cnt += (R.A.col[p] >= r0 && R.A.col[p] < r1);
            B.ptr[r - r0 + 1] = B.ptr[r - r0] + cnt;
        }
        B.col.resize(B.ptr[nb]);
        B.val.resize(B.ptr[nb]);
        R.Lsrc[t].resize(B.ptr[nb]);
        R.Ldiag[t].resize(nb);
        for (int64_t r = r0; r < r1; ++r) {
            int64_t q = B.ptr[r - r0];
            for (int64_t p = R.A.ptr[r]; p < R.A.ptr[r + 1]; ++p) {
                const int32_t c = R.A.col[p];
                if (c < r0 || c >= r1) continue;
                if (c == r) R.Ldiag[t][r - r0] = q;
                B.col[q] = int32_t(c - r0);
                R.Lsrc[t][q] = p;
                ++q;
            }
        }
    }
}

// IKJ ILU(0) in the given ordering, in place on the block copy (same algorithm as oracle/newton.py:ilu0 = PETSc's default)
void ilu_numeric(Ref &R) {
    const double t0 = omp_get_wtime();
#pragma omp parallel num_threads(R.P)
    {
        const int t = omp_get_thread_num();
        Csr &B = R.L[t];
        const std::vector<int64_t> &dg = R.Ldiag[t], &src = R.Lsrc[t];
        for (size_t q = 0; q < B.val.size(); ++q) B.val[q] = R.A.val[src[q]];
        std::vector<int64_t> where(B.n, -1);
        for (int64_t i = 0; i < B.n; ++i) {
            const int64_t lo = B.ptr[i], hi = B.ptr[i + 1];
            for (int64_t p = lo; p < hi; ++p) where[B.col[p]] = p;
            for (int64_t kk = lo; kk < dg[i]; ++kk) {
                const int32_t k = B.col[kk];
                const double lik = B.val[kk] / B.val[dg[k]];
                B.val[kk] = lik;
                for (int64_t p = dg[k] + 1; p < B.ptr[k + 1]; ++p) {
                    const int64_t pos = where[B.col[p]];
                    if (pos >= 0) B.val[pos] -= lik * B.val[p];
                }
            }
            for (int64_t p = lo; p < hi; ++p) where[B.col[p]] = -1;
        }
    }
    R.t_ilu += omp_get_wtime() - t0;
}

// z = M^-1 r: forward / backward substitution in every block (must be called inside a parallel region by thread t)
inline void ilu_apply_block(const Ref &R, int t, const double *r, double *z) {
    const Csr &B = R.L[t];
    const std::vector<int64_t> &dg = R.Ldiag[t];
    const int64_t r0 = 3 * R.lo[t];
    double *y = z + r0;
    const double *rr = r + r0;
    for (int64_t i = 0; i < B.n; ++i) {
        double s = rr[i];
        for (int64_t p = B.ptr[i]; p < dg[i]; ++p) s -= B.val[p] * y[B.col[p]];
        y[i] = s;
    }
    for (int64_t i = B.n - 1; i >= 0; --i) {
        double s = y[i];
        for (int64_t p = dg[i] + 1; p < B.ptr[i + 1]; ++p) s -= B.val[p] * y[B.col[p]];
        y[i] = s / B.val[dg[i]];
    }
}

// ---- Krylov -------------------------------------------------------------------------------------------------------------
struct Work {
    double red[64 * 72];
};

// w = M^-1 (A v)   (two phases separated by a barrier: the block solve needs the whole product of its own rows only)
void apply_pa(Ref &R, const double *v, double *w) {
#pragma omp parallel num_threads(R.P)
    {
        const int t = omp_get_thread_num();
        const int64_t r0 = 3 * R.lo[t], r1 = 3 * R.lo[t + 1];
        double *tmp = R.tmp.data();
        for (int64_t r = r0; r < r1; ++r) {
            double s = 0.0;
            for (int64_t p = R.A.ptr[r]; p < R.A.ptr[r + 1]; ++p) s += R.A.val[p] * v[R.A.col[p]];
            tmp[r] = s;
        }
        ilu_apply_block(R, t, tmp, w);
    }
}

void apply_p(Ref &R, const double *r, double *z) {
#pragma omp parallel num_threads(R.P)
    ilu_apply_block(R, omp_get_thread_num(), r, z);
}

double dot(const Ref &R, const double *a, const double *b) {
    double part[256] = {0};
#pragma omp parallel num_threads(R.P)
    {
        const int t = omp_get_thread_num();
        double s = 0.0;
        for (int64_t i = 3 * R.lo[t]; i < 3 * R.lo[t + 1]; ++i) s += a[i] * b[i];
        part[t] = s;
    }
    double s = 0.0;
    for (int t = 0; t < R.P; ++t) s += part[t];       // fixed order: deterministic for a given P
    return s;
}

// GMRES(restart), left preconditioned, classical Gram-Schmidt, zero initial guess (PETSc KSPGMRES defaults)
int gmres(Ref &R, const double *b, double *x, int restart, double rtol, double atol, int max_it, int *converged) {
    const int64_t n = R.ndof;
    if (R.restart_alloc < restart) {
        R.V.assign(size_t(restart + 1) * n, 0.0);
        R.restart_alloc = restart;
    }
    double *V = R.V.data(), *w = R.w.data();
    std::vector<double> H(size_t(restart + 1) * restart, 0.0), cs(restart), sn(restart), g(restart + 1), h(restart + 1);
    auto Hm = [&](int i, int j) -> double & { return H[size_t(i) * restart + j]; };
    memset(x, 0, n * sizeof(double));
    apply_p(R, b, w);
    double beta = sqrt(dot(R, w, w));
    const double tol = std::max(rtol * beta, atol);
    *converged = 0;
    if (!(beta == beta)) return 0;
    if (beta <= tol) { *converged = 1; return 0; }
    int its = 0;
    std::vector<double> part(size_t(R.P) * (restart + 2));
    while (its < max_it) {
        const double ib = 1.0 / beta;
#pragma omp parallel for schedule(static) num_threads(R.P)
        for (int64_t i = 0; i < n; ++i) V[i] = w[i] * ib;
        std::fill(g.begin(), g.end(), 0.0);
        g[0] = beta;
        int k_used = 0;
        bool done = false;
        for (int k = 0; k < restart && its < max_it; ++k) {
            apply_pa(R, V + int64_t(k) * n, w);
            // h = V^T w (one pass), w -= V h, norm
#pragma omp parallel num_threads(R.P)
            {
                const int t = omp_get_thread_num();
                double *pp = &part[size_t(t) * (restart + 2)];
                for (int j = 0; j <= k; ++j) pp[j] = 0.0;
                const int64_t i0 = 3 * R.lo[t], i1 = 3 * R.lo[t + 1];
                for (int j = 0; j <= k; ++j) {
                    const double *vj = V + int64_t(j) * n;
                    double s = 0.0;
                    for (int64_t i = i0; i < i1; ++i) s += vj[i] * w[i];
                    pp[j] = s;
                }
            }
            for (int j = 0; j <= k; ++j) {
                double s = 0.0;
                for (int t = 0; t < R.P; ++t) s += part[size_t(t) * (restart + 2) + j];
                h[j] = s;
            }
            double nrm2 = 0.0;
#pragma omp parallel num_threads(R.P)
            {
                const int t = omp_get_thread_num();
                const int64_t i0 = 3 * R.lo[t], i1 = 3 * R.lo[t + 1];
                for (int j = 0; j <= k; ++j) {
                    const double *vj = V + int64_t(j) * n;
                    const double hj = h[j];
                    for (int64_t i = i0; i < i1; ++i) w[i] -= hj * vj[i];
                }
                double s = 0.0;
                for (int64_t i = i0; i < i1; ++i) s += w[i] * w[i];
                part[size_t(t) * (restart + 2)] = s;
            }
            for (int t = 0; t < R.P; ++t) nrm2 += part[size_t(t) * (restart + 2)];
            const double hk1 = sqrt(nrm2);
            if (hk1 > 0.0) {
                double *vn = V + int64_t(k + 1) * n;
                const double ih = 1.0 / hk1;
#pragma omp parallel for schedule(static) num_threads(R.P)
                for (int64_t i = 0; i < n; ++i) vn[i] = w[i] * ih;
            }
            for (int i = 0; i <= k; ++i) Hm(i, k) = h[i];
            Hm(k + 1, k) = hk1;
            for (int i = 0; i < k; ++i) {
                const double tt = cs[i] * Hm(i, k) + sn[i] * Hm(i + 1, k);
                Hm(i + 1, k) = -sn[i] * Hm(i, k) + cs[i] * Hm(i + 1, k);
                Hm(i, k) = tt;
            }
            const double d = hypot(Hm(k, k), Hm(k + 1, k));
            if (d == 0.0 || !(d == d)) return its;
            cs[k] = Hm(k, k) / d;
            sn[k] = Hm(k + 1, k) / d;
            Hm(k, k) = d;
            Hm(k + 1, k) = 0.0;
            g[k + 1] = -sn[k] * g[k];
            g[k] = cs[k] * g[k];
            ++its;
            k_used = k + 1;
            if (fabs(g[k + 1]) <= tol) { done = true; break; }
        }
        std::vector<double> y(k_used);
        for (int i = k_used - 1; i >= 0; --i) {
            double s = g[i];
            for (int j = i + 1; j < k_used; ++j) s -= Hm(i, j) * y[j];
            y[i] = s / Hm(i, i);
        }
#pragma omp parallel num_threads(R.P)
        {
            const int t = omp_get_thread_num();
            for (int j = 0; j < k_used; ++j) {
                const double *vj = V + int64_t(j) * n;
                const double yj = y[j];
                for (int64_t i = 3 * R.lo[t]; i < 3 * R.lo[t + 1]; ++i) x[i] += yj * vj[i];
            }
        }
        if (done) { *converged = 1; break; }
        if (its >= max_it) break;
        // restart: r = M^-1 (b - A x)
#pragma omp parallel num_threads(R.P)
        {
            const int t = omp_get_thread_num();
            double *tmp = R.tmp.data();
            for (int64_t r = 3 * R.lo[t]; r < 3 * R.lo[t + 1]; ++r) {
                double s = 0.0;
                for (int64_t p = R.A.ptr[r]; p < R.A.ptr[r + 1]; ++p) s += R.A.val[p] * x[R.A.col[p]];
                tmp[r] = b[r] - s;
            }
            ilu_apply_block(R, t, tmp, w);
        }
        beta = sqrt(dot(R, w, w));
        if (beta <= tol) { *converged = 1; break; }
    }
    return its;
}

}  // namespace

extern "C" {

void *sref_create(int64_t nv, const double *coords, int64_t nc, const int32_t *cells, int nthreads) {
    Ref *R = new Ref();
    R->nv = nv;
    R->nc = nc;
    R->ndof = 3 * nv;
    R->P = std::max(1, std::min(nthreads, 256));
    R->coords.assign(coords, coords + 2 * nv);
    R->cells.assign(cells, cells + 3 * nc);
    omp_set_dynamic(0);
    omp_set_num_threads(R->P);
    build_pattern(*R);
    partition(*R);
    R->tfacets.assign(R->P, {});
    R->is_bc.assign(nv, 0);
    R->bc_val.assign(nv, 0.0);
    R->uold.assign(R->ndof, 0.0);
    R->uold1.assign(R->ndof, 0.0);
    R->w.assign(R->ndof, 0.0);
    R->tmp.assign(R->ndof, 0.0);
    ilu_symbolic(*R);
    return R;
}

void sref_destroy(void *h) { delete (Ref *)h; }

int sref_threads(void *h) { return ((Ref *)h)->P; }

void sref_set_bc(void *h, int64_t n, const int32_t *nodes, const double *vals) {
    Ref *R = (Ref *)h;
    for (int64_t i = 0; i < n; ++i) { R->is_bc[nodes[i]] = 1; R->bc_val[nodes[i]] = vals[i]; }
}

// facets: [n][3] = (a, b, opposite vertex) of every exterior facet that carries the electron drift flux
void sref_set_facets(void *h, int64_t n, const int32_t *fv) {
    Ref *R = (Ref *)h;
    R->facets.assign(fv, fv + 3 * n);
    R->tfacets.assign(R->P, {});
    for (int64_t f = 0; f < n; ++f) {
        const int ta = thread_of(*R, fv[3 * f]), tb = thread_of(*R, fv[3 * f + 1]);
        R->tfacets[ta].push_back(int32_t(f));
        if (tb != ta) R->tfacets[tb].push_back(int32_t(f));
    }
}

void sref_set_history(void *h, const double *uold, const double *uold1, double dt, double dt_old) {
    Ref *R = (Ref *)h;
    memcpy(R->uold.data(), uold, R->ndof * sizeof(double));
    memcpy(R->uold1.data(), uold1, R->ndof * sizeof(double));
    R->dt = dt;
    R->dt_old = dt_old;
}

void sref_residual(void *h, const double *x, double *b) { assemble<false>(*(Ref *)h, x, b); }
void sref_jacobian(void *h, const double *x, double *b) { assemble<true>(*(Ref *)h, x, b); }

int64_t sref_nnz(void *h) { return ((Ref *)h)->A.ptr.back(); }
void sref_get_csr(void *h, int64_t *ptr, int32_t *col, double *val) {
    Ref *R = (Ref *)h;
    memcpy(ptr, R->A.ptr.data(), R->A.ptr.size() * sizeof(int64_t));
    memcpy(col, R->A.col.data(), R->A.col.size() * sizeof(int32_t));
    memcpy(val, R->A.val.data(), R->A.val.size() * sizeof(double));
}

// z = M^-1 r with the current factorisation (for the ILU(0) parity test)
void sref_factor(void *h) { ilu_numeric(*(Ref *)h); }
void sref_pc_apply(void *h, const double *r, double *z) { apply_p(*(Ref *)h, r, z); }

int sref_gmres(void *h, const double *b, double *x, int restart, double rtol, int max_it, int *converged) {
    return gmres(*(Ref *)h, b, x, restart, rtol, 1e-50, max_it, converged);
}

// SNES newtonls/basic (same stopping rules as oracle/newton.py:newton).  Returns the iteration count, -1 on failure.
int sref_newton(void *h, double *x, double rtol, double atol, double stol, int max_it, int restart, double ksp_rtol,
                int ksp_max_it, int *krylov_its, double *fnorms) {
    Ref &R = *(Ref *)h;
    std::vector<double> b(R.ndof), dx(R.ndof);
    *krylov_its = 0;
    assemble<true>(R, x, b.data());
    double fnorm0 = sqrt(dot(R, b.data(), b.data())), fnorm = fnorm0;
    if (fnorms) fnorms[0] = fnorm0;
    if (!(fnorm0 == fnorm0) || isinf(fnorm0)) return -1;
    if (fnorm0 < atol) return 0;
    bool have_jac = true;
    for (int it = 1; it <= max_it; ++it) {
        if (!have_jac) assemble<true>(R, x, nullptr);
        ilu_numeric(R);
        int conv = 0;
        const double t0 = omp_get_wtime();
        *krylov_its += gmres(R, b.data(), dx.data(), restart, ksp_rtol, 1e-50, ksp_max_it, &conv);
        R.t_krylov += omp_get_wtime() - t0;
        if (!conv) return -1;
#pragma omp parallel for schedule(static) num_threads(R.P)
        for (int64_t i = 0; i < R.ndof; ++i) x[i] -= dx[i];
        // SNES evaluates F first; the Jacobian of the next iteration is assembled together with it (one pass), as libfedmx does
        assemble<true>(R, x, b.data());
        have_jac = true;
        fnorm = sqrt(dot(R, b.data(), b.data()));
        if (fnorms && it < 64) fnorms[it] = fnorm;
        if (!(fnorm == fnorm) || isinf(fnorm)) return -1;
        if (fnorm < atol || fnorm <= rtol * fnorm0) return it;
        const double dxn = sqrt(dot(R, dx.data(), dx.data())), xn = sqrt(dot(R, x, x));
        if (dxn < stol * xn) return it;
    }
    return -1;
}

void sref_timers(void *h, double out[4], int reset) {
    Ref *R = (Ref *)h;
    out[0] = R->t_asm_jac; out[1] = R->t_asm_res; out[2] = R->t_ilu; out[3] = R->t_krylov;
    if (reset) R->t_asm_jac = R->t_asm_res = R->t_ilu = R->t_krylov = 0.0;
}

}  // extern "C"
