// Residual / Jacobian assembly for node-interleaved mixed P1 spaces on triangles (sm_100a).
//
// Replaces, for the FEDM form family, what FFC's generated tabulate_tensor kernels + dolfin's Assembler +
// DirichletBC::apply do behind `Problem.F` / `Problem.J` (reference: fedm/functions.py:188-202).
//
// Design (HBM-bound FP64 work, no tensor cores): one thread owns one block row (mesh node).  It walks the
// node's cell fan (consecutive cells share an edge); for every cell it evaluates row `a` of the element
// tensor -- the residual entries R_a[k] and the three n_eq x n_eq blocks J[a][a], J[a][p], J[a][n] --
// entirely in registers.  The block of an interior edge (i,j) is the sum of exactly two consecutive cells
// of the fan, so it is finished in registers ("carry") and written ONCE, in a fixed order: the scatter-add is
// deterministic by construction and every CSR value costs 8 bytes of DRAM write traffic, no atomics, no
// read-modify-write (except one L2-resident wrap-around block per interior node).
// Storage is SELL-32 with entry-major groups (fedmx_internal.h): the 32 lanes of a warp own 32 consecutive
// rows, so every store instruction writes 256 contiguous bytes.
// Dirichlet rows are fused into the epilogue: b_i = x_i - g, Jacobian row = unit row (pattern kept).
#include <algorithm>

#include "quadrature_tables.h"

#include "row_kernel.cuh"
#include "tiled_kernel.cuh"

namespace {

// =====================================================================================================
// Streamer model (LFA, N_eq = 3: ln n_i, ln n_e, Phi).
// Reference: tests/integrated_tests/streamer_discharge/fedm_streamer.py:243-279 (coefficients, fluxes, sources,
// balance equations, Poisson, Neumann drift flux on the electrodes) over fedm/functions.py:219-237, 350-368,
// 401, 523-524.  Closed form: SURVEY.md Appendix A.  Quadrature: degree 2 (3-point) on cells, 2-point
// Gauss-Legendre on facets (parameters["form_compiler"]["quadrature_degree"] = 2, fedm_streamer.py:29).
// params: [0]=mu_c [1]=mu_p [2]=D_c [3]=D_p [4]=a_c1 [5]=a_c2 [6]=a_c4 [7]=a_c3 [8]=e/eps0
//         [9]=bitmask of facet markers carrying the electron 'Neumann' drift flux
// fields: [0]=hist_i [1]=hist_e  (BDF2 history, fedmx_bdf2_history)
// =====================================================================================================
struct StreamerModel {
    static constexpr int MINB = 13, MINB_RES = 4;       // Jacobian: 3 CTAs/SM with the accumulator blocks in shared memory
    static constexpr unsigned ZERO_MASK = 1u << 3;      // d R_e / d U_i == 0: the electron equation does not see the ions
    static constexpr int NEQ = 3;
    static constexpr int NF = 2;
    static constexpr int NC = 2;      // nodal c_s = exp(u_s/6): e^{u(q)} at the 3-point rule is a product of them
    typedef Vtx<3, 2, 2> V;

    // ---- tiled (cell-batch) kernel, tiled_kernel.cuh ---------------------------------------------------------------
    // staged vertex fields: r, z, u_i, u_e, Phi, hist_i, hist_e, c_i = exp(u_i/6), c_e = exp(u_e/6)
    // cell record (residual): P.g_k [0..2], G.g_k [3..5], S [6], w0 [7];  Jacobian adds Q.g_k [8..10], g_i.g_j [11..16]
    // (0,0 1,1 2,2 0,1 0,2 1,2), 1/|E| [17], mu [18], D [19], dS/d|E| [20]      (P = -D grad u_e + mu grad Phi, Q = dP/d|E|)
    static constexpr bool TILED = true;
    static constexpr int NVF = 9, NREC_RES = 8, NREC_JAC = 21, TILED_MINB = 3;

    struct VtxRaw { double r, z, u0, u1, u2, h0, h1; };
    static __device__ __forceinline__ VtxRaw stage_load(const AsmArgs &A, int32_t node) {
        VtxRaw v;
        const double2 c = *reinterpret_cast<const double2 *>(A.coords + 2 * int64_t(node));
        v.u0 = __ldg(A.x + 3 * int64_t(node)); v.u1 = __ldg(A.x + 3 * int64_t(node) + 1); v.u2 = __ldg(A.x + 3 * int64_t(node) + 2);
        const double2 h = *reinterpret_cast<const double2 *>(A.fields + 2 * int64_t(node));
        v.r = c.x; v.z = c.y; v.h0 = h.x; v.h1 = h.y;
        return v;
    }
    static __device__ __forceinline__ void stage_store(const TileSmem &sm, int i, const VtxRaw &v) {
        sm.v(0, i) = v.r; sm.v(1, i) = v.z;
        sm.v(2, i) = v.u0; sm.v(3, i) = v.u1; sm.v(4, i) = v.u2;
        sm.v(5, i) = v.h0; sm.v(6, i) = v.h1;
        sm.v(7, i) = exp(v.u0 * (1.0 / 6.0));
        sm.v(8, i) = exp(v.u1 * (1.0 / 6.0));
    }
    static __device__ __forceinline__ void stage(const AsmArgs &A, int32_t node, const TileSmem &sm, int i) {
        stage_store(sm, i, stage_load(A, node));
    }

    // Cell record in registers: computed for two cells at a time (two independent dependency chains through sqrt / log /
    // exp keep the FP64 pipe fed), then stored.
    template <bool JAC>
    struct CellRec { double f[JAC ? NREC_JAC : NREC_RES]; };

    template <bool JAC>
    static __device__ __forceinline__ CellRec<JAC> cell_compute(const ModelParams &mp, const TileSmem &sm, int l0, int l1, int l2) {
        CellRec<JAC> R;
        const double r0 = sm.v(0, l0), z0 = sm.v(1, l0), r1 = sm.v(0, l1), z1 = sm.v(1, l1), r2 = sm.v(0, l2), z2 = sm.v(1, l2);
        const double ue0 = sm.v(3, l0), ue1 = sm.v(3, l1), ue2 = sm.v(3, l2);
        const double ph0 = sm.v(4, l0), ph1 = sm.v(4, l1), ph2 = sm.v(4, l2);
        const double e1r = r1 - r0, e1z = z1 - z0, e2r = r2 - r0, e2z = z2 - z0;
        const double det = e1r * e2z - e1z * e2r;
        const double inv = 1.0 / det;
        double gr[3], gz[3];
        gr[1] = e2z * inv; gz[1] = -e2r * inv;
        gr[2] = -e1z * inv; gz[2] = e1r * inv;
        gr[0] = -gr[1] - gr[2]; gz[0] = -gz[1] - gz[2];
        const double Gr = ph0 * gr[0] + ph1 * gr[1] + ph2 * gr[2], Gz = ph0 * gz[0] + ph1 * gz[1] + ph2 * gz[2];
        const double Hr = ue0 * gr[0] + ue1 * gr[1] + ue2 * gr[2], Hz = ue0 * gz[0] + ue1 * gz[1] + ue2 * gz[2];
        const double Em = sqrt(Gr * Gr + Gz * Gz);
        const double iEm = 1.0 / Em;
        const double lnE = log(Em);
        const double mu = mp.p[0] * exp(mp.p[1] * lnE);
        const double D = mp.p[2] * exp(mp.p[3] * lnE);
        const double iE3 = iEm * iEm * iEm;
        const double ex = exp(-mp.p[6] * iEm);
        const double pre = mp.p[4] + mp.p[5] * iE3;
        const double alpha = pre * ex - mp.p[7];
        const double Pr = -D * Hr + mu * Gr, Pz = -D * Hz + mu * Gz;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            R.f[k] = Pr * gr[k] + Pz * gz[k];
            R.f[3 + k] = Gr * gr[k] + Gz * gz[k];
        }
        R.f[6] = alpha * mu * Em;
        R.f[7] = fabs(det) * (1.0 / 6.0) * kTwoPi;
        if constexpr (JAC) {
            const double dmu = mp.p[1] * mu * iEm;
            const double dD = mp.p[3] * D * iEm;
            const double dalpha = (-3.0 * mp.p[5] * iE3 * iEm + pre * mp.p[6] * iEm * iEm) * ex;
            const double Qr = -dD * Hr + dmu * Gr, Qz = -dD * Hz + dmu * Gz;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                R.f[8 + k] = Qr * gr[k] + Qz * gz[k];
                R.f[11 + k] = gr[k] * gr[k] + gz[k] * gz[k];
            }
            R.f[14] = gr[0] * gr[1] + gz[0] * gz[1];
            R.f[15] = gr[0] * gr[2] + gz[0] * gz[2];
            R.f[16] = gr[1] * gr[2] + gz[1] * gz[2];
            R.f[17] = iEm;
            R.f[18] = mu;
            R.f[19] = D;
            R.f[20] = dalpha * mu * Em + alpha * dmu * Em + alpha * mu;
        }
        return R;
    }

    template <bool JAC>
    static __device__ __forceinline__ void cell_store(const TileSmem &sm, int c, const CellRec<JAC> &R) {
#pragma unroll
        for (int k = 0; k < (JAC ? NREC_JAC : NREC_RES); ++k) sm.r(k, c) = R.f[k];
    }

    template <bool JAC>
    static __device__ __forceinline__ void cell(const ModelParams &mp, const TileSmem &sm, int l0, int l1, int l2, int c) {
        cell_store<JAC>(sm, c, cell_compute<JAC>(mp, sm, l0, l1, l2));
    }

    static __device__ __forceinline__ void load_staged(const TileSmem &sm, int l, V &v) {
        v.id = l;
        v.r = sm.v(0, l); v.z = sm.v(1, l);
        v.u[0] = sm.v(2, l); v.u[1] = sm.v(3, l); v.u[2] = sm.v(4, l);
        v.f[0] = sm.v(5, l); v.f[1] = sm.v(6, l);
        v.c[0] = sm.v(7, l); v.c[1] = sm.v(8, l);
    }

    struct MiniParams { double p[10], dt, dt_old; };     // what row<JAC>() reads, passed by value to the boundary-row path
    static __device__ __forceinline__ MiniParams compact(const ModelParams &mp) {
        MiniParams q;
#pragma unroll
        for (int i = 0; i < 10; ++i) q.p[i] = mp.p[i];
        q.dt = mp.dt;
        q.dt_old = mp.dt_old;
        return q;
    }
    static __device__ __forceinline__ ModelParams expand(const MiniParams &q) {
        ModelParams mp;
#pragma unroll
        for (int i = 0; i < 10; ++i) mp.p[i] = q.p[i];
        mp.dt = q.dt;
        mp.dt_old = q.dt_old;
        return mp;
    }

    // per-row registers of the tiled walk: the row's own vertex and the "next" vertex of the previous cell (= the "previous"
    // vertex of the current one: consecutive cells of a fan share an edge), fields r, u_i, u_e, hist_i, hist_e, c_i, c_e
    struct RowState { double a[7], n[7]; };
    static __device__ __forceinline__ void load7(const TileSmem &sm, int l, double *d) {
        d[0] = sm.v(0, l); d[1] = sm.v(2, l); d[2] = sm.v(3, l); d[3] = sm.v(5, l); d[4] = sm.v(6, l); d[5] = sm.v(7, l); d[6] = sm.v(8, l);
    }
    static __device__ __forceinline__ void row_begin(const ModelParams &, const TileSmem &sm, int la, RowState &st) {
        load7(sm, la, st.a);
#pragma unroll
        for (int k = 0; k < 7; ++k) st.n[k] = 0.0;
    }

    static __device__ __forceinline__ int sym6(int i, int j) { return i == j ? i : i + j + 2; }

    // Row `a` of the element tensor from the staged vertices (a, p, n) and the cell record.  3-point rule: the point
    // belonging to vertex q has lambda_q = 2/3 and 1/6 elsewhere, i.e. f(q) = (f_a + f_p + f_n)/6 + f_q/2.
    template <bool JAC>
    static __device__ __forceinline__ void row_rec(const ModelParams &mp, const TileSmem &sm, RowState &st, int c, int ka, int kp, int kn,
                                                   int lp, int ln, bool first, RowOut<3> &o) {
        const double sixth = 1.0 / 6.0;
        double pv[7];
        if (first) load7(sm, lp, pv);                    // a new fan starts here (first cell, or a non-manifold vertex)
        else {
#pragma unroll
            for (int k = 0; k < 7; ++k) pv[k] = st.n[k];
        }
        load7(sm, ln, st.n);
        double r[3], ui[3], ue[3], hi[3], he[3], ci[3], ce[3];
        r[0] = st.a[0]; ui[0] = st.a[1]; ue[0] = st.a[2]; hi[0] = st.a[3]; he[0] = st.a[4]; ci[0] = st.a[5]; ce[0] = st.a[6];
        r[1] = pv[0]; ui[1] = pv[1]; ue[1] = pv[2]; hi[1] = pv[3]; he[1] = pv[4]; ci[1] = pv[5]; ce[1] = pv[6];
        r[2] = st.n[0]; ui[2] = st.n[1]; ue[2] = st.n[2]; hi[2] = st.n[3]; he[2] = st.n[4]; ci[2] = st.n[5]; ce[2] = st.n[6];
        const double Pga = sm.r(ka, c), Gga = sm.r(3 + ka, c), S = sm.r(6, c), w0 = sm.r(7, c);
        const double kappa = mp.p[8];
        const double tr = mp.dt / mp.dt_old;
        const double ct = (1.0 + 2.0 * tr) / ((1.0 + tr) * mp.dt);
        const double sr = (r[0] + r[1] + r[2]) * sixth, sui = (ui[0] + ui[1] + ui[2]) * sixth, sue = (ue[0] + ue[1] + ue[2]) * sixth;
        const double shi = (hi[0] + hi[1] + hi[2]) * sixth, she = (he[0] + he[1] + he[2]) * sixth;
        const double Ti = ci[0] * ci[1] * ci[2], Te = ce[0] * ce[1] * ce[2];
        double X0[3], X1[3], X2[3], Wne[3], Wni[3], bic[3], bec[3];
        double Wsum = 0.0, Mne = 0.0;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const double W = w0 * fma(0.5, r[q], sr);
            const double bi = fma(ct, fma(0.5, ui[q], sui), fma(0.5, hi[q], shi));
            const double be = fma(ct, fma(0.5, ue[q], sue), fma(0.5, he[q], she));
            Wni[q] = W * (Ti * (ci[q] * ci[q] * ci[q]));
            Wne[q] = W * (Te * (ce[q] * ce[q] * ce[q]));
            X0[q] = Wni[q] * bi - S * Wne[q];
            X1[q] = Wne[q] * be - S * Wne[q];
            X2[q] = Wni[q] - Wne[q];
            bic[q] = bi + ct;
            bec[q] = be + ct;
            Wsum += W;
            Mne += Wne[q];
        }
        o.R[0] = fma(0.5, X0[0], (X0[0] + X0[1] + X0[2]) * sixth);
        o.R[1] = fma(0.5, X1[0], (X1[0] + X1[1] + X1[2]) * sixth) - Mne * Pga;
        o.R[2] = Wsum * Gga - kappa * fma(0.5, X2[0], (X2[0] + X2[1] + X2[2]) * sixth);
        if (JAC) {
            const double Qga = sm.r(8 + ka, c), iEm = sm.r(17, c), mu = sm.r(18, c), D = sm.r(19, c), dS = sm.r(20, c);
            // moments sum_q T_q lambda_a(q) lambda_b(q):  t'_q = T_q lambda_a(q),  M_b = (t'_a + t'_p + t'_n)/6 + t'_b/2
            const double la6[3] = {2.0 / 3.0, sixth, sixth};
            double Mii[3], Mee[3], Mne2[3], Mni2[3], Nb[3];
            {
                double t0[3], t1[3], t2[3], t3[3];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    t0[q] = Wni[q] * bic[q] * la6[q];
                    t1[q] = (Wne[q] * bec[q] - S * Wne[q]) * la6[q];
                    t2[q] = Wne[q] * la6[q];
                    t3[q] = Wni[q] * la6[q];
                }
                const double s0 = (t0[0] + t0[1] + t0[2]) * sixth, s1 = (t1[0] + t1[1] + t1[2]) * sixth;
                const double s2 = (t2[0] + t2[1] + t2[2]) * sixth, s3 = (t3[0] + t3[1] + t3[2]) * sixth;
                const double mne6 = Mne * sixth;
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    Mii[b] = fma(0.5, t0[b], s0);
                    Mee[b] = fma(0.5, t1[b], s1);
                    Mne2[b] = fma(0.5, t2[b], s2);
                    Mni2[b] = fma(0.5, t3[b], s3);
                    Nb[b] = fma(0.5, Wne[b], mne6);
                }
            }
            const double Mne0 = Nb[0];
            const int kv[3] = {ka, kp, kn};
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                double *J = o.J[b];
                const double Ghb = sm.r(3 + kv[b], c) * iEm;           // d|E| / dPhi_b
                const double gbga = sm.r(11 + sym6(ka, kv[b]), c);
                const double j2 = -dS * Ghb * Mne0;
                J[0] = Mii[b];
                J[1] = -S * Mne2[b];
                J[2] = j2;
                J[3] = 0.0;
                J[4] = Mee[b] + Mne * D * gbga - Nb[b] * Pga;
                J[5] = j2 - Mne * (Qga * Ghb + mu * gbga);
                J[6] = -kappa * Mni2[b];
                J[7] = kappa * Mne2[b];
                J[8] = Wsum * gbga;
            }
        }
    }

    // facet (a,b): b is local vertex LB (1 = p, 2 = n), c the opposite vertex
    template <bool JAC, int LB>
    static __device__ __forceinline__ void facet(const V &a, const V &b, const V &c, const double *gr,
                                                 const double *gz, double Gr, double Gz, double iEm, double mu, double dmu,
                                                 RowOut<3> &o) {
        const double tr_ = b.r - a.r, tz_ = b.z - a.z;
        const double L = sqrt(tr_ * tr_ + tz_ * tz_);
        double nr = tz_ / L, nz = -tr_ / L;
        if (nr * (c.r - a.r) + nz * (c.z - a.z) > 0.0) { nr = -nr; nz = -nz; }
        const double Gn = Gr * nr + Gz * nz;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const double t = kGL2_t[q];
            const double la = 1.0 - t;
            const double rq = la * a.r + t * b.r;
            const double ne = exp(la * a.u[1] + t * b.u[1]);
            const double W = 0.5 * L * kTwoPi * rq * ne * la;     // weight * n_e * test function
            o.R[1] += W * mu * Gn;
            if (JAC) {
                o.J[0][4] += W * mu * Gn * la;
                o.J[LB][4] += W * mu * Gn * t;
#pragma unroll
                for (int bb = 0; bb < 3; ++bb) {
                    const double Ghb = (Gr * gr[bb] + Gz * gz[bb]) * iEm;
                    o.J[bb][5] += W * (dmu * Ghb * Gn + mu * (gr[bb] * nr + gz[bb] * nz));
                }
            }
        }
    }

    template <bool JAC>
    static __device__ __forceinline__ void row(const ModelParams &mp, const V &a, const V &p,
                                               const V &n, uint32_t mk_ap, uint32_t mk_an,
                                               RowOut<3> &o) {
        // --- geometry (fedm uses x[0] = r, x[1] = z)
        const double e1r = p.r - a.r, e1z = p.z - a.z, e2r = n.r - a.r, e2z = n.z - a.z;
        const double det = e1r * e2z - e1z * e2r;
        const double inv = 1.0 / det;
        const double gpr = e2z * inv, gpz = -e2r * inv;
        const double gnr = -e1z * inv, gnz = e1r * inv;
        const double gar = -gpr - gnr, gaz = -gpz - gnz;
        const double gr[3] = {gar, gpr, gnr}, gz[3] = {gaz, gpz, gnz};
        const double adet = fabs(det);
        // --- cell-constant fields
        const double Gr = a.u[2] * gar + p.u[2] * gpr + n.u[2] * gnr;
        const double Gz = a.u[2] * gaz + p.u[2] * gpz + n.u[2] * gnz;
        const double Hr = a.u[1] * gar + p.u[1] * gpr + n.u[1] * gnr;
        const double Hz = a.u[1] * gaz + p.u[1] * gpz + n.u[1] * gnz;
        const double Em = sqrt(Gr * Gr + Gz * Gz);
        const double iEm = 1.0 / Em;
        const double lnE = log(Em);
        const double mu = mp.p[0] * exp(mp.p[1] * lnE);
        const double D = mp.p[2] * exp(mp.p[3] * lnE);
        const double iE3 = iEm * iEm * iEm;
        const double ex = exp(-mp.p[6] * iEm);
        const double pre = mp.p[4] + mp.p[5] * iE3;
        const double alpha = pre * ex - mp.p[7];
        const double S = alpha * mu * Em;
        const double kappa = mp.p[8];
        const double tr = mp.dt / mp.dt_old;
        const double ct = (1.0 + 2.0 * tr) / ((1.0 + tr) * mp.dt);
        // drift-diffusion vector P = -D H + mu G   (Gamma = n_e P)
        const double Pr = -D * Hr + mu * Gr, Pz = -D * Hz + mu * Gz;
        const double Pga = Pr * gar + Pz * gaz;
        double dmu = 0, dD = 0, dS = 0;
        if (JAC) {
            dmu = mp.p[1] * mu * iEm;
            dD = mp.p[3] * D * iEm;
            const double dalpha = (-3.0 * mp.p[5] * iE3 * iEm + pre * mp.p[6] * iEm * iEm) * ex;
            dS = dalpha * mu * Em + alpha * dmu * Em + alpha * mu;
        }
        // --- quadrature moments: q-th point has lambda = 2/3 at local vertex q and 1/6 elsewhere
        double Ri = 0, Re = 0, Rphi_s = 0, Wsum = 0, Mne0 = 0, Mne = 0;
        double Mii[3] = {0, 0, 0}, Mne2[3] = {0, 0, 0}, Mee[3] = {0, 0, 0}, Mni2[3] = {0, 0, 0}, Nb[3] = {0, 0, 0};
        const double w0 = adet * (1.0 / 6.0) * kTwoPi;
        // e^{u(q)} with u(q) = (u_a+u_p+u_n)/6 + u_q/2  ==  (c_a c_p c_n) c_q^3,  c_v = exp(u_v/6) precomputed per node:
        // 6 exp() per cell visit become a handful of multiplications
        const double Ti = a.c[0] * p.c[0] * n.c[0], Te = a.c[1] * p.c[1] * n.c[1];
        const double ci3[3] = {a.c[0] * a.c[0] * a.c[0], p.c[0] * p.c[0] * p.c[0], n.c[0] * n.c[0] * n.c[0]};
        const double ce3[3] = {a.c[1] * a.c[1] * a.c[1], p.c[1] * p.c[1] * p.c[1], n.c[1] * n.c[1] * n.c[1]};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const double l0 = (q == 0) ? (2.0 / 3.0) : (1.0 / 6.0);
            const double l1 = (q == 1) ? (2.0 / 3.0) : (1.0 / 6.0);
            const double l2 = (q == 2) ? (2.0 / 3.0) : (1.0 / 6.0);
            const double rq = l0 * a.r + l1 * p.r + l2 * n.r;
            const double W = w0 * rq;
            const double ui = l0 * a.u[0] + l1 * p.u[0] + l2 * n.u[0];
            const double ue = l0 * a.u[1] + l1 * p.u[1] + l2 * n.u[1];
            const double hi = l0 * a.f[0] + l1 * p.f[0] + l2 * n.f[0];
            const double he = l0 * a.f[1] + l1 * p.f[1] + l2 * n.f[1];
            const double ni = Ti * ci3[q], ne = Te * ce3[q];
            const double bi = ct * ui + hi, be = ct * ue + he;
            const double Wni = W * ni, Wne = W * ne;
            Ri += (Wni * bi - S * Wne) * l0;
            Re += (Wne * be - S * Wne) * l0;
            Rphi_s += (Wni - Wne) * l0;
            Wsum += W;
            Mne += Wne;
            if (JAC) {
                Mne0 += Wne * l0;
                const double lb[3] = {l0, l1, l2};
                const double ti = Wni * (bi + ct) * l0, te = (Wne * (be + ct) - S * Wne) * l0;
                const double tn = Wne * l0, tni = Wni * l0;
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    Mii[b] += ti * lb[b];
                    Mee[b] += te * lb[b];
                    Mne2[b] += tn * lb[b];
                    Mni2[b] += tni * lb[b];
                    Nb[b] += Wne * lb[b];
                }
            }
        }
        o.R[0] = Ri;
        o.R[1] = Re - Mne * Pga;
        o.R[2] = Wsum * (Gr * gar + Gz * gaz) - kappa * Rphi_s;
        if (JAC) {
            const double Qr = -dD * Hr + dmu * Gr, Qz = -dD * Hz + dmu * Gz;   // dP/dEm
            const double Qga = Qr * gar + Qz * gaz;
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                double *J = o.J[b];
                const double Ghb = (Gr * gr[b] + Gz * gz[b]) * iEm;   // dEm/dPhi_b
                const double gbga = gr[b] * gar + gz[b] * gaz;
                J[0] = Mii[b];                       // d R_i / d U_i
                J[1] = -S * Mne2[b];                 // d R_i / d U_e
                J[2] = -dS * Ghb * Mne0;             // d R_i / d Phi
                J[3] = 0.0;                          // d R_e / d U_i (stored: dolfin keeps the block)
                J[4] = Mee[b];
                J[5] = -dS * Ghb * Mne0 - Mne * (Qga * Ghb + mu * gbga);
                J[6] = -kappa * Mni2[b];
                J[7] = kappa * Mne2[b];
                J[8] = Wsum * gbga;
                // diffusion and drift parts of d R_e / d U_e :  + n_e D g_b.g_a  - n_e lambda_b P.g_a
                J[4] += Mne * D * gbga - Nb[b] * Pga;
            }
        }
        // --- exterior facets (a,p) and (a,n): 'Neumann' drift flux  2 pi r (-mu)(E.n) n_e v,  E = -G
        const uint32_t mask = uint32_t(mp.p[9]);
        if (mk_ap != 0 && ((mask >> mk_ap) & 1u))
            facet<JAC, 1>(a, p, n, gr, gz, Gr, Gz, iEm, mu, dmu, o);
        if (mk_an != 0 && ((mask >> mk_an) & 1u))
            facet<JAC, 2>(a, n, p, gr, gz, Gr, Gz, iEm, mu, dmu, o);
    }
};

// =====================================================================================================
// Poisson model (N_eq = 1): 2 pi r (grad u . grad v - f v) dx with f = kappa (exp(f0) - exp(f1)), f0/f1 P1
// log-density fields.  Reference: weak_form_Poisson_equation (fedm/functions.py:379-401) as used for the
// initial potential, fedm_streamer.py:210-223.  params: [0] = kappa = e/eps0.  Degree-2 quadrature.
// =====================================================================================================
struct PoissonModel {
    static constexpr bool TILED = false;
    static constexpr int MINB = 2, MINB_RES = 2;
    static constexpr unsigned ZERO_MASK = 0;
    static constexpr int NC = 0;
    static constexpr int NEQ = 1;
    static constexpr int NF = 2;

    template <bool JAC>
    static __device__ __forceinline__ void row(const ModelParams &mp, const Vtx<1, 2> &a, const Vtx<1, 2> &p,
                                               const Vtx<1, 2> &n, uint32_t, uint32_t, RowOut<1> &o) {
        const double e1r = p.r - a.r, e1z = p.z - a.z, e2r = n.r - a.r, e2z = n.z - a.z;
        const double det = e1r * e2z - e1z * e2r;
        const double inv = 1.0 / det;
        const double gpr = e2z * inv, gpz = -e2r * inv;
        const double gnr = -e1z * inv, gnz = e1r * inv;
        const double gar = -gpr - gnr, gaz = -gpz - gnz;
        const double Gr = a.u[0] * gar + p.u[0] * gpr + n.u[0] * gnr;
        const double Gz = a.u[0] * gaz + p.u[0] * gpz + n.u[0] * gnz;
        const double w0 = fabs(det) * (1.0 / 6.0) * kTwoPi;
        double Wsum = 0, src = 0;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const double l0 = (q == 0) ? (2.0 / 3.0) : (1.0 / 6.0);
            const double l1 = (q == 1) ? (2.0 / 3.0) : (1.0 / 6.0);
            const double l2 = (q == 2) ? (2.0 / 3.0) : (1.0 / 6.0);
            const double W = w0 * (l0 * a.r + l1 * p.r + l2 * n.r);
            const double f0 = l0 * a.f[0] + l1 * p.f[0] + l2 * n.f[0];
            const double f1 = l0 * a.f[1] + l1 * p.f[1] + l2 * n.f[1];
            Wsum += W;
            src += W * (exp(f0) - exp(f1)) * l0;
        }
        o.R[0] = Wsum * (Gr * gar + Gz * gaz) - mp.p[0] * src;
        if (JAC) {
            o.J[0][0] = Wsum * (gar * gar + gaz * gaz);
            o.J[1][0] = Wsum * (gpr * gar + gpz * gaz);
            o.J[2][0] = Wsum * (gnr * gar + gnz * gaz);
        }
    }
};

// =====================================================================================================
// Time-of-flight model (N_eq = 1: ln n_e), drift-diffusion-reaction with constant D, w and an analytic source.
// Reference: tests/integrated_tests/time_of_flight/fedm_tof.py:116-127 (u_analytical, Gamma = -grad(D e^u) + w e^u,
// f declared degree=2 => FFC evaluates its P2 nodal interpolant: 3 vertices + 3 edge midpoints) over
// fedm/functions.py:350-368.  Quadrature: no global degree is set, UFL estimates 8 => collapsed Gauss-Jacobi
// m = 5, 25 points (pinned by the golden VTU).
// params: [0]=D [1]=w_z [2]=alpha ;  mp.t = time of the new level ;  fields: [0]=BDF2 history
// =====================================================================================================
struct TofModel {
    static constexpr bool TILED = false;
    static constexpr int MINB = 2, MINB_RES = 2;
    static constexpr unsigned ZERO_MASK = 0;
    static constexpr int NC = 0;
    static constexpr int NEQ = 1;
    static constexpr int NF = 1;

    static __device__ __forceinline__ double source(const ModelParams &mp, double r, double z) {
        const double D = mp.p[0], w = mp.p[1], al = mp.p[2], t = mp.t;
        const double dz = z - w * t;
        const double pi = 3.14159265358979323846;
        return exp(-(dz * dz + r * r) / (4.0 * D * t) + al * w * t) * (w * al) / (8.0 * (pi * sqrt(pi)) * ((D * t) * sqrt(D * t)));
    }

    template <bool JAC>
    static __device__ __forceinline__ void row(const ModelParams &mp, const Vtx<1, 1> &a, const Vtx<1, 1> &p,
                                               const Vtx<1, 1> &n, uint32_t, uint32_t, RowOut<1> &o) {
        const double e1r = p.r - a.r, e1z = p.z - a.z, e2r = n.r - a.r, e2z = n.z - a.z;
        const double det = e1r * e2z - e1z * e2r;
        const double inv = 1.0 / det;
        const double gpr = e2z * inv, gpz = -e2r * inv;
        const double gnr = -e1z * inv, gnz = e1r * inv;
        const double gar = -gpr - gnr, gaz = -gpz - gnz;
        const double gr[3] = {gar, gpr, gnr}, gz[3] = {gaz, gpz, gnz};
        const double adet = fabs(det);
        const double Hr = a.u[0] * gar + p.u[0] * gpr + n.u[0] * gnr;
        const double Hz = a.u[0] * gaz + p.u[0] * gpz + n.u[0] * gnz;
        const double D = mp.p[0], wz = mp.p[1];
        const double tr = mp.dt / mp.dt_old;
        const double ct = (1.0 + 2.0 * tr) / ((1.0 + tr) * mp.dt);
        // P2 nodal values of the source: vertices (a,p,n) and midpoints opposite to a, p, n
        const double fa = source(mp, a.r, a.z), fp = source(mp, p.r, p.z), fn = source(mp, n.r, n.z);
        const double fma_ = source(mp, 0.5 * (p.r + n.r), 0.5 * (p.z + n.z));
        const double fmp = source(mp, 0.5 * (a.r + n.r), 0.5 * (a.z + n.z));
        const double fmn = source(mp, 0.5 * (a.r + p.r), 0.5 * (a.z + p.z));
        // Gamma . g_a = e^u ( -D H + w ) . g_a
        const double Pga = (-D * Hr) * gar + (-D * Hz + wz) * gaz;
        // The collapsed rule is not symmetric: FFC places it on the reference cell whose vertices are the cell's
        // vertices in ascending index order, so map (a,p,n) to their ranks in that order.
        const int ra = (a.id > p.id) + (a.id > n.id), rp = (p.id > a.id) + (p.id > n.id), rn = 3 - ra - rp;
        double R = 0.0, J0 = 0.0, J1 = 0.0, J2 = 0.0;
        for (int q = 0; q < 25; ++q) {
            const double l0 = kTriGJ5[q][ra], l1 = kTriGJ5[q][rp], l2 = kTriGJ5[q][rn];
            const double W = kTriGJ5[q][3] * adet * kTwoPi * (l0 * a.r + l1 * p.r + l2 * n.r);
            const double u = l0 * a.u[0] + l1 * p.u[0] + l2 * n.u[0];
            const double h = l0 * a.f[0] + l1 * p.f[0] + l2 * n.f[0];
            const double ne = exp(u);
            const double beta = ct * u + h;
            const double f = fa * l0 * (2.0 * l0 - 1.0) + fp * l1 * (2.0 * l1 - 1.0) + fn * l2 * (2.0 * l2 - 1.0) +
                             4.0 * (fma_ * l1 * l2 + fmp * l0 * l2 + fmn * l0 * l1);
            R += W * ((ne * beta - f) * l0 - ne * Pga);
            if (JAC) {
                const double c0 = W * ne * ((beta + ct) * l0 - Pga);     // multiplies lambda_b
                const double c1 = W * ne * D;                             // multiplies g_b . g_a
                J0 += c0 * l0 + c1 * (gr[0] * gar + gz[0] * gaz);
                J1 += c0 * l1 + c1 * (gr[1] * gar + gz[1] * gaz);
                J2 += c0 * l2 + c1 * (gr[2] * gar + gz[2] * gaz);
            }
        }
        o.R[0] = R;
        if (JAC) { o.J[0][0] = J0; o.J[1][0] = J1; o.J[2][0] = J2; }
    }
};

// =====================================================================================================
// L2 projection onto P1 (N_eq = 1): find x with (x, v) = (g, v), Cartesian measure, degree-4 (6-point) rule.
// Reference: dolfin project(): n_num = project(exp(u_new), V) (fedm_tof.py:165) and the reduced field
// project(1e21*sqrt(dot(-grad(Phi),-grad(Phi)))/N0) (fedm_gd.py:309, 432).  Written as the linear residual
// R_a = sum_q w |detJ| (x - g) lambda_a so that the common Newton/GMRES path solves it in one iteration.
// params: [0] = mode (0: g = exp(field0), 1: g = scale*|grad field0|)  [1] = scale ;  fields: [0] = P1 input
// =====================================================================================================
struct ProjectionModel {
    static constexpr bool TILED = false;
    static constexpr int MINB = 2, MINB_RES = 2;
    static constexpr unsigned ZERO_MASK = 0;
    static constexpr int NC = 0;
    static constexpr int NEQ = 1;
    static constexpr int NF = 1;

    template <bool JAC>
    static __device__ __forceinline__ void row(const ModelParams &mp, const Vtx<1, 1> &a, const Vtx<1, 1> &p,
                                               const Vtx<1, 1> &n, uint32_t, uint32_t, RowOut<1> &o) {
        const double e1r = p.r - a.r, e1z = p.z - a.z, e2r = n.r - a.r, e2z = n.z - a.z;
        const double det = e1r * e2z - e1z * e2r;
        const double adet = fabs(det);
        const int mode = int(mp.p[0]);
        double gconst = 0.0;
        if (mode == 1) {
            const double inv = 1.0 / det;
            const double gpr = e2z * inv, gpz = -e2r * inv, gnr = -e1z * inv, gnz = e1r * inv;
            const double gar = -gpr - gnr, gaz = -gpz - gnz;
            const double Gr = a.f[0] * gar + p.f[0] * gpr + n.f[0] * gnr;
            const double Gz = a.f[0] * gaz + p.f[0] * gpz + n.f[0] * gnz;
            gconst = mp.p[1] * sqrt(Gr * Gr + Gz * Gz);
        }
        double R = 0.0, J0 = 0.0, J1 = 0.0, J2 = 0.0;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const double l0 = kTriDeg4[q][0], l1 = kTriDeg4[q][1], l2 = kTriDeg4[q][2];
            const double W = kTriDeg4[q][3] * adet;
            const double x = l0 * a.u[0] + l1 * p.u[0] + l2 * n.u[0];
            const double g = mode == 1 ? gconst : exp(l0 * a.f[0] + l1 * p.f[0] + l2 * n.f[0]);
            R += W * (x - g) * l0;
            if (JAC) { J0 += W * l0 * l0; J1 += W * l0 * l1; J2 += W * l0 * l2; }
        }
        o.R[0] = R;
        if (JAC) { o.J[0][0] = J0; o.J[1][0] = J1; o.J[2][0] = J2; }
    }
};

#include "glow_model.cuh"

__global__ void bdf2_history_kernel(int64_t n_nodes, int n_eq, int n_hist, const double *__restrict__ x_old,
                                    const double *__restrict__ x_old1, double c_old, double c_old1,
                                    double *__restrict__ fields, int n_fields, int field_offset) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_nodes * n_hist) return;
    const int64_t v = i / n_hist;
    const int k = int(i % n_hist);
    fields[v * n_fields + field_offset + k] = c_old * x_old[v * n_eq + k] + c_old1 * x_old1[v * n_eq + k];
}

// c_s = exp(u_s / 6) for the first NC components of every local vertex (StreamerModel)
__global__ void nodal_exp6_kernel(const double *__restrict__ x, int n_eq, int nc, int64_t n_nodes, double *__restrict__ out) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_nodes * nc) return;
    const int64_t v = i / nc;
    const int k = int(i % nc);
    out[i] = exp(x[v * n_eq + k] * (1.0 / 6.0));
}

template <class Model>
int launch_assembly(fedmx_ctx *ctx, const double *d_x, double *d_b, bool jac) {
    AsmArgs A{};
    // cell-batch (tiled) kernel for the models that provide cell records; FEDMX_ASM_TILED=0 forces the generic row kernel
    static const bool env_tiled = !(getenv("FEDMX_ASM_TILED") && atoi(getenv("FEDMX_ASM_TILED")) == 0);
    if constexpr (Model::TILED) {
        if (env_tiled && ctx->tiles.n_tiles > 0) {
            A.nodal = nullptr;
            for (int k = 0; k < FEDMX_MAX_REFS; ++k) A.refs[k] = ctx->plugin_refs[k];
            A.S = ctx->sell;
            A.coords = ctx->d_coords;
            A.x = d_x;
            A.fields = ctx->d_fields;
            A.val = jac ? ctx->d_val : nullptr;
            A.res = d_b;
            A.mp = ctx->mp;
            ProfScope prof(ctx, jac ? FEDMX_PROF_ASM_JAC : FEDMX_PROF_ASM_RES);
            if (launch_tiled<Model>(A, ctx->tiles, jac, ctx->stream)) {
                ctx->launches++;
                FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
                return FEDMX_OK;
            }
        }
    }
    if (Model::NC > 0) {
        if (!ctx->d_nodal) {
            int rc = fedmx_dev_alloc(ctx, &ctx->d_nodal, size_t(ctx->n_vertices) * Model::NC);
            if (rc) return rc;
        }
        const int64_t cnt = ctx->n_vertices * Model::NC;
        nodal_exp6_kernel<<<unsigned((cnt + 255) / 256), 256, 0, ctx->stream>>>(d_x, Model::NEQ, Model::NC, ctx->n_vertices, ctx->d_nodal);
        ctx->launches++;
    }
    A.nodal = ctx->d_nodal;
    for (int k = 0; k < FEDMX_MAX_REFS; ++k) A.refs[k] = ctx->plugin_refs[k];
    A.S = ctx->sell;
    A.coords = ctx->d_coords;
    A.x = d_x;
    A.fields = ctx->d_fields;
    A.val = jac ? ctx->d_val : nullptr;
    A.res = d_b;
    A.mp = ctx->mp;
    ProfScope prof(ctx, jac ? FEDMX_PROF_ASM_JAC : FEDMX_PROF_ASM_RES);
    // resident CTAs per SM the kernel is compiled for (register cap 65536/(128*MINB)); 1x = the same with the diag/carry
    // accumulators in shared memory (row_kernel.cuh).  FEDMX_ASM_MINB overrides for tuning.  Measured on B200 at 10 M DOF
    // (tools/bench_assembly.py, incl. nodal exp + block-Jacobi update): Jacobian 1.10 / 1.14 ms for 2 / 3 CTAs with register
    // accumulators (3 CTAs: 464 B of spills), 0.97 / 1.30 ms for 3 / 4 CTAs with shared-memory accumulators (136 B / 408 B of
    // spills); residual 0.77 / 0.58 / 0.56 ms for 2 / 3 / 4 CTAs.  All variants produce bit-identical matrices.
    static const int env_minb = getenv("FEDMX_ASM_MINB") ? atoi(getenv("FEDMX_ASM_MINB")) : 0;
    const int minb = env_minb ? env_minb : (jac ? Model::MINB : Model::MINB_RES);
    launch_rows<Model>(A, jac, minb, ctx->stream);
    ctx->launches++;
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    return FEDMX_OK;
}

int assemble_dispatch(fedmx_ctx *ctx, const double *d_x, double dt, double dt_old, double t,
                      const double *h_bc_values, int n_bc_values, double *d_b, bool jac) {
    FEDMX_REQUIRE(ctx, ctx->n_vertices > 0, "assemble: no mesh");
    FEDMX_REQUIRE(ctx, ctx->mp.model_id != 0, "assemble: no model");
    FEDMX_REQUIRE(ctx, d_x != nullptr, "assemble: null state");
    FEDMX_REQUIRE(ctx, n_bc_values >= 0 && n_bc_values < FEDMX_MAX_BC, "assemble: too many Dirichlet values");
    FEDMX_REQUIRE(ctx, ctx->mp.n_fields == 0 || ctx->d_fields || ctx->mp.model_id == FEDMX_MODEL_PLUGIN, "assemble: coefficient fields not bound");
    ctx->mp.dt = dt;
    ctx->mp.dt_old = dt_old;
    ctx->mp.t = t;
    ctx->mp.bc_comp = ctx->bc_comp;
    for (int i = 0; i < n_bc_values; ++i) ctx->mp.bc_values[i] = h_bc_values[i];
    switch (ctx->mp.model_id) {
        case FEDMX_MODEL_STREAMER:
            return launch_assembly<StreamerModel>(ctx, d_x, d_b, jac);
        case FEDMX_MODEL_POISSON:
            return launch_assembly<PoissonModel>(ctx, d_x, d_b, jac);
        case FEDMX_MODEL_TOF:
            return launch_assembly<TofModel>(ctx, d_x, d_b, jac);
        case FEDMX_MODEL_GLOW:
            return launch_assembly<GlowModel>(ctx, d_x, d_b, jac);
        case FEDMX_MODEL_PROJECTION:
            return launch_assembly<ProjectionModel>(ctx, d_x, d_b, jac);
        case FEDMX_MODEL_PLUGIN: {
            FEDMX_REQUIRE(ctx, ctx->plugin_launch, "assemble: no plug-in loaded");
            AsmArgs A{};
            A.S = ctx->sell;
            A.coords = ctx->d_coords;
            A.x = d_x;
            A.fields = nullptr;
            A.nodal = nullptr;
            for (int k = 0; k < FEDMX_MAX_REFS; ++k) A.refs[k] = ctx->plugin_refs[k];
            A.val = jac ? ctx->d_val : nullptr;
            A.res = d_b;
            A.mp = ctx->mp;
            // Cell-wise path (row_kernel.cuh): the plug-in evaluates every cell ONCE into a scratch record and the rows
            // gather; needs cells * doubles-per-cell of scratch (5.1 GB for n_eq = 3 at 6.7 M cells), allocated on first
            // use.  FEDMX_JIT_CELLS=0 or a refused allocation (cap FEDMX_CELL_SCRATCH_MAX_GB, default 32) = row-wise path.
            static const bool env_cells = !(getenv("FEDMX_JIT_CELLS") && atoi(getenv("FEDMX_JIT_CELLS")) == 0);
            if (env_cells && ctx->plugin_cell_doubles && ctx->n_cells > 0 && !ctx->cell_scratch_refused) {
                const size_t need = size_t((ctx->n_cells + FEDMX_SLICE - 1) / FEDMX_SLICE * FEDMX_SLICE) * size_t(std::max(ctx->plugin_cell_doubles(1), ctx->plugin_cell_doubles(0)));
                if (need > 0 && need > ctx->cell_scratch_doubles) {
                    static const double cap_gb = getenv("FEDMX_CELL_SCRATCH_MAX_GB") ? atof(getenv("FEDMX_CELL_SCRATCH_MAX_GB")) : 32.0;
                    size_t free_b = 0, total_b = 0;
                    cudaMemGetInfo(&free_b, &total_b);
                    void *buf = nullptr;
                    if (ctx->d_cell_scratch) {                      // another plug-in with larger records was loaded
                        auto &od = ctx->owned_device;
                        od.erase(std::remove(od.begin(), od.end(), (void *)ctx->d_cell_scratch), od.end());
                        cudaStreamSynchronize(ctx->stream);
                        cudaFree(ctx->d_cell_scratch);
                        ctx->d_cell_scratch = nullptr;
                        ctx->cell_scratch_doubles = 0;
                    }
                    if (double(need) * 8.0 <= cap_gb * 1e9 && need * 8 + (size_t(2) << 30) <= free_b &&
                        cudaMalloc(&buf, need * 8) == cudaSuccess) {
                        ctx->owned_device.push_back(buf);
                        ctx->d_cell_scratch = (double *)buf;
                        ctx->cell_scratch_doubles = need;
                    } else {
                        cudaGetLastError();
                        ctx->cell_scratch_refused = true;
                    }
                }
                if (need > 0 && need <= ctx->cell_scratch_doubles) {
                    A.cells = ctx->d_cells;
                    A.n_cells = ctx->n_cells;
                    A.scratch = ctx->d_cell_scratch;
                }
            }
            ProfScope prof(ctx, jac ? FEDMX_PROF_ASM_JAC : FEDMX_PROF_ASM_RES);
            int rc = ctx->plugin_launch(&A, jac ? 1 : 0, 2, (void *)ctx->stream);
            ctx->launches += A.scratch ? 2 : 1;
            if (rc) { ctx->err = "assemble: plug-in launch failed"; return FEDMX_ECUDA; }
            FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
            return FEDMX_OK;
        }
        default:
            ctx->err = "assemble: model not implemented";
            return FEDMX_EUNSUPPORTED;
    }
}

}  // namespace

int fedmx_assemble_internal(fedmx_ctx *ctx, const double *d_x, double dt, double dt_old, double t,
                            const double *h_bc_values, int n_bc_values, double *d_b, bool jac) {
    return assemble_dispatch(ctx, d_x, dt, dt_old, t, h_bc_values, n_bc_values, d_b, jac);
}

extern "C" int fedmx_model_set(fedmx_ctx *ctx, int model_id, int n_params, const double *h_params) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, n_params >= 0 && n_params <= FEDMX_MAX_PARAMS, "fedmx_model_set: too many parameters");
    int neq = 0, nf = 0;
    unsigned zmask = 0;
    if (model_id == FEDMX_MODEL_PLUGIN) {
        ctx->zero_mask = 0;
        FEDMX_REQUIRE(ctx, ctx->plugin_launch, "fedmx_model_set: load the plug-in first (fedmx_plugin_load)");
        for (int i = 0; i < FEDMX_MAX_PARAMS; ++i) ctx->mp.p[i] = i < n_params ? h_params[i] : 0.0;
        return FEDMX_OK;
    }
    switch (model_id) {
        case FEDMX_MODEL_STREAMER: neq = StreamerModel::NEQ; nf = StreamerModel::NF; zmask = StreamerModel::ZERO_MASK; break;
        case FEDMX_MODEL_POISSON: neq = PoissonModel::NEQ; nf = PoissonModel::NF; break;
        case FEDMX_MODEL_TOF: neq = TofModel::NEQ; nf = TofModel::NF; break;
        case FEDMX_MODEL_GLOW: neq = GlowModel::NEQ; nf = GlowModel::NF; break;
        case FEDMX_MODEL_PROJECTION: neq = ProjectionModel::NEQ; nf = ProjectionModel::NF; break;
        default:
            ctx->err = "fedmx_model_set: unknown model id";
            return FEDMX_EUNSUPPORTED;
    }
    FEDMX_REQUIRE(ctx, ctx->n_eq == 0 || ctx->n_eq == neq, "fedmx_model_set: n_eq of mesh and model differ");
    ctx->mp.model_id = model_id;
    ctx->mp.n_eq = neq;
    ctx->mp.n_fields = nf;
    ctx->zero_mask = zmask;
    for (int i = 0; i < FEDMX_MAX_PARAMS; ++i) ctx->mp.p[i] = i < n_params ? h_params[i] : 0.0;
    return FEDMX_OK;
}

extern "C" int fedmx_ident_zeros_set(fedmx_ctx *ctx, int on) {
    if (!ctx) return FEDMX_EINVAL;
    ctx->mp.ident_zeros = on != 0;
    return FEDMX_OK;
}

extern "C" int fedmx_fields_bind(fedmx_ctx *ctx, int n_fields, const double *d_fields) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, ctx->mp.model_id != 0, "fedmx_fields_bind: set the model first");
    FEDMX_REQUIRE(ctx, n_fields == ctx->mp.n_fields, "fedmx_fields_bind: wrong number of fields for this model");
    ctx->d_fields = d_fields;
    return FEDMX_OK;
}

extern "C" int fedmx_assemble_residual(fedmx_ctx *ctx, const double *d_x, double dt, double dt_old, double t,
                                       const double *h_bc_values, int n_bc_values, double *d_b) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, d_b != nullptr, "fedmx_assemble_residual: null output");
    return assemble_dispatch(ctx, d_x, dt, dt_old, t, h_bc_values, n_bc_values, d_b, false);
}

extern "C" int fedmx_assemble_jacobian(fedmx_ctx *ctx, const double *d_x, double dt, double dt_old, double t,
                                       const double *h_bc_values, int n_bc_values, double *d_b) {
    if (!ctx) return FEDMX_EINVAL;
    int rc = assemble_dispatch(ctx, d_x, dt, dt_old, t, h_bc_values, n_bc_values, d_b, true);
    if (rc) return rc;
    return fedmx_update_block_jacobi(ctx);
}

extern "C" int fedmx_bdf2_history(fedmx_ctx *ctx, int n_hist, const double *d_x_old, const double *d_x_old1,
                                  double dt, double dt_old, double *d_fields, int n_fields, int field_offset) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, ctx->n_vertices > 0, "fedmx_bdf2_history: no mesh");
    FEDMX_REQUIRE(ctx, n_hist > 0 && n_hist <= ctx->n_eq && field_offset >= 0 && field_offset + n_hist <= n_fields,
                  "fedmx_bdf2_history: bad field layout");
    const double tr = dt / dt_old;
    const double trp1 = 1.0 + tr;
    const double c_old = -(trp1 * trp1) / (trp1 * dt);
    const double c_old1 = (tr * tr) / (trp1 * dt);
    const int64_t n = ctx->n_vertices * n_hist;
    const int threads = 256;
    bdf2_history_kernel<<<unsigned((n + threads - 1) / threads), threads, 0, ctx->stream>>>(
        ctx->n_vertices, ctx->n_eq, n_hist, d_x_old, d_x_old1, c_old, c_old1, d_fields, n_fields, field_offset);
    ctx->launches++;
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    return FEDMX_OK;
}

// plain-vector form of the same history term (used by the 1-D P2 path, fedmx_p2line.cu)
int fedmx_bdf2_history_raw(fedmx_ctx *ctx, int64_t n, const double *d_x_old, const double *d_x_old1, double dt, double dt_old,
                           double *d_out) {
    const double tr = dt / dt_old;
    const double trp1 = 1.0 + tr;
    const int threads = 256;
    bdf2_history_kernel<<<unsigned((n + threads - 1) / threads), threads, 0, ctx->stream>>>(
        n, 1, 1, d_x_old, d_x_old1, -(trp1 * trp1) / (trp1 * dt), (tr * tr) / (trp1 * dt), d_out, 1, 0);
    ctx->launches++;
    FEDMX_CHECK_CUDA(ctx, cudaGetLastError());
    return FEDMX_OK;
}

#include <dlfcn.h>

// Load a row kernel generated by the form compiler (fedm_b200/formcompiler.py).  The .so exports
//   int fedmx_plugin_launch(const void *asm_args, int jac, int minb, void *stream);  int fedmx_plugin_neq(); int fedmx_plugin_nf();
extern "C" int fedmx_plugin_load(fedmx_ctx *ctx, const char *path) {
    if (!ctx || !path) return FEDMX_EINVAL;
    void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) { ctx->err = std::string("fedmx_plugin_load: ") + dlerror(); return FEDMX_EINVAL; }
    auto launch = (int (*)(const void *, int, int, void *))dlsym(h, "fedmx_plugin_launch");
    auto neq = (int (*)())dlsym(h, "fedmx_plugin_neq");
    auto nf = (int (*)())dlsym(h, "fedmx_plugin_nf");
    auto sz = (int (*)())dlsym(h, "fedmx_plugin_args_size");
    FEDMX_REQUIRE(ctx, launch && neq && nf && sz, "fedmx_plugin_load: the library does not export the plug-in entry points");
    FEDMX_REQUIRE(ctx, sz() == int(sizeof(AsmArgs)), "fedmx_plugin_load: plug-in was built against a different libfedmx (AsmArgs size)");
    FEDMX_REQUIRE(ctx, ctx->n_eq == 0 || ctx->n_eq == neq(), "fedmx_plugin_load: n_eq of mesh and plug-in differ");
    FEDMX_REQUIRE(ctx, nf() <= FEDMX_MAX_REFS, "fedmx_plugin_load: too many coefficient fields");
    ctx->plugin_handle = h;
    ctx->plugin_launch = launch;
    ctx->plugin_cell_doubles = (int (*)(int))dlsym(h, "fedmx_plugin_cell_doubles");      // optional entry point
    ctx->mp.model_id = FEDMX_MODEL_PLUGIN;
    ctx->mp.n_eq = neq();
    ctx->mp.n_fields = nf();
    return FEDMX_OK;
}

extern "C" int fedmx_plugin_set_field(fedmx_ctx *ctx, int k, const double *d_ptr, int stride, int offset) {
    if (!ctx) return FEDMX_EINVAL;
    FEDMX_REQUIRE(ctx, k >= 0 && k < FEDMX_MAX_REFS && d_ptr && stride >= 1 && offset >= 0 && offset < stride, "fedmx_plugin_set_field: bad arguments");
    ctx->plugin_refs[k] = FieldRef{d_ptr, stride, offset};
    return FEDMX_OK;
}
