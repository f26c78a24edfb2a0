// Glow-discharge model (LMEA, N_eq = 5: ln w_e, ln n_Ar*, ln n_Ar+, ln n_e, Phi) for the row-assembly kernel.
//
// Reference: tests/integrated_tests/glow_discharge/fedm_gd.py:345-388 (fluxes, sources, balance equations, energy
// equation, Poisson, 'flux source' boundary terms with secondary emission) over fedm/functions.py:205-216 (Max),
// :219-237 (Flux), :350-368 (log-BDF2 balance equation), :401 (Poisson), :514-522 (Boundary_flux), :768-774
// (semi-implicit coefficients), :835-843 (Source_term), :901-912 (Energy_Source_term).  Recipe: SURVEY.md App. B.
// Quadrature: degree 4 (6-point rule) on cells, 3-point Gauss-Legendre on facets (fedm_gd.py:34).
//
// The Jacobian is obtained by first-order forward-mode AD (Dual<15>) of an explicit closed form in which the spatial
// gradients of the semi-implicit composites, e.g. grad(D_si e^u), are expanded by hand through every P1 field;
// the CPU oracle reaches the same quantities by nested AD, so the two derivations are independent.
//
// fields (per vertex, NF = 28): [0..3] BDF2 history of u0..u3, [4] eps_old (mean energy, previous step),
//   [5] ln n_e (previous step), [6] D_e, [7] dD_e/deps, [8] mu_e, [9] dmu_e/deps, [10] mu_i, [11] D_i, [12] D_*,
//   [13] eps (current mean energy, for the electron thermal velocity), [14..20] k_r, [21..27] dk_r/deps
// params: [0] N0, [1] e/eps0, [2] number of reactions R (<= 7), [3] facet-marker mask of the 'flux source' boundaries,
//   [4] v_th(Ar*), [5] v_th(Ar+), [6] ref(Ar*), [7] ref(Ar+), [8] ref(e), [9] gamma, [10] u_see,
//   [11] 16 e/(3 pi m_e), [12] 1.3333 (energy-flux thermal factor), [16+4r+j] power matrix, [48+4r+j] gain-loss,
//   [80+r] energy loss of reaction r, [88+r] 1 if reaction r is 'Umean'-dependent (semi-implicit)
#pragma once

#include "dual.cuh"

struct GlowModel {
    static constexpr bool TILED = false;
    static constexpr int MINB = 2, MINB_RES = 2;
    static constexpr unsigned ZERO_MASK = 0;
    static constexpr int NC = 0;
    static constexpr int NEQ = 5;
    static constexpr int NF = 28;
    static constexpr int ND = 15;
    typedef Dual<15> D;
    typedef Vtx<5, 28> V;

    // value of unknown k at a point with barycentric weights (l0,l1,l2); derivative index = b*5 + k
    static __device__ __forceinline__ D uval(const V &a, const V &p, const V &n, int k, double l0, double l1, double l2) {
        D r = dconst<15>(l0 * a.u[k] + l1 * p.u[k] + l2 * n.u[k]);
        r.d[k] = l0; r.d[5 + k] = l1; r.d[10 + k] = l2;
        return r;
    }
    static __device__ __forceinline__ D ugrad(const V &a, const V &p, const V &n, int k, const double g[3]) {
        D r = dconst<15>(a.u[k] * g[0] + p.u[k] * g[1] + n.u[k] * g[2]);
        r.d[k] = g[0]; r.d[5 + k] = g[1]; r.d[10 + k] = g[2];
        return r;
    }
    static __device__ __forceinline__ double fval(const V &a, const V &p, const V &n, int f, double l0, double l1, double l2) {
        return l0 * a.f[f] + l1 * p.f[f] + l2 * n.f[f];
    }
    static __device__ __forceinline__ double fgrad(const V &a, const V &p, const V &n, int f, const double g[3]) {
        return a.f[f] * g[0] + p.f[f] * g[1] + n.f[f] * g[2];
    }

    template <bool JAC>
    static __device__ __noinline__ void row(const ModelParams &mp, const V &a, const V &p, const V &n, uint32_t mk_ap,
                                            uint32_t mk_an, RowOut<5> &o) {
        const double e1r = p.r - a.r, e1z = p.z - a.z, e2r = n.r - a.r, e2z = n.z - a.z;
        const double det = e1r * e2z - e1z * e2r;
        const double inv = 1.0 / det;
        const double gr[3] = {-(e2z - e1z) * inv, e2z * inv, -e1z * inv};
        const double gz[3] = {-(e1r - e2r) * inv, -e2r * inv, e1r * inv};
        const double adet = fabs(det);
        const double N0 = mp.p[0], kappa = mp.p[1];
        const int NR = int(mp.p[2]);
        const double tr = mp.dt / mp.dt_old;
        const double ct = (1.0 + 2.0 * tr) / ((1.0 + tr) * mp.dt);
        // cell-constant gradients of the unknowns (duals) and of the coefficient fields (plain)
        D gur[5], guz[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) { gur[k] = ugrad(a, p, n, k, gr); guz[k] = ugrad(a, p, n, k, gz); }
        const D Er = -gur[4], Ez = -guz[4];
        const double g_eo_r = fgrad(a, p, n, 4, gr), g_eo_z = fgrad(a, p, n, 4, gz);      // grad eps_old
        const double g_uo_r = fgrad(a, p, n, 5, gr), g_uo_z = fgrad(a, p, n, 5, gz);      // grad ln n_e(old)
        const double g_D3_r = fgrad(a, p, n, 6, gr), g_D3_z = fgrad(a, p, n, 6, gz);
        const double g_D3d_r = fgrad(a, p, n, 7, gr), g_D3d_z = fgrad(a, p, n, 7, gz);
        const double g_D1_r = fgrad(a, p, n, 12, gr), g_D1_z = fgrad(a, p, n, 12, gz);

        D R[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) R[k] = dconst<15>(0.0);
        const double ga_r = gr[0], ga_z = gz[0];

        for (int q = 0; q < 6; ++q) {
            const double l0 = kTriDeg4[q][0], l1 = kTriDeg4[q][1], l2 = kTriDeg4[q][2];
            const double W = kTriDeg4[q][3] * adet * kTwoPi * (l0 * a.r + l1 * p.r + l2 * n.r);
            D u[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) u[k] = uval(a, p, n, k, l0, l1, l2);
            const double eo = fval(a, p, n, 4, l0, l1, l2), uo3 = fval(a, p, n, 5, l0, l1, l2);
            const double D3 = fval(a, p, n, 6, l0, l1, l2), D3d = fval(a, p, n, 7, l0, l1, l2);
            const double mu3 = fval(a, p, n, 8, l0, l1, l2), mu3d = fval(a, p, n, 9, l0, l1, l2);
            const double mu2 = fval(a, p, n, 10, l0, l1, l2), D2 = fval(a, p, n, 11, l0, l1, l2);
            const double D1 = fval(a, p, n, 12, l0, l1, l2);
            // semi-implicit mean energy: eps_e = eps_old + (e^{u0} - e^{u3} eps_old)/e^{u3_old};  delta = eps_e - eps_old
            const D q0 = dexp(u[0] - uo3), q3 = dexp(u[3] - uo3);
            const D delta = q0 - q3 * eo;
            const D D3si = delta * D3d + D3, mu3si = delta * mu3d + mu3;
            // grad(delta) = q0 (grad u0 - grad uo3) - q3 eps_old (grad u3 - grad uo3) - q3 grad eps_old
            const D gd_r = q0 * (gur[0] - g_uo_r) - (q3 * eo) * (gur[3] - g_uo_r) - q3 * g_eo_r;
            const D gd_z = q0 * (guz[0] - g_uo_z) - (q3 * eo) * (guz[3] - g_uo_z) - q3 * g_eo_z;
            const D gDs_r = delta * g_D3d_r + gd_r * D3d + g_D3_r;        // grad D3si
            const D gDs_z = delta * g_D3d_z + gd_z * D3d + g_D3_z;
            const D w = dexp(u[0]), n1 = dexp(u[1]), n2 = dexp(u[2]), n3 = dexp(u[3]);
            // fluxes
            const D ge_r = -(n3 * (gDs_r + D3si * gur[3])) - mu3si * Er * n3;      // Gamma_e = -grad(D3si e^{u3}) - mu3si E e^{u3}
            const D ge_z = -(n3 * (gDs_z + D3si * guz[3])) - mu3si * Ez * n3;
            const D gen_r = -((5.0 / 3.0) * (w * (gDs_r + D3si * gur[0]))) - (5.0 * mu3si * (1.0 / 3.0)) * Er * w;
            const D gen_z = -((5.0 / 3.0) * (w * (gDs_z + D3si * guz[0]))) - (5.0 * mu3si * (1.0 / 3.0)) * Ez * w;
            const D g2_r = -(D2 * (n2 * gur[2])) + mu2 * (Er * n2);                 // Gamma_i = -D2 e^{u2} grad u2 + mu2 E e^{u2}
            const D g2_z = -(D2 * (n2 * guz[2])) + mu2 * (Ez * n2);
            const D gD1_r = n1 * (gur[1] * D1 + g_D1_r), gD1_z = n1 * (guz[1] * D1 + g_D1_z);   // grad(D1 e^{u1})
            // reaction rates and sources
            D f1 = dconst<15>(0.0), f2 = dconst<15>(0.0), f3 = dconst<15>(0.0), fen = dconst<15>(0.0);
            for (int r = 0; r < NR; ++r) {
                const double k = fval(a, p, n, 14 + r, l0, l1, l2);
                D rate = (mp.p[88 + r] != 0.0) ? (delta * fval(a, p, n, 21 + r, l0, l1, l2) + k) : dconst<15>(k);
                for (int c = 0; c < int(mp.p[16 + 4 * r + 0]); ++c) rate = rate * N0;
                for (int c = 0; c < int(mp.p[16 + 4 * r + 1]); ++c) rate = rate * n1;
                for (int c = 0; c < int(mp.p[16 + 4 * r + 2]); ++c) rate = rate * n2;
                for (int c = 0; c < int(mp.p[16 + 4 * r + 3]); ++c) rate = rate * n3;
                f1 = f1 + rate * mp.p[48 + 4 * r + 1];
                f2 = f2 + rate * mp.p[48 + 4 * r + 2];
                f3 = f3 + rate * mp.p[48 + 4 * r + 3];
                fen = fen - rate * mp.p[80 + r];
            }
            fen = fen - (ge_r * Er + ge_z * Ez);
            // time-derivative factors e^{u}(c_t u + hist)
            const D t0 = w * (u[0] * ct + fval(a, p, n, 0, l0, l1, l2));
            const D t1 = n1 * (u[1] * ct + fval(a, p, n, 1, l0, l1, l2));
            const D t2 = n2 * (u[2] * ct + fval(a, p, n, 2, l0, l1, l2));
            const D t3 = n3 * (u[3] * ct + fval(a, p, n, 3, l0, l1, l2));
            R[0] = R[0] + ((t0 - fen) * l0 - (gen_r * ga_r + gen_z * ga_z)) * W;
            R[1] = R[1] + ((t1 - f1) * l0 + (gD1_r * ga_r + gD1_z * ga_z)) * W;
            R[2] = R[2] + ((t2 - f2) * l0 - (g2_r * ga_r + g2_z * ga_z)) * W;
            R[3] = R[3] + ((t3 - f3) * l0 - (ge_r * ga_r + ge_z * ga_z)) * W;
            R[4] = R[4] + ((gur[4] * ga_r + guz[4] * ga_z) - (n2 - n3) * (kappa * l0)) * W;
        }
        // ---- 'flux source' boundary terms on the exterior facets (a,p) and (a,n)
        const uint32_t mask = uint32_t(mp.p[3]);
        for (int side = 0; side < 2; ++side) {
            const uint32_t mk = side == 0 ? mk_ap : mk_an;
            if (mk == 0 || !((mask >> mk) & 1u)) continue;
            const V &b = side == 0 ? p : n;
            const V &c = side == 0 ? n : p;
            const double tr_ = b.r - a.r, tz_ = b.z - a.z;
            const double L = sqrt(tr_ * tr_ + tz_ * tz_);
            double nr = tz_ / L, nz = -tr_ / L;
            if (nr * (c.r - a.r) + nz * (c.z - a.z) > 0.0) { nr = -nr; nz = -nz; }
            const D En = -(gur[4] * nr + guz[4] * nz);
            const D gu2n = gur[2] * nr + guz[2] * nz;
            const double cs1 = (1.0 - mp.p[6]) / (1.0 + mp.p[6]), cs2 = (1.0 - mp.p[7]) / (1.0 + mp.p[7]);
            const double cs3 = (1.0 - mp.p[8]) / (1.0 + mp.p[8]);
            for (int q = 0; q < 3; ++q) {
                const double la = kGaussLegendre3[q][0], lb = kGaussLegendre3[q][1];
                // barycentric weights in (a,p,n) order: the facet is (a,b) with b = p (side 0) or n (side 1)
                const double l0 = la, l1 = side == 0 ? lb : 0.0, l2 = side == 0 ? 0.0 : lb;
                const double W = kGaussLegendre3[q][3] * L * kTwoPi * (l0 * a.r + l1 * p.r + l2 * n.r) * la;
                D u[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) u[k] = uval(a, p, n, k, l0, l1, l2);
                const double eo = fval(a, p, n, 4, l0, l1, l2), uo3 = fval(a, p, n, 5, l0, l1, l2);
                const double mu3 = fval(a, p, n, 8, l0, l1, l2), mu3d = fval(a, p, n, 9, l0, l1, l2);
                const double mu2 = fval(a, p, n, 10, l0, l1, l2), D2 = fval(a, p, n, 11, l0, l1, l2);
                const double eps = fval(a, p, n, 13, l0, l1, l2);
                const D delta = dexp(u[0] - uo3) - dexp(u[3] - uo3) * eo;
                const D mu3si = delta * mu3d + mu3;
                const D w = dexp(u[0]), n1 = dexp(u[1]), n2 = dexp(u[2]), n3 = dexp(u[3]);
                const D g2n = -(D2 * (n2 * gu2n)) + mu2 * (En * n2);
                const D ion = (g2n + dabs(g2n)) * 0.5;                                   // Max(Gamma_i . n, 0)
                const double vth_e = sqrt(mp.p[11] * eps);
                const D r1 = (n1 * (0.5 * mp.p[4])) * cs1;
                const D r2 = ((dabs(En * mu2) + 0.5 * mp.p[5]) * n2) * cs2;
                const D r3 = ((dabs(-(mu3si * En)) + 0.5 * vth_e) * n3) * cs3 - ion * (2.0 * mp.p[9] / (1.0 + mp.p[8]));
                const D r0 = ((dabs(-((5.0 * mu3si * (1.0 / 3.0)) * En)) + 0.5 * (mp.p[12] * vth_e)) * w) * cs3 -
                             ion * (2.0 * (mp.p[9] * mp.p[10]) / (1.0 + mp.p[8]));
                R[0] = R[0] + r0 * W;
                R[1] = R[1] + r1 * W;
                R[2] = R[2] + r2 * W;
                R[3] = R[3] + r3 * W;
            }
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            o.R[k] = R[k].v;
            if (JAC) {
#pragma unroll
                for (int bb = 0; bb < 3; ++bb)
#pragma unroll
                    for (int l = 0; l < 5; ++l) o.J[bb][k * 5 + l] = R[k].d[bb * 5 + l];
            }
        }
    }
};
