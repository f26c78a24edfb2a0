// Host build of a GENERATED row model (tests/test_formcompiler_cpu.py; stub headers in tools/jitcheck/stub): the AD Jacobian of
// JitModel::row<true> against central differences of row<false>, and the residual of both instantiations.
#include FORM
#include <algorithm>
#include <random>
int main() {
    constexpr int NEQ = JitModel::NEQ, NF = JitModel::NF;
    typedef JitModel::V V;
    std::mt19937 g(3); std::uniform_real_distribution<double> U(0, 1);
    ModelParams mp; for (int i = 0; i < FEDMX_MAX_PARAMS; ++i) mp.p[i] = 0.5 + U(g);
    int bad = 0, cells_checked = 0; double worst = 0.0;
    for (int trial = 0; trial < 8; ++trial) {
        V v[3];
        const double xy[3][2] = {{0.1, 0.2}, {0.4, 0.25}, {0.2, 0.6}};
        for (int b = 0; b < 3; ++b) {
            v[b].id = (b * 7 + 3 + trial) % 5; v[b].r = xy[b][0] + 0.05 * U(g); v[b].z = xy[b][1] + 0.05 * U(g);
            for (int k = 0; k < NEQ; ++k) v[b].u[k] = U(g) + 0.2;
            for (int j = 0; j < NF; ++j) v[b].f[j] = U(g) + 0.2;
        }
        const uint32_t m1 = trial & 3, m2 = (trial >> 1) & 3;
        RowOut<NEQ> o, o0;
        JitModel::row<true>(mp, v[0], v[1], v[2], m1, m2, o);
        JitModel::row<false>(mp, v[0], v[1], v[2], m1, m2, o0);
        for (int k = 0; k < NEQ; ++k) if (o.R[k] != o0.R[k]) { printf("residual differs k %d\n", k); ++bad; }
        if (JitModel::CELLS) {
            // cell-wise variant: record i of cell(v0, v1, v2) == row(a = v_i, p = v_j, n = v_k) without facet terms
            constexpr int NB = NEQ * NEQ, RJ = JitModel::REC_JAC, RR = JitModel::REC_RES;
            constexpr int RS = FEDMX_SLICE;                        // record element stride
            static double rec[3 * RJ * RS], rec0[3 * RR * RS];
            JitModel::cell<true>(mp, v[0], v[1], v[2], rec);
            JitModel::cell<false>(mp, v[0], v[1], v[2], rec0);
            for (int i = 0; i < 3; ++i) {
                const int j = (i + 1) % 3, k3 = (i + 2) % 3, perm[3] = {i, j, k3};
                RowOut<NEQ> r;
                JitModel::row<true>(mp, v[i], v[j], v[k3], 0, 0, r);
                for (int k = 0; k < NEQ; ++k) {
                    double scale = fabs(r.R[k]);
                    for (int b = 0; b < 3; ++b) for (int l = 0; l < NEQ; ++l) scale = std::max(scale, fabs(r.J[b][k * NEQ + l]));
                    if (fabs(rec[(i * RJ + k) * RS] - r.R[k]) > 1e-11 * scale || rec0[(i * RR + k) * RS] != rec[(i * RJ + k) * RS]) { printf("cell R differs i %d k %d\n", i, k); ++bad; }
                    for (int b = 0; b < 3; ++b) for (int l = 0; l < NEQ; ++l) {
                        const double c = rec[(i * RJ + NEQ + perm[b] * NB + k * NEQ + l) * RS];
                        if (fabs(c - r.J[b][k * NEQ + l]) > 1e-11 * scale) { printf("cell J differs i %d k %d b %d l %d: %g vs %g\n", i, k, b, l, c, r.J[b][k * NEQ + l]); ++bad; }
                    }
                }
            }
            ++cells_checked;
        }
        for (int k = 0; k < NEQ; ++k) {
            double jmax = 0.0;
            for (int b = 0; b < 3; ++b) for (int l = 0; l < NEQ; ++l) jmax = std::max(jmax, fabs(o.J[b][k * NEQ + l]));
            for (int b = 0; b < 3; ++b) for (int l = 0; l < NEQ; ++l) {
                const double h = 1e-5;
                V vp[3] = {v[0], v[1], v[2]}, vm[3] = {v[0], v[1], v[2]};
                vp[b].u[l] += h; vm[b].u[l] -= h;
                RowOut<NEQ> op, om;
                JitModel::row<false>(mp, vp[0], vp[1], vp[2], m1, m2, op);
                JitModel::row<false>(mp, vm[0], vm[1], vm[2], m1, m2, om);
                const double fd = (op.R[k] - om.R[k]) / (2 * h), ad = o.J[b][k * NEQ + l];
                const double noise = 1e-15 * fabs(o0.R[k]) / h;                 // rounding of the difference quotient
                const double e = fabs(fd - ad) / (jmax + noise + 1e-300);
                worst = std::max(worst, e);
                if (e > 1e-6 + 10 * noise / (jmax + noise + 1e-300)) { printf("trial %d k %d b %d l %d ad %g fd %g\n", trial, k, b, l, ad, fd); ++bad; }
            }
        }
    }
    printf("worst %g cellwise %d bad %d\n", worst, cells_checked, bad);
    return bad != 0;
}
