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
    int bad = 0; double worst = 0.0;
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
    printf("worst %g bad %d\n", worst, bad);
    return bad != 0;
}
