// Host build of fedm_b200/csrc/cdual.cuh (tests/test_formcompiler_cpu.py): sparse duals vs finite differences and vs the
// dense Dual<9> of dual.cuh on an expression using every operator the form compiler emits.
#include <cmath>
#include <cstdio>
#define __device__
#define __host__
#define __forceinline__ inline
using std::exp; using std::sqrt; using std::log; using std::pow; using std::fabs;
#include "cdual.cuh"
#include "dual.cuh"
template <class A, class B, class C> auto F(A u0, B u1, C g2) {
    auto E = dsqrt(g2 * g2 + 1e-3);
    auto mu = dpow(E, -0.3) * 2.0 + 1.0 / (E + 0.5);
    auto ne = dexp(u0), ni = dexp(u1);
    auto al = dexp(-(3.0 / E)) * E;
    auto S = al * mu * ne - (ni - ne) / (E * ni) + dabs(g2) * dlog(ne + ni) - (2.0 - u1) * dsign(g2);
    return S * u0 + (-g2) / 3.0;
}
static Dual<9> seed9(double v, int comp, const double *w) {
    Dual<9> r; r.v = v; for (int i = 0; i < 9; ++i) r.d[i] = 0.0;
    for (int b = 0; b < 3; ++b) r.d[b * 3 + comp] = w[b];
    return r;
}
int main() {
    int bad = 0;
    double x[3] = {0.3, -0.2, 0.7};
    double w[3][3] = {{0.2, 0.3, 0.5}, {0.2, 0.3, 0.5}, {-1.5, 0.5, 1.0}};
    auto u0 = cd_seed<1u>(x[0], w[0][0], w[0][1], w[0][2]);
    auto u1 = cd_seed<2u>(x[1], w[1][0], w[1][1], w[1][2]);
    auto g2 = cd_seed<4u>(x[2], w[2][0], w[2][1], w[2][2]);
    auto r = F(u0, u1, g2);
    CD<7u> R = cd_zero<7u>();
    R = R + r * 2.0;
    R = R + (u1 * u1) * 1.0;
    auto rv = F(cd_seed<0u>(x[0],0,0,0), cd_seed<0u>(x[1],0,0,0), cd_seed<0u>(x[2],0,0,0));
    printf("v %g vs %g\n", R.v, 2*rv.v + x[1]*x[1]);
    for (int c = 0; c < 3; ++c) {
        double h = 1e-6, xp[3] = {x[0], x[1], x[2]}, xm[3] = {x[0], x[1], x[2]};
        xp[c] += h; xm[c] -= h;
        auto fp = F(cd_seed<0u>(xp[0],0,0,0), cd_seed<0u>(xp[1],0,0,0), cd_seed<0u>(xp[2],0,0,0));
        auto fm = F(cd_seed<0u>(xm[0],0,0,0), cd_seed<0u>(xm[1],0,0,0), cd_seed<0u>(xm[2],0,0,0));
        double fd = 2 * (fp.v - fm.v) / (2 * h) + (c == 1 ? 2 * x[1] : 0);
        for (int b = 0; b < 3; ++b) {
            printf("c%d b%d ad %.10g fd %.10g\n", c, b, R.d[3 * c + b], fd * w[c][b]);
            if (fabs(R.d[3 * c + b] - fd * w[c][b]) > 1e-7 * fabs(fd * w[c][b]) + 1e-9) ++bad;
        }
    }
    // dense reference: identical operation order, so the live lanes must agree to the last bit
    auto d = F(seed9(x[0], 0, w[0]), seed9(x[1], 1, w[1]), seed9(x[2], 2, w[2]));
    for (int c = 0; c < 3; ++c) for (int b = 0; b < 3; ++b) if (d.d[b * 3 + c] != r.d[3 * c + b]) { printf("dense mismatch c%d b%d\n", c, b); ++bad; }
    if (d.v != r.v) ++bad;
    // disjoint masks: structurally absent lanes stay absent
    static_assert(decltype(u0 * u1)::N == 6 && decltype(dexp(g2))::N == 3 && decltype(rv)::N == 0, "lane counts");
    printf("bad %d\n", bad);
    return bad != 0;
}
