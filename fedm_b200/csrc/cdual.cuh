// Component-sparse forward-mode dual numbers for the row kernels the form compiler generates.
//
// A row of the element tensor depends on the 3 * n_eq unknowns of its cell, but most sub-expressions of a FEDM form depend on
// one or two COMPONENTS only (e^{u_i}, |grad Phi|, mu(|E|) e^{u_e}, ...).  CD<M> carries derivatives with respect to the
// components in the bit mask M only: 3 lanes (one per cell vertex) per component, lane 3 * rank(c in M) + vertex.  Binary
// operations produce CD<MA | MB>; every lane index is a compile-time constant, so the compiler sees straight-line scalar code
// with no work for structurally zero derivatives (the dense Dual<9> of dual.cuh did 9 lanes for every operation).
// M = 0 is the value-only instantiation used by residual-only kernels.
#pragma once

__host__ __device__ constexpr int cd_popc(unsigned m) { int n = 0; for (; m != 0u; m &= m - 1u) ++n; return n; }
__host__ __device__ constexpr int cd_rank(unsigned m, int c) { return cd_popc(m & ((1u << c) - 1u)); }

template <unsigned M>
struct CD {
    static constexpr int N = 3 * cd_popc(M);
    double v;
    double d[N > 0 ? N : 1];
};

// Static loops over the components of the result mask: every lane index is a constexpr (no run-time indexing of d[], which
// would push the duals into local memory).  F provides one(da) for unary maps; both(da, db) / left(da) / right(db) for binary.
template <unsigned MR, unsigned MA, int C, class F>
__device__ __forceinline__ void cd_map1(const CD<MA> &a, CD<MR> &r, const F &f) {
    if constexpr ((MR >> C) != 0u) {
        if constexpr (((MR >> C) & 1u) != 0u) {
            constexpr int lr = 3 * cd_rank(MR, C);
            if constexpr (((MA >> C) & 1u) != 0u) {
                constexpr int la = 3 * cd_rank(MA, C);
                r.d[lr] = f.one(a.d[la]); r.d[lr + 1] = f.one(a.d[la + 1]); r.d[lr + 2] = f.one(a.d[la + 2]);
            } else {
                r.d[lr] = 0.0; r.d[lr + 1] = 0.0; r.d[lr + 2] = 0.0;
            }
        }
        cd_map1<MR, MA, C + 1, F>(a, r, f);
    }
}

template <unsigned MA, unsigned MB, int C, class F>
__device__ __forceinline__ void cd_map2(const CD<MA> &a, const CD<MB> &b, CD<(MA | MB)> &r, const F &f) {
    constexpr unsigned MR = MA | MB;
    if constexpr ((MR >> C) != 0u) {
        if constexpr (((MR >> C) & 1u) != 0u) {
            constexpr int lr = 3 * cd_rank(MR, C);
            constexpr bool ha = ((MA >> C) & 1u) != 0u, hb = ((MB >> C) & 1u) != 0u;
            if constexpr (ha && hb) {
                constexpr int la = 3 * cd_rank(MA, C), lb = 3 * cd_rank(MB, C);
                r.d[lr] = f.both(a.d[la], b.d[lb]); r.d[lr + 1] = f.both(a.d[la + 1], b.d[lb + 1]); r.d[lr + 2] = f.both(a.d[la + 2], b.d[lb + 2]);
            } else if constexpr (ha) {
                constexpr int la = 3 * cd_rank(MA, C);
                r.d[lr] = f.left(a.d[la]); r.d[lr + 1] = f.left(a.d[la + 1]); r.d[lr + 2] = f.left(a.d[la + 2]);
            } else {
                constexpr int lb = 3 * cd_rank(MB, C);
                r.d[lr] = f.right(b.d[lb]); r.d[lr + 1] = f.right(b.d[lb + 1]); r.d[lr + 2] = f.right(b.d[lb + 2]);
            }
        }
        cd_map2<MA, MB, C + 1, F>(a, b, r, f);
    }
}

struct CdAdd { __device__ __forceinline__ double both(double x, double y) const { return x + y; }
               __device__ __forceinline__ double left(double x) const { return x; }
               __device__ __forceinline__ double right(double y) const { return y; } };
struct CdSub { __device__ __forceinline__ double both(double x, double y) const { return x - y; }
               __device__ __forceinline__ double left(double x) const { return x; }
               __device__ __forceinline__ double right(double y) const { return -y; } };
struct CdMul { double av, bv;                                   // d(ab) = da b + a db
               __device__ __forceinline__ double both(double x, double y) const { return x * bv + av * y; }
               __device__ __forceinline__ double left(double x) const { return x * bv; }
               __device__ __forceinline__ double right(double y) const { return av * y; } };
struct CdDiv { double ib, q;                                    // d(a/b) = (da - q db) / b, q = a/b
               __device__ __forceinline__ double both(double x, double y) const { return (x - q * y) * ib; }
               __device__ __forceinline__ double left(double x) const { return x * ib; }
               __device__ __forceinline__ double right(double y) const { return -(q * y) * ib; } };
struct CdScale { double s; __device__ __forceinline__ double one(double x) const { return s * x; } };

// ---- arithmetic ---------------------------------------------------------------------------------------------------------
template <unsigned MA, unsigned MB> __device__ __forceinline__ CD<(MA | MB)> operator+(const CD<MA> &a, const CD<MB> &b) {
    CD<(MA | MB)> r; r.v = a.v + b.v;
    cd_map2<MA, MB, 0>(a, b, r, CdAdd{});
    return r;
}
template <unsigned MA, unsigned MB> __device__ __forceinline__ CD<(MA | MB)> operator-(const CD<MA> &a, const CD<MB> &b) {
    CD<(MA | MB)> r; r.v = a.v - b.v;
    cd_map2<MA, MB, 0>(a, b, r, CdSub{});
    return r;
}
template <unsigned MA> __device__ __forceinline__ CD<MA> cd_chain(const CD<MA> &a, double value, double slope) {
    CD<MA> r; r.v = value;
    cd_map1<MA, MA, 0>(a, r, CdScale{slope});
    return r;
}
template <unsigned MA> __device__ __forceinline__ CD<MA> operator-(const CD<MA> &a) { return cd_chain(a, -a.v, -1.0); }
template <unsigned MA, unsigned MB> __device__ __forceinline__ CD<(MA | MB)> operator*(const CD<MA> &a, const CD<MB> &b) {
    CD<(MA | MB)> r; r.v = a.v * b.v;
    cd_map2<MA, MB, 0>(a, b, r, CdMul{a.v, b.v});
    return r;
}
template <unsigned MA, unsigned MB> __device__ __forceinline__ CD<(MA | MB)> operator/(const CD<MA> &a, const CD<MB> &b) {
    CD<(MA | MB)> r; const double ib = 1.0 / b.v; r.v = a.v * ib;
    cd_map2<MA, MB, 0>(a, b, r, CdDiv{ib, r.v});
    return r;
}
// with plain doubles
template <unsigned MA> __device__ __forceinline__ CD<MA> operator+(const CD<MA> &a, double c) { CD<MA> r = a; r.v += c; return r; }
template <unsigned MA> __device__ __forceinline__ CD<MA> operator+(double c, const CD<MA> &a) { return a + c; }
template <unsigned MA> __device__ __forceinline__ CD<MA> operator-(const CD<MA> &a, double c) { CD<MA> r = a; r.v -= c; return r; }
template <unsigned MA> __device__ __forceinline__ CD<MA> operator-(double c, const CD<MA> &a) { CD<MA> r = -a; r.v += c; return r; }
template <unsigned MA> __device__ __forceinline__ CD<MA> operator*(const CD<MA> &a, double c) {
    return cd_chain(a, a.v * c, c);
}
template <unsigned MA> __device__ __forceinline__ CD<MA> operator*(double c, const CD<MA> &a) { return a * c; }
template <unsigned MA> __device__ __forceinline__ CD<MA> operator/(const CD<MA> &a, double c) { return a * (1.0 / c); }
template <unsigned MA> __device__ __forceinline__ CD<MA> operator/(double c, const CD<MA> &b) {
    const double ib = 1.0 / b.v, q = c * ib;
    return cd_chain(b, q, -q * ib);
}
// ---- functions ----------------------------------------------------------------------------------------------------------
template <unsigned MA> __device__ __forceinline__ CD<MA> dexp(const CD<MA> &a) { const double e = exp(a.v); return cd_chain(a, e, e); }
template <unsigned MA> __device__ __forceinline__ CD<MA> dsqrt(const CD<MA> &a) { const double s = sqrt(a.v); return cd_chain(a, s, 0.5 / s); }
template <unsigned MA> __device__ __forceinline__ CD<MA> dlog(const CD<MA> &a) { return cd_chain(a, log(a.v), 1.0 / a.v); }
template <unsigned MA> __device__ __forceinline__ CD<MA> dpow(const CD<MA> &a, double c) { const double p = pow(a.v, c); return cd_chain(a, p, c * p / a.v); }
template <unsigned MA> __device__ __forceinline__ CD<MA> dabs(const CD<MA> &a) {
    const double s = (a.v > 0.0) ? 1.0 : ((a.v < 0.0) ? -1.0 : 0.0);
    return a * s;
}
template <unsigned MA> __device__ __forceinline__ double dsign(const CD<MA> &a) { return (a.v > 0.0) ? 1.0 : ((a.v < 0.0) ? -1.0 : 0.0); }

// seed: value + the three vertex weights of component COMP (unused lanes do not exist)
template <unsigned M>
__device__ __forceinline__ CD<M> cd_seed(double value, double w0, double w1, double w2) {
    CD<M> r; r.v = value;
    if constexpr (M != 0) { r.d[0] = w0; r.d[1] = w1; r.d[2] = w2; }
    return r;
}
template <unsigned M>
__device__ __forceinline__ CD<M> cd_zero() {
    CD<M> r; r.v = 0.0;
#pragma unroll
    for (int i = 0; i < CD<M>::N; ++i) r.d[i] = 0.0;
    return r;
}
#ifndef FEDMX_DSIGN_DOUBLE
#define FEDMX_DSIGN_DOUBLE
__device__ __forceinline__ double dsign(double a) { return (a > 0.0) ? 1.0 : ((a < 0.0) ? -1.0 : 0.0); }
#endif
