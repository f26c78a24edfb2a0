// Forward-mode dual numbers for in-kernel Jacobians (value + N partial derivatives), used by GlowModel and by the row
// kernels the form compiler generates.
#pragma once

// N = 0 is the value-only instantiation (residual-only kernels run the same generated expression code without paying for
// derivative lanes); its one-element array is never touched.
template <int N>
struct Dual {
    double v;
    double d[N > 0 ? N : 1];
};

template <int N> __device__ __forceinline__ Dual<N> dconst(double c) {
    Dual<N> r; r.v = c;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = 0.0;
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator+(const Dual<N> &a, const Dual<N> &b) {
    Dual<N> r; r.v = a.v + b.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator-(const Dual<N> &a, const Dual<N> &b) {
    Dual<N> r; r.v = a.v - b.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator-(const Dual<N> &a) {
    Dual<N> r; r.v = -a.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator*(const Dual<N> &a, const Dual<N> &b) {
    Dual<N> r; r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator*(const Dual<N> &a, double c) {
    Dual<N> r; r.v = a.v * c;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * c;
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator*(double c, const Dual<N> &a) { return a * c; }
template <int N> __device__ __forceinline__ Dual<N> operator+(const Dual<N> &a, double c) { Dual<N> r = a; r.v += c; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator+(double c, const Dual<N> &a) { return a + c; }
template <int N> __device__ __forceinline__ Dual<N> operator-(const Dual<N> &a, double c) { Dual<N> r = a; r.v -= c; return r; }
template <int N> __device__ __forceinline__ Dual<N> dexp(const Dual<N> &a) {
    Dual<N> r; r.v = exp(a.v);
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = r.v * a.d[i];
    return r;
}
// |x| with derivative sign(x) (0 at 0), as UFL differentiates abs()
template <int N> __device__ __forceinline__ Dual<N> dabs(const Dual<N> &a) {
    const double s = (a.v > 0.0) ? 1.0 : ((a.v < 0.0) ? -1.0 : 0.0);
    return a * s;
}

template <int N> __device__ __forceinline__ Dual<N> operator-(double c, const Dual<N> &a) { Dual<N> r = -a; r.v += c; return r; }
template <int N> __device__ __forceinline__ Dual<N> operator/(const Dual<N> &a, double c) { return a * (1.0 / c); }
template <int N> __device__ __forceinline__ Dual<N> operator/(const Dual<N> &a, const Dual<N> &b) {
    Dual<N> r; const double ib = 1.0 / b.v; r.v = a.v * ib;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator/(double c, const Dual<N> &b) {
    Dual<N> r; const double ib = 1.0 / b.v; r.v = c * ib; const double f = -r.v * ib;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = f * b.d[i];
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> dsqrt(const Dual<N> &a) {
    Dual<N> r; r.v = sqrt(a.v); const double f = 0.5 / r.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = f * a.d[i];
    return r;
}
template <int N> __device__ __forceinline__ Dual<N> dlog(const Dual<N> &a) {
    Dual<N> r; r.v = log(a.v); const double f = 1.0 / a.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = f * a.d[i];
    return r;
}
// a^c for a constant real exponent
template <int N> __device__ __forceinline__ Dual<N> dpow(const Dual<N> &a, double c) {
    Dual<N> r; r.v = pow(a.v, c); const double f = c * r.v / a.v;          // one pow(): d/da a^c = c a^c / a
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = f * a.d[i];
    return r;
}
template <int N> __device__ __forceinline__ double dsign(const Dual<N> &a) { return (a.v > 0.0) ? 1.0 : ((a.v < 0.0) ? -1.0 : 0.0); }
#ifndef FEDMX_DSIGN_DOUBLE
#define FEDMX_DSIGN_DOUBLE
__device__ __forceinline__ double dsign(double a) { return (a > 0.0) ? 1.0 : ((a < 0.0) ? -1.0 : 0.0); }
#endif
