#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#define __device__
#define __host__
#define __constant__
#define __noinline__
#define __forceinline__ inline
using std::exp; using std::sqrt; using std::log; using std::pow; using std::fabs;
#define FEDMX_MAX_PARAMS 128
#define FEDMX_SLICE 32
struct ModelParams { double p[FEDMX_MAX_PARAMS]; };
struct AsmArgs { int x; };
typedef void *cudaStream_t;
template <int NEQ, int NF, int NC = 0> struct Vtx { int32_t id; double r, z; double u[NEQ]; double f[NF > 0 ? NF : 1]; double c[NC > 0 ? NC : 1]; };
template <int NEQ> struct RowOut { double R[NEQ]; double J[3][NEQ * NEQ]; };
template <class M> void launch_rows(const AsmArgs &, bool, int, cudaStream_t) {}
inline int cudaGetLastError() { return 0; }
#define cudaSuccess 0
