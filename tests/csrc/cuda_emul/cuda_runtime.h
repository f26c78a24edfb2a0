// Host emulation of the little CUDA that libfedmx's topology code and row kernels need (tests/test_kernel_emulation_cpu.py):
// device memory is host memory, a kernel is a plain function run once per (block, thread) by emul_launch.h.  Only kernels
// without barriers or warp intrinsics can run this way -- the row, cell and gather assembly kernels qualify.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FEDMX_HOST_EMULATION 1
#define __global__
#define __device__
#define __host__
#define __constant__
#define __shared__ static
#define __forceinline__ inline
#define __noinline__
#define __grid_constant__
#define __restrict__
#define __launch_bounds__(...)

struct dim3 { unsigned x = 1, y = 1, z = 1; };
static dim3 blockIdx, threadIdx, blockDim, gridDim;
struct double2 { double x, y; };
template <class T> static inline T __ldg(const T *p) { return *p; }

typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
typedef void *cudaGraphExec_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? cudaSuccess : 2; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
