// TEST INFRASTRUCTURE ONLY — never linked into the product library.
// Stand-in for vpp_b200/csrc/common.cuh that lets g++ compile the *stateless* CUDA kernels of the library
// (one thread = a few independent loads/stores: no shared memory, no barriers, no shuffles) as ordinary C++
// and run them thread by thread on the CPU.  tests/emu/build_emu.py copies the .cu sources next to this file,
// rewrites `kernel<<<grid, block, smem, stream>>>(args)` into emu::launch(grid, block, [&]{ kernel(args); }) and
// builds tests/emu/_build/libvppb_emu.so with -fsanitize=alignment,bounds so that a misaligned vector access
// (a fault on the GPU, silently fine on x86) is reported (UBSAN_OPTIONS=log_path, checked after every test).  The launch order can be reversed
// (vppb_emu_set_reverse) to expose results that depend on the order in which threads run.
#pragma once

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "vppb.h"

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)

struct dim3 { unsigned x, y, z; };
extern thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

// CUDA vector types with the alignment the hardware demands of 64- / 128-bit accesses
struct alignas(8) int2 { int x, y; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) float4 { float x, y, z, w; };
inline int2 make_int2(int x, int y) { return int2{x, y}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaMemcpyDeviceToHost = 2, cudaMemcpyHostToDevice = 1 };
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n); return *p ? cudaSuccess : 2; }
template <typename T> inline cudaError_t cudaMalloc(T** p, size_t n) { *p = static_cast<T*>(malloc(n)); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }

// device intrinsics used by the stateless kernels (the emu library is built with -ffp-contract=off, no fast-math)
template <typename T> inline T __ldg(const T* p) { return *p; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fmul_rn(float a, float b) { return a * b; }
// kernels that communicate between threads are NOT emulated: these stubs only let their translation unit compile
inline int __shfl_down_sync(unsigned, int v, int) { return v; }
inline int atomicAdd(int* p, int v) { int o = *p; *p += v; return o; }

namespace emu {
extern bool reverse_order;
// run `body` once per thread of a <<<grid, block>>> launch (1-D launches only, as the library uses)
template <typename F>
inline void launch(long long grid, long long block, F body) {
  gridDim = dim3{(unsigned)grid, 1, 1};
  blockDim = dim3{(unsigned)block, 1, 1};
  for (long long b = 0; b < grid; b++)
    for (long long t = 0; t < block; t++) {
      blockIdx = dim3{(unsigned)(reverse_order ? grid - 1 - b : b), 0, 0};
      threadIdx = dim3{(unsigned)(reverse_order ? block - 1 - t : t), 0, 0};
      body();
    }
}
}  // namespace emu

namespace vppb {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define VPPB_CUDA(call)                                   \
  do {                                                    \
    cudaError_t e__ = (call);                             \
    if (e__ != cudaSuccess) return ::vppb::cuda_fail(e__, #call); \
  } while (0)
#define VPPB_LAUNCH_CHECK(name) do { } while (0)
#define VPPB_REQUIRE(cond, code, ...)                     \
  do {                                                    \
    if (!(cond)) { ::vppb::set_error(__VA_ARGS__); return (code); } \
  } while (0)

inline cudaStream_t as_stream(void* s) { return s; }
inline int sm_count() { return 2; }  // small grids: every thread runs several trips of its grid-stride loop

struct Img {
  unsigned char* base;
  int nrows, ncols, pitch, border;
};
inline Img view(const vppb_img* i) {
  Img v;
  v.base = static_cast<unsigned char*>(i->base);
  v.nrows = i->nrows; v.ncols = i->ncols; v.pitch = i->pitch; v.border = i->border;
  return v;
}
inline bool same_domain(const vppb_img* a, const vppb_img* b) { return a->nrows == b->nrows && a->ncols == b->ncols; }
template <typename T>
inline T* row_ptr(const Img& im, int r) { return reinterpret_cast<T*>(im.base + (long long)r * im.pitch); }
inline int4 ld_stream(const int4* p) { return *p; }
inline void st_stream(int4* p, const int4& v) { *p = v; }

}  // namespace vppb
