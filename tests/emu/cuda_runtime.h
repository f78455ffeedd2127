// TEST INFRASTRUCTURE ONLY — never seen by nvcc or the product library.
// The part of CUDA C++ the stateless kernels and the C++ host API headers use, as plain C++ for g++:
// execution-space keywords, threadIdx & co, aligned vector types, a few runtime calls backed by malloc/memcpy,
// and emu::launch, which runs a kernel body once per thread of a 1-D launch (see tests/emu/common.cuh).
#pragma once

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __constant__ static
#define __grid_constant__
#define __align__(n) alignas(n)
#define __restrict__

using std::max;
using std::min;

struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };
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
enum { cudaSuccess = 0, cudaMemcpyDeviceToHost = 2, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToDevice = 3 };
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
// cudaMalloc hands out >= 256-byte aligned memory and vppb_alloc relies on it
inline cudaError_t cudaMalloc(void** p, size_t n) { return posix_memalign(p, 256, n ? n : 1) == 0 ? cudaSuccess : 2; }
template <typename T> inline cudaError_t cudaMalloc(T** p, size_t n) { void* q = nullptr; cudaError_t e = cudaMalloc(&q, n); *p = static_cast<T*>(q); return e; }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }  // cudaFree(0) == free(NULL)
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
typedef void* cudaEvent_t;
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = reinterpret_cast<void*>(1); return cudaSuccess; }  // everything runs in issue order here
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = reinterpret_cast<void*>(1); return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
// CUDA IPC inside one process: the handle carries the pointer itself
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { memset(h, 0, sizeof(*h)); memcpy(h->reserved, &p, sizeof(p)); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof(*p)); return cudaSuccess; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
enum { cudaDevAttrMultiProcessorCount = 16 };
inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 2; return cudaSuccess; }
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };
enum { cudaEnableDefault = 0, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
cudaError_t cudaGetDriverEntryPoint(const char* symbol, void** fn, unsigned long long flags, cudaDriverEntryPointQueryResult* q);
template <typename K> inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, int, cudaStream_t) {
  for (size_t r = 0; r < h; r++) memcpy(static_cast<char*>(d) + r * dp, static_cast<const char*>(s) + r * sp, w);
  return cudaSuccess;
}

// ---- the emulated machine (emu_runtime.cc): every CUDA thread of a block is a fiber; blocks run one after another.
// A thread runs until it reaches __syncthreads or a warp collective, where it yields to the next thread of the block;
// a collective completes when every live lane named in its mask has arrived (exited lanes count as arrived).
// A full round over the block in which nothing progressed is a deadlock and aborts the process with a message.
namespace emu {
extern bool reverse_order;
extern unsigned shuffle_seed;
void run_block(unsigned nthreads, void (*entry)(void*), void* arg);
void block_barrier();
int lane_id();
// all-to-all exchange of one 64-bit value among the lanes of `mask`; returns the mask of lanes that contributed
unsigned warp_exchange(unsigned mask, uint64_t mine, uint64_t out[32]);

void* dyn_smem();                    // the launch's dynamic shared memory (1024-byte aligned)
void set_dyn_smem(size_t bytes);
void yield();
template <typename F>
inline void launch(long long grid, long long block, F body, size_t smem_bytes = 0) {  // 1-D launches only, as the library uses
  set_dyn_smem(smem_bytes);
  gridDim = dim3{(unsigned)grid, 1, 1};
  blockDim = dim3{(unsigned)block, 1, 1};
  for (long long b = 0; b < grid; b++) {
    blockIdx = dim3{(unsigned)(reverse_order ? grid - 1 - b : b), 0, 0};
    run_block((unsigned)block, [](void* f) { (*static_cast<F*>(f))(); }, &body);
  }
}
template <typename F>
inline void launch(dim3 grid, dim3 block, F body, size_t smem_bytes = 0) {  // 2-D launches (the header template kernels): blocks one after another
  set_dyn_smem(smem_bytes);
  gridDim = grid;
  blockDim = block;
  const long long nb = (long long)grid.x * grid.y;
  for (long long b = 0; b < nb; b++) {
    const long long bb = reverse_order ? nb - 1 - b : b;
    blockIdx = dim3{(unsigned)(bb % grid.x), (unsigned)(bb / grid.x), 0};
    run_block(block.x * block.y, [](void* f) { (*static_cast<F*>(f))(); }, &body);
  }
}
template <typename T> inline uint64_t to_bits(T v) { static_assert(sizeof(T) <= 8, "warp collectives move <= 64 bits"); uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
template <typename T> inline T shfl_from(unsigned mask, T v, int src_lane) {
  uint64_t all[32];
  const unsigned got = warp_exchange(mask, to_bits(v), all);
  return (src_lane >= 0 && src_lane < 32 && ((got >> src_lane) & 1u)) ? from_bits<T>(all[src_lane]) : v;
}
}  // namespace emu

#define __shared__ static  // one copy per kernel instantiation; blocks run one at a time

inline void __syncthreads() { emu::block_barrier(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { uint64_t d[32]; emu::warp_exchange(mask, 0, d); }
template <typename T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  const int lane = emu::lane_id(), base = lane & ~(width - 1);
  return emu::shfl_from(mask, v, base + (src & (width - 1)));
}
template <typename T> inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  const int lane = emu::lane_id(), base = lane & ~(width - 1), src = lane - (int)delta;
  return emu::shfl_from(mask, v, src < base ? lane : src);
}
template <typename T> inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  const int lane = emu::lane_id(), base = lane & ~(width - 1), src = lane + (int)delta;
  return emu::shfl_from(mask, v, src > base + width - 1 ? lane : src);
}
template <typename T> inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int width = 32) {
  const int lane = emu::lane_id(), base = lane & ~(width - 1), src = lane ^ lane_mask;
  return emu::shfl_from(mask, v, src > base + width - 1 ? lane : src);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
  uint64_t all[32];
  const unsigned got = emu::warp_exchange(mask, pred ? 1 : 0, all);
  unsigned r = 0;
  for (int l = 0; l < 32; l++)
    if (((got >> l) & 1u) && all[l]) r |= 1u << l;
  return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) {
  uint64_t all[32];
  const unsigned got = emu::warp_exchange(mask, pred ? 1 : 0, all);
  for (int l = 0; l < 32; l++)
    if (((got >> l) & 1u) && !all[l]) return 0;
  return 1;
}

// threads of a block never run at the same time here, so atomics are plain read-modify-writes
template <typename T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <typename T> inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }

// device intrinsics (the emu objects are built with -ffp-contract=off and without fast-math: IEEE single ops)
template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T __ldcs(const T* p) { return *p; }
template <typename T> inline T __ldcg(const T* p) { return *p; }
inline void __threadfence() {}
inline void __nanosleep(unsigned) {}
template <typename T> inline void __stcs(T* p, T v) { *p = v; }
template <typename T> inline void __stcg(T* p, T v) { *p = v; }
inline float __uint_as_float(unsigned v) { float f; memcpy(&f, &v, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned v; memcpy(&v, &f, 4); return v; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fsqrt_rn(float a) { return sqrtf(a); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
inline unsigned __vabsdiffu4(unsigned a, unsigned b) {  // VABSDIFF4.U8: per-byte |a - b|
  unsigned r = 0;
  for (int i = 0; i < 4; i++) {
    const int x = (a >> (8 * i)) & 0xFF, y = (b >> (8 * i)) & 0xFF;
    r |= (unsigned)(x > y ? x - y : y - x) << (8 * i);
  }
  return r;
}
inline unsigned __vcmpgtu4(unsigned a, unsigned b) {  // per byte: 0xFF where a > b (unsigned), else 0
  unsigned r = 0;
  for (int i = 0; i < 4; i++)
    if (((a >> (8 * i)) & 0xFF) > ((b >> (8 * i)) & 0xFF)) r |= 0xFFu << (8 * i);
  return r;
}
inline unsigned __vsadu4(unsigned a, unsigned b) {  // sum of the four per-byte absolute differences
  unsigned r = 0;
  for (int i = 0; i < 4; i++) {
    const int x = (a >> (8 * i)) & 0xFF, y = (b >> (8 * i)) & 0xFF;
    r += (unsigned)(x > y ? x - y : y - x);
  }
  return r;
}
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift) {  // SHF.L.W: upper word of (hi:lo) << (shift & 31)
  shift &= 31;
  return shift ? (hi << shift) | (lo >> (32 - shift)) : hi;
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {  // SHF.R.W: lower word of (hi:lo) >> (shift & 31)
  shift &= 31;
  return shift ? (lo >> shift) | (hi << (32 - shift)) : lo;
}
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {  // PRMT, default mode, selectors 0..7
  const uint64_t v = ((uint64_t)y << 32) | x;
  unsigned r = 0;
  for (int i = 0; i < 4; i++) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xFF) << (8 * i);
  return r;
}
