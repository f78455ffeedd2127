// Shared helpers for the sm_100a kernels behind include/vppb.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/vppb.h"

namespace vppb {

// thread-local error text returned by vppb_last_error()
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define VPPB_CUDA(call)                                   \
  do {                                                    \
    cudaError_t e__ = (call);                             \
    if (e__ != cudaSuccess) return ::vppb::cuda_fail(e__, #call); \
  } while (0)

#define VPPB_LAUNCH_CHECK(name)                           \
  do {                                                    \
    cudaError_t e__ = cudaGetLastError();                 \
    if (e__ != cudaSuccess) return ::vppb::cuda_fail(e__, name); \
  } while (0)

#define VPPB_REQUIRE(cond, code, ...)                     \
  do {                                                    \
    if (!(cond)) { ::vppb::set_error(__VA_ARGS__); return (code); } \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// launch `kernel` so that it may begin before the previous kernel of the stream has finished (it must call grid_dependency_wait())
template <typename... KArgs, typename... Args>
inline cudaError_t launch_dependent(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3((unsigned)block, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Number of SMs of the current device (148 on B200); cached per process.
int sm_count();

// cooperative launch: every CTA of the grid is resident at the same time, so the kernel may use grid-wide barriers.
// cooperative_grid_limit = the largest such grid of `kernel` with `block` threads and no dynamic shared memory.
template <typename K>
inline int cooperative_grid_limit(K kernel, int block) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  return per_sm * sm_count();
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_cooperative(void (*kernel)(KArgs...), int grid, int block, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3((unsigned)block, 1, 1);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Device-side image view (a trimmed vppb_img passed by value to kernels).
struct Img {
  unsigned char* base;  // pixel (0,0)
  int nrows, ncols, pitch, border;
};

inline Img view(const vppb_img* i) {
  Img v;
  v.base = static_cast<unsigned char*>(i->base);
  v.nrows = i->nrows; v.ncols = i->ncols; v.pitch = i->pitch; v.border = i->border;
  return v;
}

inline bool same_domain(const vppb_img* a, const vppb_img* b) {
  return a->nrows == b->nrows && a->ncols == b->ncols;
}

template <typename T>
__device__ __forceinline__ T* row_ptr(const Img& im, int r) {
  return reinterpret_cast<T*>(im.base + (long long)r * im.pitch);
}

// streaming 128-bit accesses: every byte of a map is touched exactly once, keep it out of L1
__device__ __forceinline__ int4 ld_stream(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(int4* p, const int4& v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// flags other CTAs write (dataflow sweeps): acquire load / release store at GPU scope, and the pause of a spin loop
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void spin_pause() { __nanosleep(20); }
// programmatic dependent launch: a kernel launched with launch_dependent() may start (barrier init, descriptor prefetch, loads
// of data the previous kernel does not write) while the previous kernel of the stream drains; grid_dependency_wait() returns
// when that kernel has completed and its writes are visible.
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// q = i / d, r = i % d for a work-item index: 32-bit arithmetic whenever the index fits (always, short of 2^31 items per launch) -
// a 64-bit division costs ~4x the instructions of a 32-bit one, and the item loops do one or two per item
__device__ __forceinline__ void item_divmod(long long i, int d, int& q, int& r) {
  if (i < 0x7fffffffLL) { const unsigned u = (unsigned)i, qq = u / (unsigned)d; q = (int)qq; r = (int)(u - qq * (unsigned)d); }
  else { const long long qq = i / d; q = (int)qq; r = (int)(i - qq * d); }
}

// Work item i of copy + fill_border_mirror in one pass (k_copy_mirror of pixelwise.cu, phase 0 of vppb_pyrlk_prepare): rows * nvec
// 16-byte vectors, then rows * tail single bytes, then one item per border pixel of dst - read from the mirrored position in SRC (the
// same value dst's domain receives), so nothing depends on the copy having landed.
__host__ __device__ __forceinline__ long long copy_mirror_items(const Img& dst, int nvec, int tail) {
  const long long b = dst.border;
  return (long long)dst.nrows * (nvec + tail) + 2 * b * (dst.ncols + 2 * b) + 2 * b * dst.nrows;
}
__device__ __forceinline__ void copy_mirror_item(const Img& src, const Img& dst, int nvec, int tail, int elem, long long i) {
  const int b = dst.border, nr = dst.nrows, nc = dst.ncols;
  const long long n_vec = (long long)nr * nvec, n_tail = (long long)nr * tail;
  const long long wfull = nc + 2LL * b, n_top = (long long)b * wfull, n_side = (long long)nr * b;
  if (i < n_vec) {
    int r, k;
    item_divmod(i, nvec, r, k);
    st_stream(reinterpret_cast<int4*>(dst.base + (long long)r * dst.pitch + k * 16), ld_stream(reinterpret_cast<const int4*>(src.base + (long long)r * src.pitch + k * 16)));
  } else if (i < n_vec + n_tail) {
    const long long j = i - n_vec, r = j / tail;
    const long long off = (long long)nvec * 16 + (j - r * tail);
    dst.base[r * dst.pitch + off] = src.base[r * src.pitch + off];
  } else {
    long long j = i - n_vec - n_tail;
    int r, c;
    if (j < n_top) { r = (int)(j / wfull) - b; c = (int)(j % wfull) - b; }
    else if (j < 2 * n_top) { j -= n_top; r = nr + (int)(j / wfull); c = (int)(j % wfull) - b; }
    else if (j < 2 * n_top + n_side) { j -= 2 * n_top; r = (int)(j / b); c = (int)(j % b) - b; }
    else { j -= 2 * n_top + n_side; r = (int)(j / b); c = nc + (int)(j % b); }
    const int sr = r < 0 ? -r - 1 : (r >= nr ? 2 * nr - r - 1 : r);  // fill.hh:59-82
    const int sc = c < 0 ? -c - 1 : (c >= nc ? 2 * nc - c - 1 : c);
    const unsigned char* s = src.base + (long long)sr * src.pitch + (long long)sc * elem;
    unsigned char* d = dst.base + (long long)r * dst.pitch + (long long)c * elem;
    for (int k = 0; k < elem; k++) d[k] = s[k];
  }
}

// all CTAs of a cooperative launch are resident: a counter barrier in global memory.  `gen` counts the barriers passed (uniform
// over the grid); the counter starts at 0 (the host zeroes it before the launch).
__device__ __forceinline__ void grid_barrier(int* bar, int& gen) {
  __syncthreads();
  gen++;
  if (threadIdx.x == 0) {
    const int target = gen * (int)gridDim.x;
    __threadfence();
    atomicAdd(bar, 1);
    while (ld_acquire(bar) < target) spin_pause();
  }
  __syncthreads();
}

}  // namespace vppb
