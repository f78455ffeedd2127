// TEST INFRASTRUCTURE ONLY — never linked into the product library.
// Stand-in for vpp_b200/csrc/common.cuh that lets g++ compile the *stateless* CUDA kernels of the library
// (one thread = a few independent loads/stores: no shared memory, no barriers, no shuffles) as ordinary C++
// and run them thread by thread on the CPU.  tests/emu/build_emu.py copies the .cu sources next to this file,
// rewrites `kernel<<<grid, block, smem, stream>>>(args)` into emu::launch(grid, block, [&]{ kernel(args); }) and
// builds tests/emu/_build/libvppb_emu.so with -fsanitize=alignment,bounds so that a misaligned vector access
// (a fault on the GPU, silently fine on x86) is reported (UBSAN_OPTIONS=log_path, checked after every test).  The launch order can be reversed
// (vppb_emu_set_reverse) to expose results that depend on the order in which threads run.
#pragma once

#include <cuda_runtime.h>  // tests/emu/cuda_runtime.h

#include "vppb.h"

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
template <typename... KArgs, typename... Args>
inline cudaError_t launch_dependent(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t, Args... args) {
  ::emu::launch((long long)grid, (long long)block, [&]() { kernel(static_cast<KArgs>(args)...); }, smem);  // kernels run to completion one after another here
  return cudaSuccess;
}
// blocks run one after another here, so a kernel with grid-wide barriers gets a grid of ONE block
template <typename K> inline int cooperative_grid_limit(K, int) { return 1; }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_cooperative(void (*kernel)(KArgs...), int grid, int block, cudaStream_t, Args... args) {
  if (grid != 1) return 1;
  ::emu::launch((long long)grid, (long long)block, [&]() { kernel(static_cast<KArgs>(args)...); }, 0);
  return cudaSuccess;
}
int sm_count();  // core.cu: cudaDeviceGetAttribute -> 2 here, so every thread runs several trips of its grid-stride loop

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

inline int ld_acquire(const int* p) { return *p; }
inline unsigned long long ld_acquire64(const unsigned long long* p) { return *p; }
inline void st_release64(unsigned long long* p, unsigned long long v) { *p = v; }
inline void st_release(int* p, int v) { *p = v; }
}  // namespace vppb
namespace emu { void yield(); }
namespace vppb {
inline void spin_pause() { emu::yield(); }
inline void grid_dependency_wait() {}
inline void grid_launch_dependents() {}  // a spinning fiber hands the CPU to the other threads of the block

inline void item_divmod(long long i, int d, int& q, int& r) { q = (int)(i / d); r = (int)(i - (long long)q * d); }

// Work item i of copy + fill_border_mirror in one pass (k_copy_mirror of pixelwise.cu, phase 0 of vppb_pyrlk_prepare): rows * nvec
// 16-byte vectors, then rows * tail single bytes, then one item per border pixel of dst - read from the mirrored position in SRC (the
// same value dst's domain receives), so nothing depends on the copy having landed.
inline long long copy_mirror_items(const Img& dst, int nvec, int tail) {
  const long long b = dst.border;
  return (long long)dst.nrows * (nvec + tail) + 2 * b * (dst.ncols + 2 * b) + 2 * b * dst.nrows;
}
inline void copy_mirror_item(const Img& src, const Img& dst, int nvec, int tail, int elem, long long i) {
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
inline void grid_barrier(int* bar, int& gen) {
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
