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

}  // namespace vppb
