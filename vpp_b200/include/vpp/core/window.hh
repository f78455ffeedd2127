// Named neighbourhood windows c9 / c8 / c5 / c4 and foreach(window, f) (reference: vpp/core/window.hh:11-62):
//   pixel_wise(relative_access(img)) | [=] VPP_KERNEL (relative_access_kernel<int> a) { foreach(c4, [&](vint2 n) { a(n) += a(0, 0); }); };
// The reference builds them from lambdas returning arrays so that the offsets inline; here a window is an empty tag
// type whose offsets come from a constexpr switch, which makes it usable inside device kernels (captured by value,
// no global memory behind it).  Offsets are (row, col), in the reference's order.
#pragma once
#include <vpp/core/vector.hh>

namespace vpp {

template <int N, int MASK>  // MASK: bit i set = the i-th cell of the 3x3 raster (-1,-1) .. (1,1) belongs to the window
struct window3x3 {
  enum { size_ = N };
  VPP_HD static constexpr int size() { return N; }
  VPP_HD static vint2 at(int i) {  // i-th member in raster order
    int seen = 0;
    for (int k = 0; k < 9; k++)
      if ((MASK >> k) & 1) {
        if (seen == i) return vint2(k / 3 - 1, k % 3 - 1);
        seen++;
      }
    return vint2(0, 0);
  }
  VPP_HD vint2 operator[](int i) const { return at(i); }
};

#if defined(__CUDACC__)
#pragma nv_exec_check_disable  // f may be a host lambda when foreach is called from host code
#endif
template <int N, int MASK, typename F>
VPP_HD void foreach(window3x3<N, MASK>, F f) {
#pragma unroll
  for (int i = 0; i < N; i++) f(window3x3<N, MASK>::at(i));
}

static constexpr window3x3<9, 0x1FF> c9{};  // the 3x3 block
static constexpr window3x3<8, 0x1EF> c8{};  // the 3x3 block without its centre
static constexpr window3x3<5, 0x0BA> c5{};  // the cross with its centre
static constexpr window3x3<4, 0x0AA> c4{};  // the cross without its centre

}  // namespace vpp
