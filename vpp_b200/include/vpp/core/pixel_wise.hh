// pixel_wise(ranges...)(options...) | kernel, relative_access, block_wise, row_wise.
// Reference: vpp/core/pixel_wise.hh:41-50, pixel_wise.hpp:14-217, block_wise.hh:26-78,
// relative_accessor.hh:18-33.
//
// The kernel runs on the GPU, so it must be device code: a functor or lambda annotated VPP_KERNEL
// (`[=] VPP_KERNEL (int& a, int& b, int& c) { a = b + c; }`, nvcc --extended-lambda) that captures by
// value.  A plain host lambda is rejected at compile time - there is no CPU fallback.
// Traversal options keep the reference's meaning: by default every pixel is independent (the
// reference runs rows in parallel); _left_to_right/_right_to_left make the columns of a row
// sequential (one thread per row), _top_to_bottom/_bottom_to_top make the rows sequential (one thread
// per column), both or _no_threads give one sequential thread in the reference's raster order.
#pragma once
#include <tuple>
#include <type_traits>
#include <utility>

#include <vpp/core/image2d.hh>

namespace vpp {

template <typename I>
struct relative_access_ {
  vint2 first_point_coordinates() const { return img.first_point_coordinates(); }
  vint2 last_point_coordinates() const { return img.last_point_coordinates(); }
  I img;  // images share their buffer, so holding a copy is an alias (pixel_wise.hpp:22-25 holds a reference)
};
template <typename I>
relative_access_<I> relative_access(const I& i) { return relative_access_<I>{i}; }

// nbh(dr, dc) -> reference to img(r+dr, c+dc)  (relative_accessor.hh:26-33)
template <typename V>
struct relative_access_kernel {
  V* p;
  int pitch;
  VPP_HD V& operator()(int dr, int dc) const { return *(V*)((char*)p + (long long)dr * pitch + (long long)dc * (long long)sizeof(V)); }
  VPP_HD V& operator()(vint2 d) const { return (*this)(d[0], d[1]); }
};

// ---- box_nbh2d<V, R, C>: the legacy neighbourhood accessor (its header is gone from the reference at this commit,
// superseded by relative_access; the API is the one its users still spell: tests/box_nbh2d.cc:8-29,
// benchmarks/box_5x5_filter.cc:163-172).  Here it is a thin alias over the same accessor as relative_access.
//   range form:  auto Anbh = box_nbh2d<int, 5, 5>(A);
//                pixel_wise(B, Anbh) | [=] VPP_KERNEL (int& b, box_nbh2d_kernel<int, 5, 5>& n) { int s = 0; n.for_all([&s](int& v) { s += v; }); b = s / 25; };
//   point form:  auto n = box_nbh2d<int, 3, 3>(A, vint2(1, 1)); n.for_all(f); n.north() = 3;   (host access, lazy mirror)
// As in the reference nothing checks bounds: the image needs a border >= R/2, C/2 (filled by the caller).
template <typename V, int R, int C>
struct box_nbh2d_kernel {  // what a pixel_wise kernel receives for a box_nbh2d range
  V* p;
  int pitch;
  VPP_HD V& operator()(int dr, int dc) const { return *(V*)((char*)p + (long long)dr * pitch + (long long)dc * (long long)sizeof(V)); }
  VPP_HD V& operator()(vint2 d) const { return (*this)(d[0], d[1]); }
  VPP_HD V& north() const { return (*this)(-1, 0); }
  VPP_HD V& south() const { return (*this)(1, 0); }
  VPP_HD V& east() const { return (*this)(0, 1); }
  VPP_HD V& west() const { return (*this)(0, -1); }
  template <typename F>
  VPP_HD void for_all(F f) const {  // row-major over the R x C window centred on the pixel
    for (int dr = -(R / 2); dr <= R / 2; dr++)
      for (int dc = -(C / 2); dc <= C / 2; dc++) f((*this)(dr, dc));
  }
};
template <typename V, int R, int C>
struct box_nbh2d_range {
  vint2 first_point_coordinates() const { return img.first_point_coordinates(); }
  vint2 last_point_coordinates() const { return img.last_point_coordinates(); }
  imageNd<V, 2> img;  // shares the buffer
};
template <typename V, int R, int C>
struct box_nbh2d_point {  // host-side accessor around one pixel; every access goes through the image's lazy host mirror
  imageNd<V, 2> img;
  vint2 p;
  V& operator()(int dr, int dc) { return img(p[0] + dr, p[1] + dc); }
  V& operator()(vint2 d) { return (*this)(d[0], d[1]); }
  V& north() { return (*this)(-1, 0); }
  V& south() { return (*this)(1, 0); }
  V& east() { return (*this)(0, 1); }
  V& west() { return (*this)(0, -1); }
  template <typename F>
  void for_all(F f) {
    for (int dr = -(R / 2); dr <= R / 2; dr++)
      for (int dc = -(C / 2); dc <= C / 2; dc++) f((*this)(dr, dc));
  }
};
template <typename V, int R, int C>
box_nbh2d_range<V, R, C> box_nbh2d(const imageNd<V, 2>& img) { return box_nbh2d_range<V, R, C>{img}; }
template <typename V, int R, int C>
box_nbh2d_point<V, R, C> box_nbh2d(const imageNd<V, 2>& img, vint2 p) { return box_nbh2d_point<V, R, C>{img, p}; }

namespace pixel_wise_internals {

// device-side views of the pixel_wise arguments
template <typename V>
struct image_view {
  unsigned char* base; int pitch;
  VPP_HD V& at(int r, int c) const { return *(V*)(base + (long long)r * pitch + (long long)c * (long long)sizeof(V)); }
};
struct box_view {
  VPP_HD vint2 at(int r, int c) const { return vint2(r, c); }
};
template <typename V>
struct relative_view {
  unsigned char* base; int pitch;
  VPP_HD relative_access_kernel<V> at(int r, int c) const {
    return relative_access_kernel<V>{(V*)(base + (long long)r * pitch + (long long)c * (long long)sizeof(V)), pitch};
  }
};

template <typename V> image_view<V> make_view(const imageNd<V, 2>& i) { const vppb_img* d = i.device_write(); return image_view<V>{(unsigned char*)d->base, d->pitch}; }
inline box_view make_view(const box2d&) { return box_view(); }
template <typename V> relative_view<V> make_view(const relative_access_<imageNd<V, 2>>& r) {
  const vppb_img* d = r.img.device_write(); return relative_view<V>{(unsigned char*)d->base, d->pitch}; }
template <typename V, int R, int C>
struct nbh_view {
  unsigned char* base; int pitch;
  VPP_HD box_nbh2d_kernel<V, R, C> at(int r, int c) const {
    return box_nbh2d_kernel<V, R, C>{(V*)(base + (long long)r * pitch + (long long)c * (long long)sizeof(V)), pitch};
  }
};
template <typename V, int R, int C> nbh_view<V, R, C> make_view(const box_nbh2d_range<V, R, C>& r) {
  const vppb_img* d = r.img.device_write(); return nbh_view<V, R, C>{(unsigned char*)d->base, d->pitch}; }

// The kernel's arguments are handed over as lvalues, so that it may take an accessor by value, by const& or - as the
// reference's users of box_nbh2d do (`auto& a_nbh`) - by non-const reference; pixel references stay references.
template <typename F, typename... A>
VPP_HD void invoke_lv(F& fun, A&&... a) { fun(a...); }

#if defined(__CUDACC__)
enum { MODE_PARALLEL = 0, MODE_ROW_THREADS = 1, MODE_COL_THREADS = 2, MODE_SERIAL = 3 };

template <typename F, typename... Views>
__global__ void pixel_wise_kernel(F fun, int r0, int c0, int nr, int nc, int mode, int rows_desc, int cols_desc, Views... views) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (mode == MODE_PARALLEL) {
    for (long long i = tid; i < (long long)nr * nc; i += stride) {
      const int r = r0 + (int)(i / nc), c = c0 + (int)(i % nc);
      invoke_lv(fun, views.at(r, c)...);
    }
  } else if (mode == MODE_ROW_THREADS) {  // rows independent, columns in order (process_row, pixel_wise.hpp:69-81)
    for (long long i = tid; i < nr; i += stride)
      for (int k = 0; k < nc; k++) {
        const int c = cols_desc ? c0 + nc - 1 - k : c0 + k;
        invoke_lv(fun, views.at(r0 + (int)i, c)...);
      }
  } else if (mode == MODE_COL_THREADS) {  // columns independent, rows in order
    for (long long i = tid; i < nc; i += stride)
      for (int k = 0; k < nr; k++) {
        const int r = rows_desc ? r0 + nr - 1 - k : r0 + k;
        invoke_lv(fun, views.at(r, c0 + (int)i)...);
      }
  } else if (tid == 0) {  // pixel_wise_row_first_serial_2d (pixel_wise.hpp:105-126)
    for (int kr = 0; kr < nr; kr++)
      for (int kc = 0; kc < nc; kc++) {
        const int r = rows_desc ? r0 + nr - 1 - kr : r0 + kr;
        const int c = cols_desc ? c0 + nc - 1 - kc : c0 + kc;
        invoke_lv(fun, views.at(r, c)...);
      }
  }
}
#endif

template <typename P> struct arg_value;  // what the kernel receives for a range
template <typename V> struct arg_value<imageNd<V, 2>> { typedef V& type; };
template <> struct arg_value<box2d> { typedef vint2 type; };
template <typename V> struct arg_value<relative_access_<imageNd<V, 2>>> { typedef relative_access_kernel<V> type; };
template <typename V, int R, int C> struct arg_value<box_nbh2d_range<V, R, C>> { typedef box_nbh2d_kernel<V, R, C>& type; };
template <typename P> using arg_value_t = typename arg_value<typename std::decay<P>::type>::type;

}  // namespace pixel_wise_internals

namespace pixel_wise_internals {
// o = fun(args...) for value-returning kernels (pixel_wise.hpp:205-209)
template <typename F, typename R>
struct assign_result {
  F fun;
  template <typename... A>
  VPP_HD void operator()(R& o, A&&... a) const { o = fun(static_cast<A&&>(a)...); }
};
}  // namespace pixel_wise_internals

template <typename OPTS, typename... Params>
struct pixel_wise_impl {
  pixel_wise_impl(std::tuple<Params...> t, OPTS opts) : ps(t), options(opts) {}

  template <typename... A>
  auto operator()(A... opts) {
    auto o = s::D(opts...);
    return pixel_wise_impl<decltype(o), Params...>(ps, o);
  }

  template <typename F, std::size_t... I>
  void run(F fun, std::index_sequence<I...>) {
#if defined(__CUDACC__)
    using namespace pixel_wise_internals;
    const vint2 p1 = std::get<0>(ps).first_point_coordinates(), p2 = std::get<0>(ps).last_point_coordinates();
    const int nr = p2[0] - p1[0] + 1, nc = p2[1] - p1[1] + 1;
    if (nr <= 0 || nc <= 0) return;
    const bool col_dep = OPTS::has(s::_left_to_right) || OPTS::has(s::_right_to_left);
    const bool row_dep = OPTS::has(s::_top_to_bottom) || OPTS::has(s::_bottom_to_top);
    int mode = MODE_PARALLEL;
    if (OPTS::has(s::_no_threads) || (col_dep && row_dep)) mode = MODE_SERIAL;
    else if (col_dep) mode = MODE_ROW_THREADS;
    else if (row_dep) mode = MODE_COL_THREADS;
    const long long items = mode == MODE_PARALLEL ? (long long)nr * nc : (mode == MODE_ROW_THREADS ? nr : (mode == MODE_COL_THREADS ? nc : 1));
    long long blocks = (items + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    pixel_wise_kernel<<<(int)blocks, 256>>>(fun, p1[0], p1[1], nr, nc, mode, OPTS::has(s::_bottom_to_top) ? 1 : 0,
                                            OPTS::has(s::_right_to_left) ? 1 : 0, make_view(std::get<I>(ps))...);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw std::runtime_error(std::string("pixel_wise launch: ") + cudaGetErrorString(e));
#else
    static_assert(sizeof(F) == 0, "pixel_wise(...) | kernel needs device code: compile this translation unit with nvcc "
                                  "(--extended-lambda) and annotate the kernel VPP_KERNEL; there is no CPU fallback");
#endif
  }

  template <typename F>
  using kernel_return_type = decltype(std::declval<F>()(std::declval<pixel_wise_internals::arg_value_t<Params>>()...));

  // void kernel: run in place.  Value-returning kernel: build and return image2d<ret> (pixel_wise.hpp:198-211).
  template <typename F>
  auto operator|(F fun) { return dispatch(fun, std::is_void<kernel_return_type<F>>()); }

  std::tuple<Params...> ps;
  OPTS options;

 private:
  template <typename F>
  void dispatch(F fun, std::true_type) { run(fun, std::make_index_sequence<sizeof...(Params)>()); }
  template <typename F>
  auto dispatch(F fun, std::false_type) {
    typedef typename std::decay<kernel_return_type<F>>::type value_type;
    const vint2 p1 = std::get<0>(ps).first_point_coordinates(), p2 = std::get<0>(ps).last_point_coordinates();
    image2d<value_type> out(box2d(p1, p2));
    auto ranges = std::tuple_cat(std::make_tuple(out), ps);
    pixel_wise_internals::assign_result<F, value_type> k{fun};
    pixel_wise_impl<OPTS, image2d<value_type>, Params...>(ranges, options).run(k, std::make_index_sequence<sizeof...(Params) + 1>());
    return out;
  }
};

struct pixel_wise_caller {
  template <typename... T>
  auto operator()(const T&... t) const {
    return pixel_wise_impl<s::options<>, T...>(std::make_tuple(t...), s::options<>());
  }
};
static const pixel_wise_caller pixel_wise = pixel_wise_caller();

// ---- block_wise / row_wise (block_wise.hh:26-78): the callback runs on the host once per block, in
// the order the options ask for, and receives `range | block_box` views; whatever it does to them
// (fill, pixel_wise, ...) is device work.
namespace internals {
template <typename V> imageNd<V, 2> restrict_to(const imageNd<V, 2>& i, const box2d& b) { return i | b; }
inline box2d restrict_to(const box2d&, const box2d& b) { return b; }
}  // namespace internals

template <typename OPTS, typename... Params>
class block_wise_runner {
 public:
  block_wise_runner(vint2 block_size, std::tuple<Params...> t, OPTS o = OPTS()) : block_size_(block_size), ranges_(t), options_(o) {}
  template <typename... A>
  auto operator()(A... opts) {
    auto o = s::D(opts...);
    return block_wise_runner<decltype(o), Params...>(block_size_, ranges_, o);
  }
  template <typename F>
  void operator|(F fun) { call(fun, std::make_index_sequence<sizeof...(Params)>()); }

 private:
  template <typename F, std::size_t... I>
  void call(F fun, std::index_sequence<I...>) {
    const vint2 p1 = std::get<0>(ranges_).first_point_coordinates(), p2 = std::get<0>(ranges_).last_point_coordinates();
    const int rstart = p1[0], rend = p2[0], cstart = p1[1], cend = p2[1];
    const int nr = (1 + rend - rstart + block_size_[0] - 1) / block_size_[0];  // ceil (block_wise.hh:37-38)
    const int nc = (1 + cend - cstart + block_size_[1] - 1) / block_size_[1];
    const bool rdesc = OPTS::has(s::_bottom_to_top), cdesc = OPTS::has(s::_right_to_left);
    for (int kr = 0; kr < nr; kr++)
      for (int kc = 0; kc < nc; kc++) {
        const int br = rdesc ? nr - 1 - kr : kr, bc = cdesc ? nc - 1 - kc : kc;
        const int r1 = rstart + br * block_size_[0], r2 = std::min(rstart + (br + 1) * block_size_[0] - 1, rend);
        const int c1 = cstart + bc * block_size_[1], c2 = std::min(cstart + (bc + 1) * block_size_[1] - 1, cend);
        const box2d b(vint2(r1, c1), vint2(r2, c2));
        fun(internals::restrict_to(std::get<I>(ranges_), b)...);
      }
  }
  vint2 block_size_;
  std::tuple<Params...> ranges_;
  OPTS options_;
};

template <typename... PS>
auto block_wise(vint2 block_size, const PS&... params) {
  return block_wise_runner<s::options<>, PS...>(block_size, std::make_tuple(params...));
}
template <typename P0, typename... PS>
auto row_wise(const P0& a0, const PS&... params) {
  const vint2 p1 = a0.first_point_coordinates(), p2 = a0.last_point_coordinates();
  return block_wise(vint2(1, 1 + p2[1] - p1[1]), a0, params...);
}

}  // namespace vpp
