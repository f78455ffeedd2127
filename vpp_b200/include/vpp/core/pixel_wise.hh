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
#include <algorithm>
#include <stdexcept>
#include <string>
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

// `ro`: the kernel takes this range by value or const reference, so the host mirror of the image stays valid
template <typename V> image_view<V> make_view(const imageNd<V, 2>& i, bool ro = false) {
  const vppb_img* d = ro ? i.device_read() : i.device_write(); return image_view<V>{(unsigned char*)d->base, d->pitch}; }
inline box_view make_view(const box2d&, bool = false) { return box_view(); }
template <typename V> relative_view<V> make_view(const relative_access_<imageNd<V, 2>>& r, bool = false) {
  const vppb_img* d = r.img.device_write(); return relative_view<V>{(unsigned char*)d->base, d->pitch}; }
template <typename V, int R, int C>
struct nbh_view {
  unsigned char* base; int pitch;
  VPP_HD box_nbh2d_kernel<V, R, C> at(int r, int c) const {
    return box_nbh2d_kernel<V, R, C>{(V*)(base + (long long)r * pitch + (long long)c * (long long)sizeof(V)), pitch};
  }
};
template <typename V, int R, int C> nbh_view<V, R, C> make_view(const box_nbh2d_range<V, R, C>& r, bool = false) {
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

// ---- vectorised lowering of the default (every pixel independent) traversal when all ranges are plain images / boxes:
// a thread owns VEC consecutive pixels of a row; every image contributes VEC * sizeof(V) bytes, moved as 16-byte
// vectors with streaming loads / stores (each byte of a map is touched once).  The kernel runs on register copies of
// the pixels; a 16-byte vector is written back only if the kernel changed it, so inputs taken as `int&` (the reference's
// users never write const) cost no store traffic.  2-D grid-stride loops: no division per pixel.
template <typename V>
struct vec_image_view {
  unsigned char* base; int pitch;   // base = pixel (r0, c0) of the traversal
  enum { is_image = 1, elem = sizeof(V) };
};
struct vec_box_view { enum { is_image = 0, elem = 0 }; };
template <typename V> vec_image_view<V> make_vec_view(const imageNd<V, 2>& i, bool ro, int r0, int c0) {
  const vppb_img* d = ro ? i.device_read() : i.device_write();
  return vec_image_view<V>{(unsigned char*)d->base + (long long)r0 * d->pitch + (long long)c0 * (long long)sizeof(V), d->pitch}; }
inline vec_box_view make_vec_view(const box2d&, bool, int, int) { return vec_box_view(); }

#if defined(__CUDACC__)
template <typename V, int VEC>
struct vec_regs {  // VEC pixels of one image in registers
  static constexpr int NB = (int)sizeof(V) * VEC, NV = NB / 16;
  union { uint4 q[NV]; V v[VEC]; };
  uint4 q0[NV];
  unsigned char* p;
  VPP_HD vec_regs() {}
  __device__ void load(const vec_image_view<V>& w, int r, int x) {
    p = w.base + (long long)r * w.pitch + (long long)x * (long long)sizeof(V);
#pragma unroll
    for (int i = 0; i < NV; i++) { q[i] = __ldcs(reinterpret_cast<const uint4*>(p) + i); q0[i] = q[i]; }
  }
  __device__ void store() {
#pragma unroll
    for (int i = 0; i < NV; i++)
      if (q[i].x != q0[i].x || q[i].y != q0[i].y || q[i].z != q0[i].z || q[i].w != q0[i].w) __stcs(reinterpret_cast<uint4*>(p) + i, q[i]);
  }
  __device__ V& at(int i, int, int) { return v[i]; }
};
template <int VEC>
struct vec_box_regs {
  vint2 cur;
  __device__ void load(const vec_box_view&, int, int) {}
  __device__ void store() {}
  __device__ vint2& at(int i, int r, int c) { cur = vint2(r, c + i); return cur; }
};
// a tuple usable in device code (std::get is a host function)
template <typename... T> struct dtuple;
template <> struct dtuple<> {};
template <typename H, typename... T> struct dtuple<H, T...> { H head; dtuple<T...> tail; };
template <std::size_t I> struct dget {
  template <typename H, typename... T> static __device__ auto& of(dtuple<H, T...>& t) { return dget<I - 1>::of(t.tail); }
};
template <> struct dget<0> {
  template <typename H, typename... T> static __device__ H& of(dtuple<H, T...>& t) { return t.head; }
};
template <int VEC, typename W> struct vec_regs_of;
template <int VEC, typename V> struct vec_regs_of<VEC, vec_image_view<V>> { typedef vec_regs<V, VEC> type; };
template <int VEC> struct vec_regs_of<VEC, vec_box_view> { typedef vec_box_regs<VEC> type; };

template <typename V> __device__ V& scalar_at(const vec_image_view<V>& w, vint2&, int r, int x, int, int) { return *(V*)(w.base + (long long)r * w.pitch + (long long)x * (long long)sizeof(V)); }
inline __device__ vint2& scalar_at(const vec_box_view&, vint2& tmp, int r, int c, int r0, int c0) { tmp = vint2(r0 + r, c0 + c); return tmp; }

template <int VEC, typename F, std::size_t... I, typename... Views>
__device__ void pixel_wise_vec_body(F& fun, int r0, int c0, int nr, int nc, std::index_sequence<I...>, const Views&... views) {
  const int nchunks = (nc + VEC - 1) / VEC;
  for (int r = blockIdx.y * blockDim.y + threadIdx.y; r < nr; r += gridDim.y * blockDim.y)
    for (int ch = blockIdx.x * blockDim.x + threadIdx.x; ch < nchunks; ch += gridDim.x * blockDim.x) {
      const int x = ch * VEC;
      if (x + VEC <= nc) {
        dtuple<typename vec_regs_of<VEC, Views>::type...> regs;
        int dummy[] = {(dget<I>::of(regs).load(views, r, x), 0)...};
        (void)dummy;
#pragma unroll
        for (int i = 0; i < VEC; i++) invoke_lv(fun, dget<I>::of(regs).at(i, r0 + r, c0 + x)...);
        int dummy2[] = {(dget<I>::of(regs).store(), 0)...};
        (void)dummy2;
      } else {  // ragged end of the row: pixel by pixel, in place
        for (int i = x; i < nc; i++) {
          vint2 tmp[sizeof...(Views)];
          invoke_lv(fun, scalar_at(views, tmp[I], r, i, r0, c0)...);
          (void)tmp;
        }
      }
    }
}
template <int VEC, typename F, typename... Views>
__global__ void pixel_wise_vec_kernel(F fun, int r0, int c0, int nr, int nc, Views... views) {
  pixel_wise_vec_body<VEC>(fun, r0, c0, nr, nc, std::make_index_sequence<sizeof...(Views)>(), views...);
}
#endif

template <typename P> struct vec_capable : std::false_type {};
template <typename V> struct vec_capable<imageNd<V, 2>> : std::integral_constant<bool, (16 % sizeof(V) == 0) || (sizeof(V) % 16 == 0)> {};
template <typename P> struct elem_bytes_of { enum { value = 0 }; };
template <typename V> struct elem_bytes_of<imageNd<V, 2>> { enum { value = sizeof(V) }; };
constexpr int cmin_nz(int a, int b) { return a == 0 ? b : (b == 0 ? a : (a < b ? a : b)); }
constexpr int cmax(int a, int b) { return a > b ? a : b; }
template <typename... P> struct min_elem;
template <> struct min_elem<> { enum { value = 0, maxv = 0 }; };
template <typename P0, typename... P> struct min_elem<P0, P...> {
  enum { value = cmin_nz(elem_bytes_of<P0>::value, min_elem<P...>::value), maxv = cmax(elem_bytes_of<P0>::value, min_elem<P...>::maxv) };
};
template <typename... P> struct all_vec_capable : std::true_type {};
template <typename P0, typename... P> struct all_vec_capable<P0, P...>
    : std::integral_constant<bool, (vec_capable<P0>::value || std::is_same<P0, box2d>::value) && all_vec_capable<P...>::value> {};

template <typename V> bool vec_aligned(const imageNd<V, 2>& i, int c0_rel) {
  const vppb_img* d = i.device_read();
  return (((uintptr_t)d->base + (long long)c0_rel * (long long)sizeof(V)) % 16) == 0 && (d->pitch % 16) == 0; }
inline bool vec_aligned(const box2d&, int) { return true; }
template <typename V> const void* buffer_of(const imageNd<V, 2>& i) { return i.device_read()->base; }
inline const void* buffer_of(const box2d&) { return nullptr; }
// can the range be addressed at every point of [p1, p2]?  (images: inside the domain with its border)
template <typename V> bool covers(const imageNd<V, 2>& i, vint2 p1, vint2 p2) {
  const int b = i.border();
  return p1[0] >= -b && p1[1] >= -b && p2[0] < i.nrows() + b && p2[1] < i.ncols() + b; }
inline bool covers(const box2d&, vint2, vint2) { return true; }
template <typename V> bool covers(const relative_access_<imageNd<V, 2>>& r, vint2 p1, vint2 p2) { return covers(r.img, p1, p2); }
template <typename V, int R, int C> bool covers(const box_nbh2d_range<V, R, C>& r, vint2 p1, vint2 p2) { return covers(r.img, p1, p2); }

// Is the kernel callable when range I is handed over as a const lvalue?  Then it takes it by value or const reference and
// cannot write it.  Only asked of kernels with a single, non-template operator() (for a generic lambda the question
// would instantiate its body with a const argument, a hard error), everything else counts as written.
template <typename F, typename = void> struct has_plain_call : std::false_type {};
template <typename F> struct has_plain_call<F, decltype((void)&F::operator())> : std::true_type {};
template <typename T> struct add_const_ref { typedef const typename std::remove_reference<T>::type& type; };
template <std::size_t K, std::size_t I, typename T> struct const_at { typedef typename std::conditional<K == I, typename add_const_ref<T>::type, T>::type type; };
template <typename F, std::size_t K, typename Seq, typename... A> struct callable_const_at_impl : std::false_type {};
template <typename...> struct voider { typedef void type; };
template <typename F, std::size_t K, typename Enable, typename Seq, typename... A> struct cc_probe : std::false_type {};
template <typename F, std::size_t K, std::size_t... I, typename... A>
struct cc_probe<F, K, typename voider<decltype(std::declval<F&>()(std::declval<typename const_at<K, I, A>::type>()...))>::type, std::index_sequence<I...>, A...> : std::true_type {};
template <typename F, std::size_t K, typename... A>
constexpr bool read_only_arg() { return has_plain_call<F>::value && cc_probe<F, K, void, std::make_index_sequence<sizeof...(A)>, A...>::value; }

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
    // the traversal is the first range's domain, applied to every range (pixel_wise.hpp:147-152): each must hold those points
    const bool ok[] = {covers(std::get<I>(ps), p1, p2)...};
    for (std::size_t k = 0; k < sizeof...(I); k++)
      if (!ok[k]) throw std::runtime_error("pixel_wise: a range does not cover the domain of the first one");
    static int sms = 0;
    if (!sms) {
      int dev = 0;
      if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    const bool ro[] = {read_only_arg<F, I, arg_value_t<Params>...>()...};
    if (mode == MODE_PARALLEL && try_vectorised(fun, p1, nr, nc, sms, ro, std::index_sequence<I...>(), all_vec_capable<typename std::decay<Params>::type...>())) return;
    const long long items = mode == MODE_PARALLEL ? (long long)nr * nc : (mode == MODE_ROW_THREADS ? nr : (mode == MODE_COL_THREADS ? nc : 1));
    long long blocks = (items + 255) / 256;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    pixel_wise_kernel<<<(int)blocks, 256>>>(fun, p1[0], p1[1], nr, nc, mode, OPTS::has(s::_bottom_to_top) ? 1 : 0,
                                            OPTS::has(s::_right_to_left) ? 1 : 0, make_view(std::get<I>(ps), ro[I])...);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw std::runtime_error(std::string("pixel_wise launch: ") + cudaGetErrorString(e));
#else
    static_assert(sizeof(F) == 0, "pixel_wise(...) | kernel needs device code: compile this translation unit with nvcc "
                                  "(--extended-lambda) and annotate the kernel VPP_KERNEL; there is no CPU fallback");
#endif
  }

#if defined(__CUDACC__)
  // all ranges are images (element size dividing 16) or boxes: 16-byte vectors per thread when the rows allow it
  template <typename F, std::size_t... I>
  bool try_vectorised(F& fun, vint2 p1, int nr, int nc, int sms, const bool* ro, std::index_sequence<I...>, std::true_type) {
    using namespace pixel_wise_internals;
    constexpr int minE = min_elem<typename std::decay<Params>::type...>::value, maxE = min_elem<typename std::decay<Params>::type...>::maxv;
    constexpr int VEC = minE > 0 && minE <= 16 ? 16 / minE : 1;
    if (minE == 0 || VEC * maxE > 64) return false;  // no image among the ranges, or too many registers per thread
    // every range is addressed at the first range's coordinates, as in the scalar traversal
    const bool aligned[] = {vec_aligned(std::get<I>(ps), p1[1])...};
    for (std::size_t k = 0; k < sizeof...(I); k++)
      if (!aligned[k]) return false;
    const void* bufs[] = {buffer_of(std::get<I>(ps))...};
    for (std::size_t a = 0; a < sizeof...(I); a++)  // the same pixels through two ranges: the kernel must see one memory
      for (std::size_t b = a + 1; b < sizeof...(I); b++)
        if (bufs[a] && bufs[a] == bufs[b]) return false;
    const int nchunks = (nc + VEC - 1) / VEC;
    dim3 block(32, 8);
    long long gx = (nchunks + 31) / 32, gy = (nr + 7) / 8;
    const long long cap = (long long)sms * 8;  // 8 CTAs of 256 threads per SM in flight, the rest by grid stride
    if (gx * gy > cap) {
      if (gx > cap) { gx = cap; gy = 1; } else { gy = std::max<long long>(1, cap / gx); }
    }
    pixel_wise_vec_kernel<VEC><<<dim3((unsigned)gx, (unsigned)gy), block>>>(fun, p1[0], p1[1], nr, nc, make_vec_view(std::get<I>(ps), ro[I], p1[0], p1[1])...);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw std::runtime_error(std::string("pixel_wise launch: ") + cudaGetErrorString(e));
    return true;
  }
  template <typename F, std::size_t... I>
  bool try_vectorised(F&, vint2, int, int, int, const bool*, std::index_sequence<I...>, std::false_type) { return false; }
#endif

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

// Device form of the block callback: a VPP_KERNEL functor that takes block_view<V> (for image ranges) / box2d (for box ranges)
// runs for ALL blocks in ONE launch - one thread per block, blocks independent unless a traversal option orders them
// (then a single thread visits them in that order).  The host form (a callback taking image2d<V> sub-images, as every
// user of the reference writes it) is kept: it is chosen whenever the functor accepts sub-images.
template <typename V>
struct block_view {
  unsigned char* base;  // pixel (0,0) of the block
  int pitch, nr, nc;
  vint2 p1;             // position of the block's first pixel in the range it was cut from
  VPP_HD V& operator()(int r, int c) const { return *(V*)(base + (long long)r * pitch + (long long)c * (long long)sizeof(V)); }
  VPP_HD V& operator()(vint2 p) const { return (*this)(p[0], p[1]); }
  VPP_HD int nrows() const { return nr; }
  VPP_HD int ncols() const { return nc; }
  VPP_HD vint2 first_point_coordinates() const { return p1; }
};
namespace internals {
template <typename V> struct block_src { unsigned char* base; int pitch; };  // device view of the whole range
struct block_box_src {};
template <typename V> block_src<V> make_block_src(const imageNd<V, 2>& i) { const vppb_img* d = i.device_write(); return block_src<V>{(unsigned char*)d->base, d->pitch}; }
inline block_box_src make_block_src(const box2d&) { return block_box_src(); }
template <typename V> VPP_HD block_view<V> cut(const block_src<V>& s, int r1, int c1, int r2, int c2) {
  return block_view<V>{s.base + (long long)r1 * s.pitch + (long long)c1 * (long long)sizeof(V), s.pitch, r2 - r1 + 1, c2 - c1 + 1, vint2(r1, c1)}; }
VPP_HD inline box2d cut(const block_box_src&, int r1, int c1, int r2, int c2) { return box2d(vint2(r1, c1), vint2(r2, c2)); }
template <typename P> struct block_arg;
template <typename V> struct block_arg<imageNd<V, 2>> { typedef block_view<V> type; };
template <> struct block_arg<box2d> { typedef box2d type; };
template <typename P> struct host_block_arg { typedef P type; };
template <typename F, typename Enable, typename... A> struct callable_with : std::false_type {};
template <typename F, typename... A>
struct callable_with<F, typename pixel_wise_internals::voider<decltype(std::declval<F&>()(std::declval<A>()...))>::type, A...> : std::true_type {};
#if defined(__CUDACC__)
template <typename F, typename... Src>
__global__ void block_wise_kernel(F fun, int rstart, int cstart, int rend, int cend, int bs0, int bs1, int nbr, int nbc, int serial, int rdesc, int cdesc, Src... src) {
  const long long total = (long long)nbr * nbc;
  const long long first = serial ? 0 : (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long step = serial ? 1 : (long long)gridDim.x * blockDim.x;
  if (serial && (blockIdx.x != 0 || threadIdx.x != 0)) return;
  for (long long i = first; i < total; i += step) {
    int br = (int)(i / nbc), bc = (int)(i - (long long)br * nbc);
    if (rdesc) br = nbr - 1 - br;
    if (cdesc) bc = nbc - 1 - bc;
    const int r1 = rstart + br * bs0, r2 = min(rstart + (br + 1) * bs0 - 1, rend);
    const int c1 = cstart + bc * bs1, c2 = min(cstart + (bc + 1) * bs1 - 1, cend);
    fun(cut(src, r1, c1, r2, c2)...);
  }
}
#endif
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
  void operator|(F fun) {
    // a functor that accepts sub-images is the reference's host callback; one that only accepts block views is device code
    typedef std::integral_constant<bool, !internals::callable_with<F, void, typename internals::host_block_arg<Params>::type...>::value &&
                                             internals::callable_with<F, void, typename internals::block_arg<typename std::decay<Params>::type>::type...>::value> device_form;
    call(fun, std::make_index_sequence<sizeof...(Params)>(), device_form());
  }

 private:
  template <typename F, std::size_t... I>
  void call(F fun, std::index_sequence<I...>, std::true_type) {
#if defined(__CUDACC__)
    const vint2 p1 = std::get<0>(ranges_).first_point_coordinates(), p2 = std::get<0>(ranges_).last_point_coordinates();
    const int nr = (1 + p2[0] - p1[0] + block_size_[0] - 1) / block_size_[0], nc = (1 + p2[1] - p1[1] + block_size_[1] - 1) / block_size_[1];
    if (nr <= 0 || nc <= 0) return;
    const bool serial = OPTS::has(s::_no_threads) || OPTS::has(s::_top_to_bottom) || OPTS::has(s::_bottom_to_top) || OPTS::has(s::_left_to_right) ||
                        OPTS::has(s::_right_to_left);
    const long long total = (long long)nr * nc;
    const int grid = serial ? 1 : (int)std::min<long long>((total + 127) / 128, 148 * 16);
    internals::block_wise_kernel<<<grid, serial ? 32 : 128>>>(fun, p1[0], p1[1], p2[0], p2[1], block_size_[0], block_size_[1], nr, nc, serial ? 1 : 0,
                                                             OPTS::has(s::_bottom_to_top) ? 1 : 0, OPTS::has(s::_right_to_left) ? 1 : 0,
                                                             internals::make_block_src(std::get<I>(ranges_))...);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw std::runtime_error(std::string("block_wise launch: ") + cudaGetErrorString(e));
#else
    static_assert(sizeof(F) == 0, "a block_wise callback over block_view<V> is device code: compile with nvcc (--extended-lambda)");
#endif
  }
  template <typename F, std::size_t... I>
  void call(F fun, std::index_sequence<I...> seq, std::false_type) { call(fun, seq); }
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
