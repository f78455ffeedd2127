// image2d<V> = imageNd<V,2> on pitched HBM (reference: vpp/core/imageNd.hh:42-177, imageNd.hpp:26-341).
// Same constructor / option / accessor surface; pixels live on the GPU (include/vppb.h).  Host element
// access (img(r,c), img[r], iteration) goes through a lazily synchronised host mirror: reading
// downloads the buffer once, writing marks the device copy stale, the next device operation uploads.
#pragma once
#include <cassert>
#include <cstring>
#include <initializer_list>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <vpp/core/boxNd.hh>
#include <vpp/core/symbols.hh>

#include "../../../../include/vppb.h"

#ifndef VPP_DEFAULT_IMAGE_ALIGNMENT
#define VPP_DEFAULT_IMAGE_ALIGNMENT 128  // reference: 16 / 32 (imageNd.hpp:10-18)
#endif

namespace vpp {

inline void vppb_check(int rc) {
  if (rc != 0) throw std::runtime_error(std::string("vppb: ") + vppb_last_error());
}

namespace internals {
// One allocation shared by every image / subimage that aliases it.
struct buffer_state {
  vppb_img dev;                     // whole image (pixel (0,0) of the ORIGINAL domain)
  unsigned char* host0 = nullptr;   // host mirror: address of pixel (0,0)
  int host_pitch = 0;
  std::vector<unsigned char> host_storage;
  bool host_valid = false, dev_valid = true;
  bool external_host = false;       // `_data=` wraps caller-owned host memory
  std::shared_ptr<void> external_holder;
  ~buffer_state() { if (dev.alloc) vppb_free(&dev); }

  void ensure_host() {
    if (!host0) {
      int pitch; long long total, origin;
      vppb_layout(dev.nrows, dev.ncols, dev.elem_bytes, dev.border, 16, &pitch, (int64_t*)&total, (int64_t*)&origin);
      host_storage.assign((size_t)total, 0);
      host0 = host_storage.data() + origin;
      host_pitch = pitch;
    }
    if (!host_valid) {
      vppb_check(vppb_download(&dev, host0, host_pitch, 1, nullptr));
      vppb_check(vppb_sync(nullptr));
      host_valid = true;
    }
  }
  void host_written() { dev_valid = false; }
  // the caller is about to overwrite EVERY byte of the mirror on the host: no need to fetch the device copy first
  void host_overwrite_all() {
    if (!host0) { host_valid = true; ensure_host(); }  // allocates the mirror; nothing to download
    host_valid = true;
    dev_valid = false;
  }
  void ensure_device() {
    if (!dev_valid) {
      vppb_check(vppb_upload(&dev, host0, host_pitch, 1, nullptr));
      vppb_check(vppb_sync(nullptr));
      dev_valid = true;
    }
  }
  void device_written() { host_valid = false; }
};
}  // namespace internals

template <typename V, unsigned N>
class imageNd;

template <typename V>
class imageNd<V, 2> {
 public:
  typedef imageNd<V, 2> self;
  typedef V value_type;
  typedef vint2 coord_type;
  typedef box2d domain_type;
  enum { dimension = 2 };

  struct iterator {
    self* img; vint2 p;
    V& operator*() { return (*img)(p); }
    iterator& operator++() { if (p[1] == img->ncols() - 1) { p[1] = 0; p[0]++; } else p[1]++; return *this; }
    bool operator!=(const iterator& o) const { return p != o.p; }
  };

  imageNd() {}
  template <typename... O>
  imageNd(int nrows, int ncols, const O&... opts) { construct(make_box2d(nrows, ncols), s::D(opts...)); }
  template <typename... O>
  imageNd(const std::initializer_list<int>& dims, const O&... opts) { construct(make_box2d(dims.begin()[0], dims.begin()[1]), s::D(opts...)); }
  template <typename... O>
  imageNd(const std::vector<int>& dims, const O&... opts) { construct(make_box2d(dims[0], dims[1]), s::D(opts...)); }
  template <typename... O>
  imageNd(const box2d& domain, const O&... opts) { construct(domain, s::D(opts...)); }
  // copies share the data (imageNd.hpp:77-87)
  imageNd(const self&) = default;
  imageNd(self&&) = default;
  self& operator=(const self&) = default;
  self& operator=(self&&) = default;

  int nrows() const { return view_.nrows; }
  int ncols() const { return view_.ncols; }
  int pitch() const { return view_.pitch; }
  int border() const { return view_.border; }
  int alignment() const { return view_.align; }
  box2d domain() const { return make_box2d(view_.nrows, view_.ncols); }
  box2d domain_with_border() const { return domain() + vpp::border(view_.border); }
  vint2 first_point_coordinates() const { return vint2(0, 0); }
  vint2 last_point_coordinates() const { return vint2(view_.nrows - 1, view_.ncols - 1); }
  bool has(const vint2& p) const { return domain().has(p); }
  bool has_data() const { return !!buf_; }
  int coords_to_offset(const vint2& p) const { return p[0] * view_.pitch + p[1] * (int)sizeof(V); }
  int offset_of(const vint2& p) const { return coords_to_offset(p); }

  // ---- host element access (lazy mirror)
  V& operator()(const vint2& p) { return *host_ptr(p[0], p[1], true); }
  const V& operator()(const vint2& p) const { return *const_cast<self*>(this)->host_ptr(p[0], p[1], false); }
  V& operator()(int r, int c) { return *host_ptr(r, c, true); }
  const V& operator()(int r, int c) const { return *const_cast<self*>(this)->host_ptr(r, c, false); }
  V* operator[](int r) { return host_ptr(r, 0, true); }
  const V* operator[](int r) const { return const_cast<self*>(this)->host_ptr(r, 0, false); }
  V* address_of(const vint2& p) { return host_ptr(p[0], p[1], true); }
  iterator begin() { return iterator{this, vint2(0, 0)}; }
  iterator end() { return iterator{this, vint2(nrows(), 0)}; }
  // raw host mirror of pixel (0,0) (the reference's data() is the buffer start; same for border 0)
  V* data() { return host_ptr(-border(), -border(), true); }

  // imageNd.hpp:280-300 (host evaluation on the mirror; the LK kernels evaluate the same on the device)
  V linear_interpolate(const vfloat2& p) const {
    vint2 x((int)p[0], (int)p[1]);
    float a0 = p[0] - x[0], a1 = p[1] - x[1];
    const V& v00 = (*this)(x[0], x[1]); const V& v10 = (*this)(x[0] + 1, x[1]);
    const V& v01 = (*this)(x[0], x[1] + 1); const V& v11 = (*this)(x[0] + 1, x[1] + 1);
    return interp_(v00, v10, v01, v11, a0, a1);
  }

  // ---- views (imageNd.hpp:324-341): alias the pixels, re-based to (0,0)
  self subimage(const box2d& d) const {
    self r;
    r.buf_ = buf_;
    vppb_check(vppb_subimage(&view_, d.p1()[0], d.p1()[1], d.p2()[0], d.p2()[1], &r.view_));
    r.r0_ = r0_ + d.p1()[0];
    r.c0_ = c0_ + d.p1()[1];
    return r;
  }
  const self const_subimage(const box2d& d) const { return subimage(d); }
  void swap(self& o) { std::swap(buf_, o.buf_); std::swap(view_, o.view_); std::swap(r0_, o.r0_); std::swap(c0_, o.c0_); }
  void set_external_data_holder(void* data, void (*deleter)(void*)) { buf_->external_holder = std::shared_ptr<void>(data, deleter); }

  // ---- device side (used by the operators of this library)
  const vppb_img* device_read() const { buf_->ensure_device(); return &view_; }
  const vppb_img* device_write() const { buf_->ensure_device(); buf_->device_written(); return &view_; }
  // fill(img, v) + border, done on the host mirror (host-side bookkeeping images: keypoint index, detector mask):
  // no kernel launch and no device -> host copy; the device copy is refreshed lazily if a device operator reads it
  void host_fill_with_border(const V& v) {
    assert(buf_ && r0_ == 0 && c0_ == 0 && view_.nrows == buf_->dev.nrows && view_.ncols == buf_->dev.ncols);  // whole images only
    buf_->host_overwrite_all();
    const int b = border();
    for (int r = -b; r < nrows() + b; r++) {
      V* row = (V*)(buf_->host0 + (long long)r * buf_->host_pitch);
      for (int c = -b; c < ncols() + b; c++) row[c] = v;
    }
  }
  // flush device results back into caller-owned host memory (`_data=` images)
  void sync_host() const { buf_->ensure_host(); }

 private:
  template <typename OPTS>
  void construct(const box2d& domain, const OPTS& options) {
    static_assert(!OPTS::has(s::_data) || OPTS::has(s::_pitch),
                  "You must provide the pitch when providing a data pointer to the image constructor.");  // imageNd.hpp:108-110
    buf_ = std::make_shared<internals::buffer_state>();
    const int b = options.get(s::_border, 0);
    const int al = options.get(s::_aligned, VPP_DEFAULT_IMAGE_ALIGNMENT);
    vppb_check(vppb_alloc(&buf_->dev, domain.nrows(), domain.ncols(), (int)sizeof(V), b, al));
    view_ = buf_->dev;
    view_.alloc = nullptr;
    void* ext = (void*)options.get(s::_data, (V*)nullptr);
    if (ext) {  // caller-owned HOST pixels (imageNd.hpp:112-136): they become the host mirror
      buf_->host0 = (unsigned char*)ext;
      buf_->host_pitch = options.get(s::_pitch, 0);
      buf_->external_host = true;
      buf_->host_valid = true;
      buf_->dev_valid = false;
    }
  }
  V* host_ptr(int r, int c, bool will_write) {
    assert(buf_);
    buf_->ensure_host();
    if (will_write) buf_->host_written();
    return (V*)(buf_->host0 + (long long)(r0_ + r) * buf_->host_pitch + (long long)(c0_ + c) * (long long)sizeof(V));
  }
  template <typename T>
  static T interp_(const T& v00, const T& v10, const T& v01, const T& v11, float a0, float a1,
                   typename std::enable_if<std::is_arithmetic<T>::value>::type* = 0) {
    return T((1 - a0) * (1 - a1) * float(v00) + a0 * (1 - a1) * float(v10) + (1 - a0) * a1 * float(v01) + a0 * a1 * float(v11));
  }
  template <typename T>
  static T interp_(const T& v00, const T& v10, const T& v01, const T& v11, float a0, float a1,
                   typename std::enable_if<!std::is_arithmetic<T>::value>::type* = 0) {
    T r;
    for (int i = 0; i < (int)T::SizeAtCompileTime; i++)
      r[i] = typename T::Scalar((1 - a0) * (1 - a1) * float(v00[i]) + a0 * (1 - a1) * float(v10[i]) + (1 - a0) * a1 * float(v01[i]) +
                                a0 * a1 * float(v11[i]));
    return r;
  }

  std::shared_ptr<internals::buffer_state> buf_;
  vppb_img view_ = vppb_img();
  int r0_ = 0, c0_ = 0;  // offset of this view's (0,0) inside the buffer's original domain
};

template <typename V>
using image2d = imageNd<V, 2>;

template <typename V>
imageNd<V, 2> operator|(const imageNd<V, 2>& img, const box2d& b) { return img.subimage(b); }

}  // namespace vpp
