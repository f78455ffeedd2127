// pyramid2d<V> (reference: vpp/core/pyramid.hh:126-215), factor 2: each level is the fused
// low-pass + subsample kernel vppb_lowpass_sub2 of the previous one, borders mirror-filled.
#pragma once
#include <vector>
#include <vpp/core/copy.hh>
#include <vpp/core/fill.hh>

namespace vpp {

namespace internals {
template <typename V> struct lowpass_kind;
template <> struct lowpass_kind<unsigned char> { enum { value = 0 }; };
template <> struct lowpass_kind<vint2> { enum { value = 1 }; };
template <> struct lowpass_kind<vfloat2> { enum { value = 2 }; };
}  // namespace internals

template <typename V, unsigned N>
struct pyramid {
  static_assert(N == 2, "only 2-d pyramids");
  typedef imageNd<V, N> image_type;

  template <typename... O>
  pyramid(box2d d, int nlevels, float factor, const O&... image_options) : levels_(nlevels), factor_(factor) {
    if (factor != 2.f) throw std::runtime_error("pyramid: only factor 2 is built (pyramid.hh:174-182)");
    for (int i = 0; i < nlevels; i++) {
      levels_[i] = image_type(d, image_options...);
      d = make_box2d(int(1 + (d.nrows() / factor)), int(1 + (d.ncols() / factor)));  // pyramid.hh:140
    }
  }
  template <typename... O>
  pyramid(const image_type& img, int nlevels, float factor, const O&... image_options)
      : pyramid(img.domain(), nlevels, factor, image_options...) { update(img); }

  image_type& operator[](unsigned i) { return levels_[i]; }
  const image_type& operator[](unsigned i) const { return levels_[i]; }

  void propagate_level0() {  // pyramid.hh:169-192
    fill_border_mirror(levels_[0]);
    propagate_from_mirrored_level0();
  }
  // levels 1.. from a level 0 whose mirror border is already in place: low-pass + subsample + mirror border of
  // each new level in one launch
  void propagate_from_mirrored_level0() {
    for (size_t i = 1; i < levels_.size(); i++)
      vppb_check(vppb_lowpass_sub2_mirror(levels_[i - 1].device_read(), levels_[i].device_write(), internals::lowpass_kind<V>::value, nullptr));
  }
  void update(const image_type& in) {  // pyramid.hh:194-198: copy, mirror, levels
    vppb_check(vppb_copy2d_mirror(in.device_read(), levels_[0].device_write(), nullptr));
    propagate_from_mirrored_level0();
  }
  float factor() const { return factor_; }
  int size() const { return (int)levels_.size(); }
  void swap(pyramid& o) { levels_.swap(o.levels_); std::swap(factor_, o.factor_); }
  std::vector<image_type>& levels() { return levels_; }
  const std::vector<image_type>& levels() const { return levels_; }

 private:
  std::vector<image_type> levels_;
  float factor_;
};
template <typename V> using pyramid2d = pyramid<V, 2>;

}  // namespace vpp
