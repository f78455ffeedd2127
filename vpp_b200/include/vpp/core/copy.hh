// copy / copy_with_border / clone / sum (reference: vpp/core/copy.hh:10-27, clone.hh:10-20, sum.hh:12-19).
#pragma once
#include <vpp/core/image2d.hh>

namespace vpp {

template <typename V>
void copy(const imageNd<V, 2>& src, imageNd<V, 2>& dst) { vppb_check(vppb_copy2d(src.device_read(), dst.device_write(), 0, nullptr)); }
template <typename V>
void copy(const imageNd<V, 2>& src, imageNd<V, 2>&& dst) { vppb_check(vppb_copy2d(src.device_read(), dst.device_write(), 0, nullptr)); }
template <typename V>
void copy_with_border(const imageNd<V, 2>& src, imageNd<V, 2>& dst) {
  assert(src.domain() == dst.domain());
  assert(src.border() <= dst.border());
  vppb_check(vppb_copy2d(src.device_read(), dst.device_write(), 1, nullptr));
}

template <typename V, typename... O>
imageNd<V, 2> clone(const imageNd<V, 2>& img, const O&... options) {
  auto o = s::D(options...);
  const int border = o.has(s::_border) ? o.get(s::_border, 0) : img.border();
  const int aligned = o.has(s::_aligned) ? o.get(s::_aligned, 0) : img.alignment();
  imageNd<V, 2> n(img.domain(), s::_border = border, s::_aligned = aligned);
  if (img.border() <= border) copy_with_border(img, n);
  else copy(img, n);
  return n;
}

// sum(img): plus_promotion<V> accumulator; device reduction for the scalar types the C-ABI covers
template <typename V>
typename std::enable_if<std::is_integral<V>::value && (sizeof(V) == 1 || sizeof(V) == 4), plus_promotion<V>>::type
sum(const imageNd<V, 2>& img) {
  int64_t out = 0;
  vppb_check(vppb_sum_i32(img.device_read(), std::is_signed<V>::value ? 1 : 0, &out, nullptr));
  return (plus_promotion<V>)out;
}

}  // namespace vpp
