// fill / border fills (reference: vpp/core/fill.hh:12-121) on the GPU through include/vppb.h.
#pragma once
#include <vpp/core/image2d.hh>
#include <vpp/core/pixel_wise.hh>

namespace vpp {

template <typename V, typename U>
void fill(imageNd<V, 2>& img, U&& value) { V v = V(value); vppb_check(vppb_fill(img.device_write(), &v, 0, nullptr)); }
template <typename V, typename U>
void fill(imageNd<V, 2>&& img, U&& value) { V v = V(value); vppb_check(vppb_fill(img.device_write(), &v, 0, nullptr)); }
template <typename V>
void fill(imageNd<V, 2>& img, V value, const box2d& box) { auto sub = img | box; vppb_check(vppb_fill(sub.device_write(), &value, 0, nullptr)); }
template <typename V, typename U>
void fill_with_border(imageNd<V, 2>& img, U&& value) { V v = V(value); vppb_check(vppb_fill(img.device_write(), &v, 1, nullptr)); }
template <typename V, typename U>
void fill_border_with_value(imageNd<V, 2>& img, U&& value) { V v = V(value); vppb_check(vppb_fill_border_value(img.device_write(), &v, nullptr)); }
template <typename V>
void fill_border_mirror(imageNd<V, 2>& img) { vppb_check(vppb_fill_border_mirror(img.device_write(), nullptr)); }
template <typename V>
void fill_border_closest(imageNd<V, 2>& img) { vppb_check(vppb_fill_border_closest(img.device_write(), nullptr)); }

}  // namespace vpp
