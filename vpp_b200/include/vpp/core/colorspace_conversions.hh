// rgb_to_graylevel / graylevel_to_rgb (reference: vpp/core/colorspace_conversions.hh:10-60) and the fused frame ingest.
//   o = (i[0] + i[1] + i[2]) / 3, int arithmetic, over in.domain_with_border(); the result has the input's border and
//   alignment.  8-bit RGB / RGBA inputs with an 8-bit output (unsigned char or vuchar1) run the dedicated kernel
//   (vppb_rgb_to_graylevel_u8); any other combination goes through pixel_wise with the same per-pixel arithmetic and
//   therefore needs nvcc.
//   ingest_rgb_frame(frame, border): clone(frame, _border = border); fill_border_mirror; rgb_to_graylevel<unsigned char>
//   (examples/video_extruder.cc:46-48) in ONE launch.
#pragma once
#include <vpp/core/pixel_wise.hh>

namespace vpp {

template <typename T, typename U>
VPP_HD void rgb_to_graylevel(const vector<U, 3>& i, vector<T, 1>& o) { o[0] = T((i[0] + i[1] + i[2]) / 3); }
template <typename T, typename U>
VPP_HD void rgb_to_graylevel(const vector<U, 3>& i, T& o) { o = T((i[0] + i[1] + i[2]) / 3); }

namespace internals {
template <typename T> struct is_u8_gray : std::integral_constant<bool, std::is_same<T, unsigned char>::value || std::is_same<T, vuchar1>::value> {};
template <typename T, typename IN>
struct gray_kernel {  // generic path: one pixel, the reference's arithmetic
  VPP_HD void operator()(vint2, IN& i, T& o) const {
    vector<typename IN::Scalar, 3> rgb(i[0], i[1], i[2]);
    rgb_to_graylevel(rgb, o);
  }
};
}  // namespace internals

namespace internals {
template <typename T, typename U, unsigned CH>
void rgb_to_graylevel_run(const imageNd<vector<U, CH>, 2>& in, imageNd<T, 2>& out, std::true_type /* 8-bit in, 8-bit out */) {
  vppb_check(vppb_rgb_to_graylevel_u8(in.device_read(), out.device_write(), nullptr));
}
template <typename T, typename U, unsigned CH>
void rgb_to_graylevel_run(const imageNd<vector<U, CH>, 2>& in, imageNd<T, 2>& out, std::false_type) {
  pixel_wise(in.domain_with_border(), in, out) | gray_kernel<T, vector<U, CH>>();
}
}  // namespace internals

template <typename T, typename U, unsigned CH>
typename std::enable_if<(CH == 3 || CH == 4), imageNd<T, 2>>::type rgb_to_graylevel(const imageNd<vector<U, CH>, 2>& in) {
  imageNd<T, 2> out(in.domain(), s::_border = in.border(), s::_aligned = in.alignment());
  internals::rgb_to_graylevel_run(in, out, std::integral_constant<bool, std::is_same<U, unsigned char>::value && internals::is_u8_gray<T>::value>());
  return out;
}

// decoded RGB / RGBA frame (any border, not read) -> gray image with a mirror-filled border, one launch
template <unsigned CH>
typename std::enable_if<(CH == 3 || CH == 4), image2d<unsigned char>>::type ingest_rgb_frame(const imageNd<vector<unsigned char, CH>, 2>& frame, int border) {
  image2d<unsigned char> out(frame.domain(), s::_border = border);
  vppb_check(vppb_rgb_to_graylevel_u8_mirror(frame.device_read(), out.device_write(), nullptr));
  return out;
}

}  // namespace vpp
