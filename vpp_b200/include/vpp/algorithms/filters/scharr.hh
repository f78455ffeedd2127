// scharr(in, out) (reference: vpp/algorithms/filters/scharr.hh:46-87): u8 -> vint2 (truncated) or vfloat2.
#pragma once
#include <vpp/core/image2d.hh>

namespace vpp {
inline void scharr(const image2d<unsigned char>& in, image2d<vint2>& out) { vppb_check(vppb_scharr_u8(in.device_read(), out.device_write(), 0, nullptr)); }
inline void scharr(const image2d<unsigned char>& in, image2d<vfloat2>& out) { vppb_check(vppb_scharr_u8(in.device_read(), out.device_write(), 1, nullptr)); }
}  // namespace vpp
