// Plain g++ -std=c++14 translation unit: everything except `pixel_wise(...) | kernel` (device code)
// is usable from host-only code; used as a syntax check by build.sh.
#include <vpp/vpp.hh>
#include <vpp/algorithms/fast_detector/fast.hh>
#include <vpp/algorithms/lucas_kanade.hh>
#include <vpp/algorithms/optical_flow.hh>
#include <vpp/algorithms/pyrlk/pyrlk_match.hh>
#include <vpp/algorithms/pyrlk/lk.hh>
#include <vpp/algorithms/lbp/lbp_transform.hh>
#include <vpp/algorithms/video_extruder.hh>

int host_only_demo() {
  using namespace vpp;
  image2d<vuchar3> a(1080, 1920, _border = 2), b(1080, 1920);
  fill(a, vuchar3(1, 2, 3));
  fill_border_mirror(a);
  vppb_check(vppb_box5x5_u8c3(a.device_read(), b.device_write(), nullptr));
  image2d<int> x(512, 512), y(512, 512), z(512, 512);
  vppb_check(vppb_pw_add_i32(x.device_write(), y.device_read(), z.device_read(), nullptr));
  // the whole video_extruder is host-callable C++14: no device lambda in user code
  image2d<unsigned char> f1(270, 480, _border = 3), f2(270, 480, _border = 3);
  auto ctx = video_extruder_init(f1.domain());
  video_extruder_update(ctx, f1, f2, _detector_th = 10, _keypoint_spacing = 10);
  image2d<unsigned char> l(270, 480);
  lbp_transform(f1, l);
  local_maxima_filter(x, 3);
  auto ranked = fast_detector9_blockwise_rank(f1, 10, 10, 3);
  return b(0, 0)[0] + sum(x) + ctx.keypoints.size() + l(0, 0) + ranked.size();
}
