// Plain g++ -std=c++14 translation unit: everything except `pixel_wise(...) | kernel` (device code)
// is usable from host-only code; used as a syntax check by build.sh.
#include <vpp/vpp.hh>
#include <vpp/algorithms/fast_detector/fast.hh>
#include <vpp/algorithms/lucas_kanade.hh>
#include <vpp/algorithms/optical_flow.hh>
#include <vpp/algorithms/pyrlk/pyrlk_match.hh>

int host_only_demo() {
  using namespace vpp;
  image2d<vuchar3> a(1080, 1920, _border = 2), b(1080, 1920);
  fill(a, vuchar3(1, 2, 3));
  fill_border_mirror(a);
  vppb_check(vppb_box5x5_u8c3(a.device_read(), b.device_write(), nullptr));
  image2d<int> x(512, 512), y(512, 512), z(512, 512);
  vppb_check(vppb_pw_add_i32(x.device_write(), y.device_read(), z.device_read(), nullptr));
  return b(0, 0)[0] + sum(x);
}
