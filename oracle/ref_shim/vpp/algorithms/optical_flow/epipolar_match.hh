// Shadow of vpp/algorithms/optical_flow/epipolar_match.hh (compile-time-off path of the semi-dense flow).
#pragma once
#include <vpp/vpp.hh>
#include <vpp/algorithms/symbols.hh>
namespace vpp {
template <typename D>
inline auto epipolar_match(vint2 p, vint2 prediction, vfloat2, const Eigen::Matrix3f&, D distance) {
  return iod::D(s::_flow = vint2(prediction - p), s::_distance = distance(p, prediction, INT_MAX));
}
}
