// semi_dense_optical_flow(keypoints, match_callback, i1, i2, _winsize =, _nscales =, _min_scale =,
// _propagation =, _patchsize =) (reference: vpp/algorithms/optical_flow.hh:29-35,
// optical_flow/semi_dense_optical_flow.hpp:46-214).  keypoints: anything with size() and operator[] -> vint2.
// match_callback(i, new_position, distance) is invoked serially, in keypoint order, as in the reference.
#pragma once
#include <vector>
#include <vpp/algorithms/fast_detector/fast.hh>
#include <vpp/core/pyramid.hh>

namespace vpp {

namespace s {
VPP_DEFINE_SYMBOL(min_scale)
VPP_DEFINE_SYMBOL(propagation)
VPP_DEFINE_SYMBOL(patchsize)
}  // namespace s

template <typename K, typename MC, typename... OPTS>
void semi_dense_optical_flow(const K& keypoints, MC match_callback, const image2d<unsigned char>& i1, const image2d<unsigned char>& i2,
                             OPTS... options) {
  auto opts = s::D(options...);
  vppb_sdof_params P;
  P.winsize = opts.get(s::_winsize, 7);   // defaults: semi_dense_optical_flow.hpp:57-61
  P.nscales = opts.get(s::_nscales, 4);
  P.min_scale = opts.get(s::_min_scale, 0);
  P.propagation = opts.get(s::_propagation, 2);
  P.patchsize = opts.get(s::_patchsize, 5);
  const int n = (int)keypoints.size();
  std::vector<vint2> kps(n);
  for (int i = 0; i < n; i++) kps[i] = keypoints[i];
  // both pyramids (semi_dense_optical_flow.hpp:70-100) in one launch
  pyramid2d<unsigned char> p1(i1.domain(), P.nscales, 2, s::_border = 2 * P.winsize), p2(i2.domain(), P.nscales, 2, s::_border = 2 * P.winsize);
  std::vector<vppb_img> a(P.nscales), b(P.nscales);
  for (int l = 0; l < P.nscales; l++) { a[l] = *p1[l].device_write(); b[l] = *p2[l].device_write(); }
  vppb_check(vppb_pyrlk_prepare(i1.device_read(), i2.device_read(), a.data(), b.data(), nullptr, P.nscales, 0, nullptr));
  const int64_t wsb = vppb_sdof_workspace_bytes(i1.nrows(), i1.ncols(), &P);
  internals::device_array ws((size_t)wsb), dk((size_t)n * 8), dpos((size_t)n * 8), ddist((size_t)n * 4), dvalid((size_t)n);
  dk.from_host(kps.data(), (size_t)n * 8);
  vppb_check(vppb_sdof_u8(a.data(), b.data(), &P, (const vppb_int2*)dk.ptr(), n, ws.ptr(), wsb, (vppb_int2*)dpos.ptr(), (int32_t*)ddist.ptr(),
                          (unsigned char*)dvalid.ptr(), nullptr));
  std::vector<vint2> pos(n);
  std::vector<int> dist(n);
  std::vector<unsigned char> valid(n);
  dpos.to_host(pos.data(), (size_t)n * 8);
  ddist.to_host(dist.data(), (size_t)n * 4);
  dvalid.to_host(valid.data(), (size_t)n);
  for (int i = 0; i < n; i++)
    if (valid[i]) match_callback(i, pos[i], dist[i]);  // :205-212
}

}  // namespace vpp
