// lucas_kanade(i1, i2, _keypoints =, _flow =, ...) (reference: vpp/algorithms/lucas_kanade.hh:21-24,
// lucas_kanade/lucas_kanade.hpp:135-184) and pyrlk_match (pyrlk/pyrlk_match.hh:15-55).
#pragma once
#include <cmath>
#include <vector>
#include <vpp/algorithms/fast_detector/fast.hh>
#include <vpp/algorithms/filters/scharr.hh>
#include <vpp/core/pyramid.hh>

namespace vpp {

namespace internals {
template <typename G>
inline void lk_run(const pyramid2d<unsigned char>& prev, const pyramid2d<unsigned char>& next, const pyramid2d<G>& grad,
                   const vppb_lk_params& P, const std::vector<vfloat2>& kps, const std::vector<vfloat2>* prediction,
                   std::vector<vfloat2>& flow, std::vector<float>& dist) {
  const size_t n = kps.size();
  std::vector<vppb_img> a(P.nlevels), b(P.nlevels), g(P.nlevels);
  for (int s = 0; s < P.nlevels; s++) { a[s] = *prev[s].device_read(); b[s] = *next[s].device_read(); g[s] = *grad[s].device_read(); }
  device_array dk(n * 8), dp(prediction ? n * 8 : 0), df(n * 8), de(n * 4);
  dk.from_host(kps.data(), n * 8);
  if (prediction) dp.from_host(prediction->data(), n * 8);
  vppb_check(vppb_lk_match_u8(a.data(), b.data(), g.data(), &P, (const vppb_float2*)dk.ptr(), prediction ? (const vppb_float2*)dp.ptr() : nullptr,
                              (int)n, (vppb_float2*)df.ptr(), (float*)de.ptr(), nullptr));
  flow.resize(n); dist.resize(n);
  df.to_host(flow.data(), n * 8);
  de.to_host(dist.data(), n * 4);
}
struct no_prediction { vfloat2 operator()(vfloat2) const { return vfloat2(0.f, 0.f); } };
}  // namespace internals

template <typename... OPTS>
void lucas_kanade(const image2d<unsigned char>& i1, const image2d<unsigned char>& i2, OPTS... opts) {
  auto options = s::D(opts...);
  const int niterations = options.get(s::_niterations, 21);
  const int winsize = options.get(s::_winsize, 11);
  const int nscales = options.get(s::_nscales, 3);
  const int min_ev = (int)options.get(s::_min_ev, 0.0001);  // stored in int by the reference (lucas_kanade.hpp:143-144)
  const int delta = (int)options.get(s::_delta, 0.1);
  auto prediction = options.get(s::_prediction, internals::no_prediction());
  auto flow = options.get(s::_flow, 0);
  auto keypoints = options.get(s::_keypoints, std::vector<vfloat2>());

  pyramid2d<unsigned char> pyramid_prev(i1, nscales, 2, s::_border = winsize / 2);
  pyramid2d<vint2> pyramid_prev_grad(i1.domain(), nscales, 2, s::_border = winsize / 2);
  pyramid2d<unsigned char> pyramid_next(i2, nscales, 2, s::_border = winsize / 2);
  scharr(pyramid_prev[0], pyramid_prev_grad[0]);
  pyramid_prev_grad.propagate_level0();

  std::vector<vfloat2> kps(keypoints.size()), pred(keypoints.size()), fl;
  std::vector<float> dist;
  for (size_t i = 0; i < keypoints.size(); i++) {
    kps[i] = vfloat2(keypoints[i][0], keypoints[i][1]);
    auto p = prediction(keypoints[i]);
    pred[i] = vfloat2(p[0], p[1]);
  }
  vppb_lk_params P;
  P.nlevels = nscales; P.min_scale = 0; P.winsize = winsize; P.max_iter = niterations; P.grad_is_float = 0;
  P.err_mode = VPPB_LK_ERR_SAD; P.gate_on_max_err = 0; P.min_ev = (float)min_ev; P.delta = (float)delta; P.max_err = 0.f;
  P.factor = pyramid_prev.factor(); P.pred_div = float(std::pow(2, nscales));
  internals::lk_run(pyramid_prev, pyramid_next, pyramid_prev_grad, P, kps, &pred, fl, dist);
  for (size_t i = 0; i < keypoints.size(); i++) flow(keypoints[i], fl[i], dist[i]);  // serial callbacks (lucas_kanade.hpp:181)
}

}  // namespace vpp
