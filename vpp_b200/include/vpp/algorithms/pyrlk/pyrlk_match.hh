// pyrlk_match (reference: vpp/algorithms/pyrlk/pyrlk_match.hh:15-55) with the square-window matcher
// lk_match_point_square_win<WS> (pyrlk/lk.hh:9-21,42-175).  The keypoint container is any type with
// size(), operator[] -> {position, alive()}, remove(i), move(i, pos) (keypoint_container.hh:30-70).
#pragma once
#include <vpp/algorithms/lucas_kanade.hh>

namespace vpp {

template <unsigned WS>
struct lk_match_point_square_win { enum { window_size = WS }; };

template <unsigned WS, typename U, typename C>
void pyrlk_match(const pyramid2d<unsigned char>& pyramid_prev, const pyramid2d<vector<U, 2>>& pyramid_prev_grad,
                 const pyramid2d<unsigned char>& pyramid_next, C& keypoints, lk_match_point_square_win<WS>, float min_ev, float max_err,
                 float max_iteration, float convergence_delta, int min_scale = 0) {
  std::vector<vfloat2> kps, flow;
  std::vector<int> idx;
  std::vector<float> dist;
  for (int i = 0; i < (int)keypoints.size(); i++)
    if (keypoints[i].alive()) { kps.push_back(vfloat2(keypoints[i].position[0], keypoints[i].position[1])); idx.push_back(i); }
  vppb_lk_params P;
  P.nlevels = pyramid_prev.size(); P.min_scale = min_scale; P.winsize = WS; P.max_iter = (int)max_iteration;
  P.grad_is_float = std::is_floating_point<U>::value ? 1 : 0; P.err_mode = VPPB_LK_ERR_SAD_OVER_MAD; P.gate_on_max_err = 1;
  P.min_ev = min_ev; P.delta = convergence_delta; P.max_err = max_err; P.factor = pyramid_prev.factor(); P.pred_div = 1.f;
  internals::lk_run(pyramid_prev, pyramid_next, pyramid_prev_grad, P, kps, nullptr, flow, dist);
  const box2d dom = pyramid_prev[0].domain();
  for (size_t k = 0; k < kps.size(); k++) {
    const vfloat2 np = kps[k] + flow[k];
    if (dist[k] > max_err || !dom.has(vint2((int)np[0], (int)np[1]))) keypoints.remove(idx[k]);  // pyrlk_match.hh:44-47
    else keypoints.move(idx[k], np);
  }
}

}  // namespace vpp
