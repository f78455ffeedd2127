// The point matchers of pyrlk (reference: vpp/algorithms/pyrlk/lk.hh).  lk_match_point_square_win<WS> is the tag pyrlk_match takes
// (pyrlk_match.hh); oriented_lk_match_point_square_win<WS> (lk.hh:23-38,180-317) keeps the reference's call operator for ONE point
// and adds a batched form - the CUDA kernel behind vppb_lk_match_oriented_u8 matches all points of a level in one launch.
#pragma once
#include <utility>
#include <vector>
#include <vpp/algorithms/pyrlk/pyrlk_match.hh>

namespace vpp {

template <unsigned WS>
struct oriented_lk_match_point_square_win {
  enum { window_size = WS };

  // all points at once: flow[i] = v - p (or the reference's failure codes (-1,-1) / (0,0)), err[i] = SAD / (cpt * MAD) or FLT_MAX
  template <typename U>
  void operator()(const std::vector<vfloat2>& p, const std::vector<vfloat2>& tr_prediction, const image2d<unsigned char>& A, const image2d<unsigned char>& B,
                  const image2d<vector<U, 2>>& Ag, float min_ev_th, int max_interations, float convergence_delta, float max_step_norm,
                  const std::vector<vfloat2>& match_direction1, const std::vector<vfloat2>& match_direction2, std::vector<vfloat2>& flow,
                  std::vector<float>& err) const {
    const size_t n = p.size();
    internals::device_array dk(n * 8), dp(n * 8), d1(n * 8), d2(n * 8), df(n * 8), de(n * 4);
    dk.from_host(p.data(), n * 8); dp.from_host(tr_prediction.data(), n * 8);
    d1.from_host(match_direction1.data(), n * 8); d2.from_host(match_direction2.data(), n * 8);
    vppb_check(vppb_lk_match_oriented_u8(A.device_read(), B.device_read(), Ag.device_read(), std::is_floating_point<U>::value ? 1 : 0, (int)WS, min_ev_th,
                                         max_interations, convergence_delta, max_step_norm, (const vppb_float2*)dk.ptr(), (const vppb_float2*)dp.ptr(),
                                         (const vppb_float2*)d1.ptr(), (const vppb_float2*)d2.ptr(), (int)n, (vppb_float2*)df.ptr(), (float*)de.ptr(), nullptr));
    flow.resize(n); err.resize(n);
    df.to_host(flow.data(), n * 8);
    de.to_host(err.data(), n * 4);
  }

  // the reference's signature (lk.hh:27-36): one point per call
  template <typename U>
  std::pair<vfloat2, float> operator()(vfloat2 p, vfloat2 tr_prediction, const image2d<unsigned char>& A, const image2d<unsigned char>& B,
                                       const image2d<vector<U, 2>>& Ag, float min_ev_th, int max_interations, float convergence_delta, float max_step_norm,
                                       vfloat2 match_direction1, vfloat2 match_direction2) const {
    std::vector<vfloat2> flow;
    std::vector<float> err;
    (*this)(std::vector<vfloat2>(1, p), std::vector<vfloat2>(1, tr_prediction), A, B, Ag, min_ev_th, max_interations, convergence_delta, max_step_norm,
            std::vector<vfloat2>(1, match_direction1), std::vector<vfloat2>(1, match_direction2), flow, err);
    return std::make_pair(flow[0], err[0]);
  }
};

}  // namespace vpp
