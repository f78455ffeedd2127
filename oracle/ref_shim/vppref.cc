// C entry points over the reference's OWN headers (/root/reference/vpp/..., included verbatim) so the
// oracle restatement and the CPU baseline can be checked against / timed on the real templates.
// Built only where /root/reference exists (oracle/ref_shim/build_ref.sh -> oracle/_ref/libvppref*.so).
// Images are described like vo_img (host pointer to pixel (0,0), pitch, border) and wrapped zero-copy
// with the reference's `_data = , _pitch = ` constructor (imageNd.hpp:99-141).
#include <vpp/vpp.hh>
#include <vpp/algorithms/filters/scharr.hh>
#include <vpp/core/colorspace_conversions.hh>
#include <vpp/algorithms/fast_detector/fast.hh>
#include <vpp/algorithms/lucas_kanade.hh>
#include <vpp/algorithms/pyrlk/lk.hh>
#include <vpp/algorithms/lbp/lbp_transform.hh>
#include <vpp/algorithms/optical_flow/semi_dense_optical_flow.hpp>
#include <climits>
#include <vpp/algorithms/video_extruder.hh>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace vpp;

extern "C" {
typedef struct vo_img { unsigned char* base; int32_t nrows, ncols, pitch, border, elem; } vo_img;
typedef struct vo_int2 { int32_t r, c; } vo_int2;
typedef struct vo_float2 { float r, c; } vo_float2;
}

template <typename V>
static image2d<V> wrap(const vo_img* d) {
  return image2d<V>(make_box2d(d->nrows, d->ncols), _data = (V*)d->base, _pitch = d->pitch, _border = d->border);
}

template <typename V>
static void copy_out_with_border(const image2d<V>& src, const vo_img* dst) {
  const int b = dst->border;
  for (int r = -b; r < src.nrows() + b; r++)
    memcpy(dst->base + (int64_t)r * dst->pitch - (int64_t)b * sizeof(V), &src(r, -b), (size_t)(src.ncols() + 2 * b) * sizeof(V));
}

extern "C" {

int vppref_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void vppref_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

// benchmarks/image_add.cc:51-57
void vppref_pw_add_i32(const vo_img* a, const vo_img* b, const vo_img* c) {
  auto A = wrap<int>(a), B = wrap<int>(b), C = wrap<int>(c);
  pixel_wise(A, B, C) | [](int& x, int& y, int& z) { x = y + z; };
}

void vppref_fill_border_mirror(const vo_img* img) {
  if (img->elem == 1) { auto I = wrap<unsigned char>(img); fill_border_mirror(I); }
  else if (img->elem == 3) { auto I = wrap<vuchar3>(img); fill_border_mirror(I); }
  else if (img->elem == 4) { auto I = wrap<int>(img); fill_border_mirror(I); }
  else { auto I = wrap<vint2>(img); fill_border_mirror(I); }
}
void vppref_fill_border_closest(const vo_img* img) {
  if (img->elem == 1) { auto I = wrap<unsigned char>(img); fill_border_closest(I); }
  else if (img->elem == 3) { auto I = wrap<vuchar3>(img); fill_border_closest(I); }
  else { auto I = wrap<int>(img); fill_border_closest(I); }
}

// benchmarks/box_5x5_filter2.cc:71-81
void vppref_box5x5_i32(const vo_img* in, const vo_img* out) {
  auto A = wrap<int>(in), B = wrap<int>(out);
  pixel_wise(B, relative_access(A)) | [&](int& b, auto a) {
    int sum = 0;
    for (int i = -2; i <= 2; i++)
      for (int j = -2; j <= 2; j++) sum += a(i, j);
    b = sum / 25;
  };
}
// the same kernel on image2d<vuchar3> with a vint3 accumulator (examples/box_filter.cc:23-32 style)
void vppref_box5x5_u8c3(const vo_img* in, const vo_img* out) {
  auto A = wrap<vuchar3>(in), B = wrap<vuchar3>(out);
  pixel_wise(B, relative_access(A)) | [&](vuchar3& b, auto a) {
    vint3 sum = vint3::Zero();
    for (int i = -2; i <= 2; i++)
      for (int j = -2; j <= 2; j++) sum += a(i, j).template cast<int>();
    b = (sum / 25).cast<unsigned char>();
  };
}

// rgb_to_graylevel<unsigned char>(image2d<vuchar3 | vuchar4>) (colorspace_conversions.hh:22-47): returns a new image with
// the input's border, converted over domain_with_border(); copied out with `out`'s border (<= the input's).
void vppref_rgb_to_graylevel(const vo_img* in, const vo_img* out) {
  if (in->elem == 3) { auto A = wrap<vuchar3>(in); auto G = rgb_to_graylevel<unsigned char>(A); copy_out_with_border(G, out); }
  else { auto A = wrap<vuchar4>(in); auto G = rgb_to_graylevel<unsigned char>(A); copy_out_with_border(G, out); }
}
// the same through the vuchar1 overload the reference's own test uses (tests/colorspace_conversions.cc:18)
void vppref_rgb_to_graylevel_v1(const vo_img* in, const vo_img* out) {
  auto A = wrap<vuchar3>(in);
  image2d<vuchar1> G = rgb_to_graylevel<vuchar1>(A);
  copy_out_with_border(G, out);
}

void vppref_scharr_u8(const vo_img* in, const vo_img* out, int as_float) {
  auto A = wrap<unsigned char>(in);
  if (as_float) { auto G = wrap<vfloat2>(out); scharr(A, G); }
  else { auto G = wrap<vint2>(out); scharr(A, G); }
}

// antialiasing_lowpass_filter (pyramid.hh:12-59), out same domain
void vppref_lowpass_u8(const vo_img* in, const vo_img* out) {
  auto A = wrap<unsigned char>(in), O = wrap<unsigned char>(out);
  antialiasing_lowpass_filter(A, O);
}

// pyramid2d<V>(img, nlevels, 2, _border = levels[0].border) (pyramid.hh:146-198); kind 0 = u8 image pyramid,
// 1 / 2 = vint2 / vfloat2 Scharr-gradient pyramid of the u8 image (lucas_kanade.hpp:153-157)
void vppref_pyramid(const vo_img* in, int nlevels, const vo_img* levels, int kind) {
  auto A = wrap<unsigned char>(in);
  const int b = levels[0].border;
  pyramid2d<unsigned char> P(A, nlevels, 2, _border = b);
  if (kind == 0) {
    for (int i = 0; i < nlevels; i++) copy_out_with_border(P[i], &levels[i]);
  } else if (kind == 1) {
    pyramid2d<vint2> G(A.domain(), nlevels, 2, _border = b);
    scharr(P[0], G[0]);
    G.propagate_level0();
    for (int i = 0; i < nlevels; i++) copy_out_with_border(G[i], &levels[i]);
  } else {
    pyramid2d<vfloat2> G(A.domain(), nlevels, 2, _border = b);
    scharr(P[0], G[0]);
    G.propagate_level0();
    for (int i = 0; i < nlevels; i++) copy_out_with_border(G[i], &levels[i]);
  }
}

// fast9(A, th, [_local_maxima | _blockwise, _block_size, _mask, _scores]) (fast.hpp:931-955); returns the count,
// keypoints sorted in raster order (the reference's order is thread-schedule dependent)
int vppref_fast9_u8(const vo_img* img, int th, const vo_img* mask, int mode, int block_size, vo_int2* kps, int32_t* scores, int capacity) {
  auto A = wrap<unsigned char>(img);
  image2d<unsigned char> M;
  if (mask && mask->base) M = wrap<unsigned char>(mask);
  std::vector<int> sc;
  std::vector<vint2> k;
  if (mode == 1) k = fast9(A, th, _local_maxima, _mask = M, _scores = &sc);
  else if (mode == 2) k = fast9(A, th, _blockwise, _block_size = block_size, _mask = M, _scores = &sc);
  else k = fast9(A, th, _mask = M, _scores = &sc);
  std::vector<int> order(k.size());
  for (size_t i = 0; i < k.size(); i++) order[i] = (int)i;
  std::sort(order.begin(), order.end(), [&](int x, int y) { return k[x][0] != k[y][0] ? k[x][0] < k[y][0] : k[x][1] < k[y][1]; });
  int n = (int)k.size();
  for (int i = 0; i < n && i < capacity; i++) {
    kps[i].r = k[order[i]][0]; kps[i].c = k[order[i]][1];
    if (scores) scores[i] = sc[order[i]];
  }
  return n <= capacity ? n : -n;
}
// blockwise FAST in the order the reference returns it (meaningful in the serial build: cells in raster order)
int vppref_fast9_blockwise_native_order(const vo_img* img, int th, const vo_img* mask, int block_size, vo_int2* kps, int capacity) {
  auto A = wrap<unsigned char>(img);
  image2d<unsigned char> M;
  if (mask && mask->base) M = wrap<unsigned char>(mask);
  auto k = fast9(A, th, _blockwise, _block_size = block_size, _mask = M);
  int n = (int)k.size();
  for (int i = 0; i < n && i < capacity; i++) { kps[i].r = k[i][0]; kps[i].c = k[i][1]; }
  return n;
}
int vppref_fast9_score(const vo_img* img, int th, int r, int c) { auto A = wrap<unsigned char>(img); return fast9_score(A, th, vint2(r, c)); }
// the scalar detector with the TRUE ring (fast.hpp:79-112)
int vppref_is_fast9_keypoint(const vo_img* img, int th, int r, int c) {
  auto A = wrap<unsigned char>(img);
  struct N { const image2d<unsigned char>& a; vint2 p; typedef int value_type; int operator()(int dr, int dc) const { return a(p[0] + dr, p[1] + dc); } };
  return FAST_internals::is_fast9_keypoint(N{A, vint2(r, c)}, th) ? 1 : 0;
}

int vppref_interp_u8(const vo_img* img, float pr, float pc) { auto A = wrap<unsigned char>(img); return A.linear_interpolate(vfloat2(pr, pc)); }

// lucas_kanade(i1, i2, _keypoints, _flow, ...) (lucas_kanade.hpp:135-184)
void vppref_lucas_kanade(const vo_img* i1, const vo_img* i2, const vo_float2* kps, const vo_float2* prediction, int n, int niterations,
                         int winsize, int nscales, double min_ev, double delta, vo_float2* flow_out, float* dist_out) {
  auto I1 = wrap<unsigned char>(i1), I2 = wrap<unsigned char>(i2);
  std::vector<vfloat2> keypoints(n);
  for (int i = 0; i < n; i++) keypoints[i] = vfloat2(kps[i].r, kps[i].c);
  int idx = 0;
  auto cb = [&](vfloat2, vfloat2 f, float d) { flow_out[idx].r = f[0]; flow_out[idx].c = f[1]; dist_out[idx] = d; idx++; };
  int pidx = 0;
  if (prediction) {
    auto pred = [&](vfloat2) { vfloat2 p(prediction[pidx].r, prediction[pidx].c); pidx++; return p; };
    lucas_kanade(I1, I2, _keypoints = keypoints, _niterations = niterations, _winsize = winsize, _nscales = nscales, _min_ev = min_ev,
                 _delta = delta, _prediction = pred, _flow = cb);
  } else {
    lucas_kanade(I1, I2, _keypoints = keypoints, _niterations = niterations, _winsize = winsize, _nscales = nscales, _min_ev = min_ev,
                 _delta = delta, _flow = cb);
  }
}

}  // extern "C"

// The pyrlk_match loop (pyrlk_match.hh:24-41) around the reference's matcher lk_match_point_square_win<WS>
// (lk.hh:42-175) on caller-provided pyramids (levels[] of prev / next u8 and float gradient).
template <unsigned WS>
static void pyrlk_levels(const vo_img* prev, const vo_img* next, const vo_img* grad, int nlevels, int min_scale, const vo_float2* kps, int n,
                         float min_ev, float max_err, float max_iter, float delta, vo_float2* flow_out, float* dist_out) {
  std::vector<image2d<unsigned char>> P, N;
  std::vector<image2d<vfloat2>> G;
  for (int s = 0; s < nlevels; s++) { P.push_back(wrap<unsigned char>(&prev[s])); N.push_back(wrap<unsigned char>(&next[s])); G.push_back(wrap<vfloat2>(&grad[s])); }
  lk_match_point_square_win<WS> matcher;
#pragma omp parallel for
  for (int i = 0; i < n; i++) {
    vfloat2 position(kps[i].r, kps[i].c);
    vfloat2 tr = vfloat2{0.f, 0.f};
    float dist = 0.f;
    for (int S = nlevels - 1; S >= min_scale; S--) {
      tr *= 2.f;
      auto match = matcher(position / std::pow(2, S), tr, P[S], N[S], G[S], min_ev, max_iter, delta);
      if (match.second < max_err) tr = match.first;
      dist = match.second;
    }
    flow_out[i].r = tr[0]; flow_out[i].c = tr[1]; dist_out[i] = dist;
  }
}
extern "C" {
void vppref_pyrlk_levels(const vo_img* prev, const vo_img* next, const vo_img* grad, int nlevels, int min_scale, int winsize, const vo_float2* kps,
                         int n, float min_ev, float max_err, float max_iter, float delta, vo_float2* flow_out, float* dist_out) {
  switch (winsize) {
    case 5: pyrlk_levels<5>(prev, next, grad, nlevels, min_scale, kps, n, min_ev, max_err, max_iter, delta, flow_out, dist_out); break;
    case 7: pyrlk_levels<7>(prev, next, grad, nlevels, min_scale, kps, n, min_ev, max_err, max_iter, delta, flow_out, dist_out); break;
    case 9: pyrlk_levels<9>(prev, next, grad, nlevels, min_scale, kps, n, min_ev, max_err, max_iter, delta, flow_out, dist_out); break;
    default: pyrlk_levels<11>(prev, next, grad, nlevels, min_scale, kps, n, min_ev, max_err, max_iter, delta, flow_out, dist_out); break;
  }
}


// semi_dense_optical_flow (semi_dense_optical_flow.hpp:46-214): out_valid[i] = 1 where match_callback fired
void vppref_semi_dense_flow(const vo_img* i1, const vo_img* i2, const vo_int2* kps, int n, int winsize, int nscales, int min_scale,
                            int propagation, int patchsize, vo_int2* out_pos, int32_t* out_dist, unsigned char* out_valid) {
  auto I1 = wrap<unsigned char>(i1), I2 = wrap<unsigned char>(i2);
  std::vector<vint2> keypoints(n);
  for (int i = 0; i < n; i++) { keypoints[i] = vint2(kps[i].r, kps[i].c); out_valid[i] = 0; out_pos[i].r = out_pos[i].c = 0; out_dist[i] = 0; }
  semi_dense_optical_flow(keypoints, [&](int i, vint2 pos, int d) { out_valid[i] = 1; out_pos[i].r = pos[0]; out_pos[i].c = pos[1]; out_dist[i] = d; },
                          I1, I2, _winsize = winsize, _nscales = nscales, _min_scale = min_scale, _propagation = propagation, _patchsize = patchsize);
}


// video_extruder_init + nframes-1 x video_extruder_update (video_extruder.hpp:15-135) over a frame sequence
// (frames[i]: u8, border >= 3, mirror-filled).  Dumps the final keypoints as rows of 6 ints
// (row, col, age, trajectory start frame, trajectory length, trajectory alive); returns their count.
int vppref_video_extruder(const vo_img* frames, int nframes, int detector_th, int keypoint_spacing, int detector_period, int max_traj, int nscales,
                          int winsize, int propagation, int32_t* out, int capacity) {
  auto ctx = video_extruder_init(make_box2d(frames[0].nrows, frames[0].ncols));
  for (int f = 1; f < nframes; f++) {
    auto f1 = wrap<unsigned char>(&frames[f - 1]), f2 = wrap<unsigned char>(&frames[f]);
    video_extruder_update(ctx, f1, f2, _detector_th = detector_th, _keypoint_spacing = keypoint_spacing, _detector_period = detector_period,
                          _max_trajectory_length = max_traj, _nscales = nscales, _winsize = winsize, _propagation = propagation);
  }
  const int n = ctx.keypoints.size();
  for (int i = 0; i < n && i < capacity; i++) {
    out[6 * i + 0] = ctx.keypoints[i].position[0]; out[6 * i + 1] = ctx.keypoints[i].position[1]; out[6 * i + 2] = ctx.keypoints[i].age;
    out[6 * i + 3] = ctx.trajectories[i].start_frame(); out[6 * i + 4] = ctx.trajectories[i].size(); out[6 * i + 5] = ctx.trajectories[i].alive() ? 1 : 0;
  }
  return n;
}

}  // extern "C"


// ---- SURVEY 8(f) N4 ----------------------------------------------------------------------------------------------
// oriented_lk_match_point_square_win<WS> (lk.hh:180-317), one level, called keypoint by keypoint
template <unsigned WS>
static void oriented_points(const vo_img* a, const vo_img* b, const vo_img* ag, float min_ev, int max_iter, float delta, float max_step,
                            const vo_float2* kps, const vo_float2* pred, const vo_float2* d1, const vo_float2* d2, int n, vo_float2* flow_out,
                            float* err_out) {
  auto A = wrap<unsigned char>(a), B = wrap<unsigned char>(b);
  auto G = wrap<vfloat2>(ag);
  oriented_lk_match_point_square_win<WS> matcher;
  for (int i = 0; i < n; i++) {
    auto m = matcher(vfloat2(kps[i].r, kps[i].c), vfloat2(pred[i].r, pred[i].c), A, B, G, min_ev, max_iter, delta, max_step, vfloat2(d1[i].r, d1[i].c),
                     vfloat2(d2[i].r, d2[i].c));
    flow_out[i].r = m.first[0]; flow_out[i].c = m.first[1]; err_out[i] = m.second;
  }
}

extern "C" {
// lbp_transform (lbp_transform.hh:7-38), unsigned char -> unsigned char
void vppref_lbp_u8(const vo_img* in, const vo_img* out) {
  auto A = wrap<unsigned char>(in), B = wrap<unsigned char>(out);
  lbp_transform(A, B);
}
// local_maxima_filter (fast.hpp:555-575), in place; serial in libvppref.so (no OpenMP)
void vppref_local_maxima_filter(const vo_img* img) {
  if (img->elem == 1) { auto A = wrap<unsigned char>(img); local_maxima_filter(A, 3); }
  else { auto A = wrap<int>(img); local_maxima_filter(A, 3); }
}
void vppref_lk_match_oriented(const vo_img* a, const vo_img* b, const vo_img* ag, int winsize, float min_ev, int max_iter, float delta, float max_step,
                              const vo_float2* kps, const vo_float2* pred, const vo_float2* d1, const vo_float2* d2, int n, vo_float2* flow_out,
                              float* err_out) {
  switch (winsize) {
    case 5: oriented_points<5>(a, b, ag, min_ev, max_iter, delta, max_step, kps, pred, d1, d2, n, flow_out, err_out); break;
    case 7: oriented_points<7>(a, b, ag, min_ev, max_iter, delta, max_step, kps, pred, d1, d2, n, flow_out, err_out); break;
    case 9: oriented_points<9>(a, b, ag, min_ev, max_iter, delta, max_step, kps, pred, d1, d2, n, flow_out, err_out); break;
    default: oriented_points<11>(a, b, ag, min_ev, max_iter, delta, max_step, kps, pred, d1, d2, n, flow_out, err_out); break;
  }
}
}  // extern "C"
