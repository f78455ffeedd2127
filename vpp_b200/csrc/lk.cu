// Pyramidal Lucas-Kanade point matcher.
// Reference: vpp/algorithms/lucas_kanade/lucas_kanade.hpp:12-131 (lk_internals::match) and
// :135-184 (driver), vpp/algorithms/pyrlk/lk.hh:42-175 (lk_match_point_square_win<WS>),
// vpp/algorithms/pyrlk/pyrlk_match.hh:15-55 (driver), vpp/core/imageNd.hpp:280-300
// (linear_interpolate: float32 4-tap, summed left to right, result truncated to the pixel type).
//
// One warp per keypoint, all pyramid levels in one launch (the pyramids stay resident in L2:
// 1080p x 3 levels x (u8 + u8 + 8-byte gradient) ~ 27 MB).  Lanes own the window pixels
// (WS*WS <= 225 -> up to 8 per lane), sample A / grad once per level and B once per iteration.
// The reference accumulates G, b_k and the error strictly in row-major window order in float32;
// because the truncating interpolation makes the iteration discontinuous, a different summation
// order can move the result by far more than 1e-4.  The reductions therefore replay the
// reference order exactly: every lane broadcasts its terms with warp shuffles and every lane
// performs the same sequential float additions (so all lanes hold the same v, G, b_k and control
// flow stays warp-uniform).  Not HBM-bound: ~2 KB of compulsory traffic per keypoint;
// throughput is set by the 2 x WS^2 dependent FADDs per iteration.
// The library is compiled with -fmad=false; products and sums below must not be contracted.
#include "common.cuh"

#include <float.h>

namespace vppb {

constexpr int LK_MAX_LEVELS = 8;
constexpr unsigned FULL = 0xffffffffu;

struct LkLevels {
  Img prev[LK_MAX_LEVELS], next[LK_MAX_LEVELS], grad[LK_MAX_LEVELS];
};

// clamp so that the 2x2 footprint stays inside the allocated frame (the reference reads whatever
// lies there; documented deviation, never hit by in-range keypoints)
__device__ __forceinline__ int clamp_tap(int x, int n, int border) { return min(max(x, -border), n + border - 2); }

// imageNd.hpp:280-300 on image2d<unsigned char>; returns the truncated uchar as int
__device__ __forceinline__ int interp_u8(const Img& im, float p0, float p1) {
  const int x0 = (int)p0, x1 = (int)p1;
  const float a0 = __fsub_rn(p0, (float)x0), a1 = __fsub_rn(p1, (float)x1);
  const unsigned char* l1 = im.base + (long long)clamp_tap(x0, im.nrows, im.border) * im.pitch + clamp_tap(x1, im.ncols, im.border);
  const unsigned char* l2 = l1 + im.pitch;
  const float b0 = __fsub_rn(1.f, a0), b1 = __fsub_rn(1.f, a1);
  float res = __fmul_rn(__fmul_rn(b0, b1), (float)l1[0]);
  res = __fadd_rn(res, __fmul_rn(__fmul_rn(a0, b1), (float)l2[0]));
  res = __fadd_rn(res, __fmul_rn(__fmul_rn(b0, a1), (float)l1[1]));
  res = __fadd_rn(res, __fmul_rn(__fmul_rn(a0, a1), (float)l2[1]));
  // static_cast<unsigned char>(float): truncate to int, keep the low byte (what the x86 reference build does
  // for the out-of-range values that extrapolated samples can produce)
  return ((int)res) & 255;
}

// The same sample for the inner loop of the v2 kernel: the image's fields live in registers for the whole level (the Img structs sit
// in the parameter bank behind a run-time level index: three constant loads per sample otherwise), the clamp bounds are precomputed,
// and the byte -> float conversions are the full-rate I2FP of the integer pipe rather than the quarter-rate I2F.U16 / I2F.U8.
struct ImgU8 {
  const unsigned char* base;
  int pitch, lo, hi_r, hi_c, nrows, ncols;
};
__device__ __forceinline__ ImgU8 sampler_of(const Img& im) {
  ImgU8 s;
  s.base = im.base; s.pitch = im.pitch; s.lo = -im.border; s.hi_r = im.nrows + im.border - 2; s.hi_c = im.ncols + im.border - 2;
  s.nrows = im.nrows; s.ncols = im.ncols;
  return s;
}
// (float)v for v < 2^23 without the conversion unit: 2^23 + v is exactly representable, subtract 2^23 (LOP3 + FADD, both full rate;
// the compiler turns a plain cast of a loaded byte into the quarter-rate I2F.U16)
__device__ __forceinline__ float byte_to_float(unsigned v) { return __fsub_rn(__uint_as_float(0x4B000000u | v), 8388608.f); }
__device__ __forceinline__ int interp_u8(const ImgU8& im, float p0, float p1) {
  const int x0 = (int)p0, x1 = (int)p1;
  const float a0 = __fsub_rn(p0, (float)x0), a1 = __fsub_rn(p1, (float)x1);
  // 32-bit offset from pixel (0,0): the host only takes this kernel for images whose frame stays below 2 GB
  const unsigned char* l1 = im.base + (min(max(x0, im.lo), im.hi_r) * im.pitch + min(max(x1, im.lo), im.hi_c));
  const unsigned char* l2 = l1 + im.pitch;
  const float b0 = __fsub_rn(1.f, a0), b1 = __fsub_rn(1.f, a1);
  float res = __fmul_rn(__fmul_rn(b0, b1), byte_to_float(l1[0]));
  res = __fadd_rn(res, __fmul_rn(__fmul_rn(a0, b1), byte_to_float(l2[0])));
  res = __fadd_rn(res, __fmul_rn(__fmul_rn(b0, a1), byte_to_float(l1[1])));
  res = __fadd_rn(res, __fmul_rn(__fmul_rn(a0, a1), byte_to_float(l2[1])));
  return ((int)res) & 255;
}
// the sample as a float (the truncated uchar converted back, as `cast<float>(B.linear_interpolate(n))` does)
__device__ __forceinline__ float interp_u8f(const ImgU8& im, float p0, float p1) { return byte_to_float((unsigned)interp_u8(im, p0, p1)); }

// same on image2d<vint2> (result truncated per component) or image2d<vfloat2>
template <bool GRAD_FLOAT>
__device__ __forceinline__ float2 interp_grad(const Img& im, float p0, float p1) {
  const int x0 = (int)p0, x1 = (int)p1;
  const float a0 = __fsub_rn(p0, (float)x0), a1 = __fsub_rn(p1, (float)x1);
  const unsigned char* l1 = im.base + (long long)clamp_tap(x0, im.nrows, im.border) * im.pitch + (long long)clamp_tap(x1, im.ncols, im.border) * 8;
  const unsigned char* l2 = l1 + im.pitch;
  float v00x, v00y, v10x, v10y, v01x, v01y, v11x, v11y;
  if (GRAD_FLOAT) {
    const float2 t00 = *reinterpret_cast<const float2*>(l1), t01 = *reinterpret_cast<const float2*>(l1 + 8);
    const float2 t10 = *reinterpret_cast<const float2*>(l2), t11 = *reinterpret_cast<const float2*>(l2 + 8);
    v00x = t00.x; v00y = t00.y; v01x = t01.x; v01y = t01.y; v10x = t10.x; v10y = t10.y; v11x = t11.x; v11y = t11.y;
  } else {
    const int2 t00 = *reinterpret_cast<const int2*>(l1), t01 = *reinterpret_cast<const int2*>(l1 + 8);
    const int2 t10 = *reinterpret_cast<const int2*>(l2), t11 = *reinterpret_cast<const int2*>(l2 + 8);
    v00x = (float)t00.x; v00y = (float)t00.y; v01x = (float)t01.x; v01y = (float)t01.y;
    v10x = (float)t10.x; v10y = (float)t10.y; v11x = (float)t11.x; v11y = (float)t11.y;
  }
  const float b0 = __fsub_rn(1.f, a0), b1 = __fsub_rn(1.f, a1);
  const float w00 = __fmul_rn(b0, b1), w10 = __fmul_rn(a0, b1), w01 = __fmul_rn(b0, a1), w11 = __fmul_rn(a0, a1);
  float rx = __fmul_rn(w00, v00x), ry = __fmul_rn(w00, v00y);
  rx = __fadd_rn(rx, __fmul_rn(w10, v10x)); ry = __fadd_rn(ry, __fmul_rn(w10, v10y));
  rx = __fadd_rn(rx, __fmul_rn(w01, v01x)); ry = __fadd_rn(ry, __fmul_rn(w01, v01y));
  rx = __fadd_rn(rx, __fmul_rn(w11, v11x)); ry = __fadd_rn(ry, __fmul_rn(w11, v11y));
  if (!GRAD_FLOAT) { rx = (float)(int)rx; ry = (float)(int)ry; }  // cast<vint2> then gx = g[0] (int -> float)
  return make_float2(rx, ry);
}

// Sum t[0..npix) in window order with the reference's sequential float additions.
template <int PPL>
__device__ __forceinline__ float seq_sum(const float (&t)[PPL], int npix) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < PPL; j++) {
    const int cnt = min(32, npix - 32 * j);
#pragma unroll 8
    for (int l = 0; l < cnt; l++) s = __fadd_rn(s, __shfl_sync(FULL, t[j], l));
  }
  return s;
}
template <int PPL>
__device__ __forceinline__ void seq_sum2(const float (&t0)[PPL], const float (&t1)[PPL], int npix, float& s0, float& s1) {
  s0 = 0.f; s1 = 0.f;
#pragma unroll
  for (int j = 0; j < PPL; j++) {
    const int cnt = min(32, npix - 32 * j);
#pragma unroll 8
    for (int l = 0; l < cnt; l++) {
      s0 = __fadd_rn(s0, __shfl_sync(FULL, t0[j], l));
      s1 = __fadd_rn(s1, __shfl_sync(FULL, t1[j], l));
    }
  }
}

__device__ __forceinline__ bool finite2(float a, float b) { return isfinite(a) && isfinite(b); }

template <int PPL, bool GRAD_FLOAT>
__global__ void __launch_bounds__(128) k_lk_match(LkLevels L, vppb_lk_params P, const vppb_float2* __restrict__ kps,
                                                  const vppb_float2* __restrict__ prediction, int n, vppb_float2* __restrict__ flow_out,
                                                  float* __restrict__ err_out) {
  const int lane = threadIdx.x & 31;
  const int kp_idx = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (kp_idx >= n) return;  // warp-uniform
  const int ws = P.winsize, hws = ws / 2, npix = ws * ws;
  const float kp0 = kps[kp_idx].r, kp1 = kps[kp_idx].c;

  // window offsets of this lane's pixels
  float off_r[PPL], off_c[PPL];
  bool have[PPL];
#pragma unroll
  for (int j = 0; j < PPL; j++) {
    const int i = lane + 32 * j;
    have[j] = i < npix;
    off_r[j] = (float)(i / ws - hws);
    off_c[j] = (float)(i % ws - hws);
  }

  float tr0 = 0.f, tr1 = 0.f, dist = 0.f;
  if (prediction) {  // lucas_kanade.hpp:163
    tr0 = __fdiv_rn(prediction[kp_idx].r, P.pred_div);
    tr1 = __fdiv_rn(prediction[kp_idx].c, P.pred_div);
  }

  for (int S = P.nlevels - 1; S >= P.min_scale; S--) {
    tr0 = __fmul_rn(tr0, P.factor);
    tr1 = __fmul_rn(tr1, P.factor);
    const Img& A = L.prev[S];
    const Img& B = L.next[S];
    const Img& Ag = L.grad[S];
    const float scale = (float)(1 << S);
    const float p0 = __fdiv_rn(kp0, scale), p1 = __fdiv_rn(kp1, scale);

    // ---- G and the cached samples (lucas_kanade.hpp:24-43, 69-83)
    float gs0[PPL], gs1[PPL], asv[PPL], t00[PPL], t01[PPL], t11[PPL];
    bool valid[PPL];
    int cpt = 0;
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      const float n0 = __fadd_rn(p0, off_r[j]), n1 = __fadd_rn(p1, off_c[j]);
      const int i0 = (int)n0, i1 = (int)n1;
      valid[j] = have[j] && i0 >= 0 && i0 < A.nrows && i1 >= 0 && i1 < A.ncols;
      gs0[j] = 0.f; gs1[j] = 0.f; asv[j] = 0.f;
      if (valid[j]) {
        const float2 g = interp_grad<GRAD_FLOAT>(Ag, n0, n1);
        gs0[j] = g.x; gs1[j] = g.y;
        asv[j] = (float)interp_u8(A, n0, n1);
      }
      t00[j] = __fmul_rn(gs0[j], gs0[j]);
      t01[j] = __fmul_rn(gs0[j], gs1[j]);
      t11[j] = __fmul_rn(gs1[j], gs1[j]);
      cpt += __popc(__ballot_sync(FULL, valid[j]));
    }
    float G00, G01, G11;
    seq_sum2<PPL>(t00, t01, npix, G00, G01);
    G11 = seq_sum<PPL>(t11, npix);

    float m0 = -1.f, m1 = -1.f, merr = FLT_MAX;  // result of this level's match
    bool done = false;

    // ---- minimum eigenvalue of G / cpt (lucas_kanade.hpp:45-52); closed form for a symmetric 2x2
    {
      const float cf = (float)cpt;
      const float a = __fdiv_rn(G00, cf), b = __fdiv_rn(G01, cf), d = __fdiv_rn(G11, cf);
      const float half = __fmul_rn(__fadd_rn(a, d), 0.5f), diff = __fmul_rn(__fsub_rn(a, d), 0.5f);
      const float root = __fsqrt_rn(__fadd_rn(__fmul_rn(diff, diff), __fmul_rn(b, b)));
      const float e1 = fabsf(__fadd_rn(half, root)), e2 = fabsf(__fsub_rn(half, root));
      float min_ev = 99999.f;
      if (e1 < min_ev) min_ev = e1;
      if (e2 < min_ev) min_ev = e2;
      if (min_ev < P.min_ev) { m0 = -1.f; m1 = -1.f; merr = FLT_MAX; done = true; }
    }

    if (!done) {
      // G^-1 (Eigen 2x2: invdet = 1/det; [d,-b;-c,a] * invdet)
      const float det = __fsub_rn(__fmul_rn(G00, G11), __fmul_rn(G01, G01));
      const float invdet = __fdiv_rn(1.f, det);
      const float I00 = __fmul_rn(G11, invdet), I01 = __fmul_rn(-G01, invdet), I11 = __fmul_rn(G00, invdet);

      float v0 = __fadd_rn(p0, tr0), v1 = __fadd_rn(p1, tr1);
      float nk0 = 1.f, nk1 = 1.f;
      bool failed = false;
      for (int k = 0; k <= P.max_iter; k++) {
        const float nrm = __fsqrt_rn(__fadd_rn(__fmul_rn(nk0, nk0), __fmul_rn(nk1, nk1)));
        if (!(nrm >= P.delta)) break;
        float c0[PPL], c1[PPL];
#pragma unroll
        for (int j = 0; j < PPL; j++) {
          c0[j] = 0.f; c1[j] = 0.f;
          if (valid[j]) {
            const float dt = __fsub_rn(asv[j], (float)interp_u8(B, __fadd_rn(v0, off_r[j]), __fadd_rn(v1, off_c[j])));
            c0[j] = __fmul_rn(gs0[j], dt);
            c1[j] = __fmul_rn(gs1[j], dt);
          }
        }
        float bk0, bk1;
        seq_sum2<PPL>(c0, c1, npix, bk0, bk1);
        nk0 = __fadd_rn(__fmul_rn(I00, bk0), __fmul_rn(I01, bk1));
        nk1 = __fadd_rn(__fmul_rn(I01, bk0), __fmul_rn(I11, bk1));
        v0 = __fadd_rn(v0, nk0);
        v1 = __fadd_rn(v1, nk1);
        const int iv0 = (int)v0, iv1 = (int)v1;
        if (!finite2(v0, v1) || iv0 < 0 || iv0 >= B.nrows || iv1 < 0 || iv1 >= B.ncols) { failed = true; break; }
      }
      if (failed) {
        m0 = 0.f; m1 = 0.f; merr = FLT_MAX;
      } else {
        // ---- matching error (lucas_kanade.hpp:116-128 / lk.hh:151-173)
        float e[PPL];
#pragma unroll
        for (int j = 0; j < PPL; j++) {
          e[j] = 0.f;
          if (have[j]) {
            const int bi = interp_u8(B, __fadd_rn(v0, off_r[j]), __fadd_rn(v1, off_c[j]));
            e[j] = fabsf((float)((int)asv[j] - bi));
          }
        }
        float err = seq_sum<PPL>(e, npix);
        const int cpt2 = cpt + npix;
        if (P.err_mode == VPPB_LK_ERR_SAD) {
          merr = __fdiv_rn(err, (float)cpt2);
        } else {
          const float wsq = (float)npix;
          float avg = __fdiv_rn(seq_sum<PPL>(asv, npix), wsq);
          float dv[PPL];
#pragma unroll
          for (int j = 0; j < PPL; j++) dv[j] = have[j] ? fabsf(__fsub_rn(avg, asv[j])) : 0.f;
          const float stddev = __fdiv_rn(seq_sum<PPL>(dv, npix), wsq);
          merr = __fdiv_rn(err, __fmul_rn((float)cpt2, stddev));
        }
        m0 = __fsub_rn(v0, p0);
        m1 = __fsub_rn(v1, p1);
      }
    }

    // ---- drivers: lucas_kanade.hpp:177-178 (unconditional) / pyrlk_match.hh:37-41 (gated)
    if (!P.gate_on_max_err || merr < P.max_err) { tr0 = m0; tr1 = m1; }
    dist = merr;
  }

  if (lane == 0) {
    flow_out[kp_idx].r = tr0;
    flow_out[kp_idx].c = tr1;
    err_out[kp_idx] = dist;
  }
}

// ------------------------------------------------------------------------------------------------
// v2: four keypoints per warp (8 lanes each).  The window samples are spread over the 8 lanes of a
// keypoint; the per-pixel terms go to shared memory and ONE lane per (keypoint, component) adds them
// in the reference's window order.  Same float sequence as k_lk_match (bit-identical results), but
// the 2 x WS^2 dependent additions are issued once per keypoint instead of once per lane, which cuts
// the warp-instruction count per keypoint ~3.5x.  Used for WS <= 11.
constexpr int LK2_LPK = 8;    // lanes per keypoint
constexpr int LK2_KPW = 4;    // keypoints per warp
constexpr int LK2_WARPS = 4;  // warps per CTA (default; the kernel is also instantiated for 1 and 2: see lk_launch_v2)
constexpr int LK2_MAXPIX = 121;
constexpr int LK2_ROW = 124;

template <int PPL, bool GRAD_FLOAT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_lk_match_v2(LkLevels L, vppb_lk_params P, const vppb_float2* __restrict__ kps,
                                                             const vppb_float2* __restrict__ prediction, int n,
                                                             vppb_float2* __restrict__ flow_out, float* __restrict__ err_out) {
  __align__(16) __shared__ float sbuf[WARPS][3][LK2_KPW][LK2_ROW];  // rows of 124 floats: 16-byte aligned, 28 banks apart (no conflicts between the 8 summing lanes)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = lane >> 3, sl = lane & 7, lead = lane & ~7;
  const int kp_idx = (blockIdx.x * WARPS + warp) * LK2_KPW + k;
  const bool kp_ok = kp_idx < n;
  float (*buf)[LK2_KPW][LK2_ROW] = sbuf[warp];
  const int ws = P.winsize, hws = ws / 2, npix = ws * ws;
  const int kload = kp_ok ? kp_idx : 0;
  const float kp0 = kps[kload].r, kp1 = kps[kload].c;

  float off_r[PPL], off_c[PPL];
  bool have[PPL];
#pragma unroll
  for (int j = 0; j < PPL; j++) {
    const int i = sl + LK2_LPK * j;
    have[j] = i < npix;
    off_r[j] = (float)(i / ws - hws);
    off_c[j] = (float)(i % ws - hws);
  }
  // ordered sums of `ncomp` components of keypoint k: lane sl (< ncomp) adds buf[sl][k][0..npix) in order
  auto ordered_sum = [&](int ncomp) -> float {
    __syncwarp();
    float acc = 0.f;
    if (sl < ncomp) {
      const float* src = buf[sl][k];
      int i = 0;
#pragma unroll 3
      for (; i + 4 <= npix; i += 4) {  // four terms per shared-memory load, added in window order
        const float4 q = *reinterpret_cast<const float4*>(src + i);
        acc = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc, q.x), q.y), q.z), q.w);
      }
      for (; i < npix; i++) acc = __fadd_rn(acc, src[i]);
    }
    __syncwarp();
    return acc;
  };

  float tr0 = 0.f, tr1 = 0.f, dist = 0.f;
  if (prediction) {
    tr0 = __fdiv_rn(prediction[kload].r, P.pred_div);
    tr1 = __fdiv_rn(prediction[kload].c, P.pred_div);
  }

  for (int S = P.nlevels - 1; S >= P.min_scale; S--) {
    tr0 = __fmul_rn(tr0, P.factor);
    tr1 = __fmul_rn(tr1, P.factor);
    const Img& A = L.prev[S];
    const Img& B = L.next[S];
    const Img& Ag = L.grad[S];
    const float scale = (float)(1 << S);
    const float p0 = __fdiv_rn(kp0, scale), p1 = __fdiv_rn(kp1, scale);

    const ImgU8 sA = sampler_of(A), sB = sampler_of(B);
    float gs0[PPL], gs1[PPL], asv[PPL];
    bool valid[PPL];
    int cpt_l = 0;
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      const float n0 = __fadd_rn(p0, off_r[j]), n1 = __fadd_rn(p1, off_c[j]);
      const int i0 = (int)n0, i1 = (int)n1;
      valid[j] = have[j] && i0 >= 0 && i0 < A.nrows && i1 >= 0 && i1 < A.ncols;
      gs0[j] = 0.f; gs1[j] = 0.f; asv[j] = 0.f;
      if (valid[j]) {
        const float2 g = interp_grad<GRAD_FLOAT>(Ag, n0, n1);
        gs0[j] = g.x; gs1[j] = g.y;
        asv[j] = interp_u8f(sA, n0, n1);
        cpt_l++;
      }
      if (have[j]) {
        const int i = sl + LK2_LPK * j;
        buf[0][k][i] = __fmul_rn(gs0[j], gs0[j]);
        buf[1][k][i] = __fmul_rn(gs0[j], gs1[j]);
        buf[2][k][i] = __fmul_rn(gs1[j], gs1[j]);
      }
    }
    // cpt: number of in-domain window pixels of this keypoint (sum over its 8 lanes)
    int cpt = cpt_l;
    cpt += __shfl_xor_sync(FULL, cpt, 1); cpt += __shfl_xor_sync(FULL, cpt, 2); cpt += __shfl_xor_sync(FULL, cpt, 4);
    const float gsum = ordered_sum(3);
    const float G00 = __shfl_sync(FULL, gsum, lead), G01 = __shfl_sync(FULL, gsum, lead + 1), G11 = __shfl_sync(FULL, gsum, lead + 2);

    float m0 = -1.f, m1 = -1.f, merr = FLT_MAX;
    bool active = true;  // this keypoint is still iterating at this level
    {
      const float cf = (float)cpt;
      const float a = __fdiv_rn(G00, cf), b = __fdiv_rn(G01, cf), d = __fdiv_rn(G11, cf);
      const float half = __fmul_rn(__fadd_rn(a, d), 0.5f), diff = __fmul_rn(__fsub_rn(a, d), 0.5f);
      const float root = __fsqrt_rn(__fadd_rn(__fmul_rn(diff, diff), __fmul_rn(b, b)));
      const float e1 = fabsf(__fadd_rn(half, root)), e2 = fabsf(__fsub_rn(half, root));
      float min_ev = 99999.f;
      if (e1 < min_ev) min_ev = e1;
      if (e2 < min_ev) min_ev = e2;
      if (min_ev < P.min_ev) active = false;  // result stays ((-1,-1), FLT_MAX)
    }
    const bool rejected = !active;
    const float det = __fsub_rn(__fmul_rn(G00, G11), __fmul_rn(G01, G01));
    const float invdet = __fdiv_rn(1.f, det);
    const float I00 = __fmul_rn(G11, invdet), I01 = __fmul_rn(-G01, invdet), I11 = __fmul_rn(G00, invdet);
    float v0 = __fadd_rn(p0, tr0), v1 = __fadd_rn(p1, tr1);
    float nk0 = 1.f, nk1 = 1.f;
    bool failed = false;
    for (int kk = 0; kk <= P.max_iter; kk++) {
      if (active) {
        const float nrm = __fsqrt_rn(__fadd_rn(__fmul_rn(nk0, nk0), __fmul_rn(nk1, nk1)));
        if (!(nrm >= P.delta)) active = false;
      }
      if (!__any_sync(FULL, active)) break;
#pragma unroll
      for (int j = 0; j < PPL; j++) {
        if (have[j]) {
          float c0 = 0.f, c1 = 0.f;
          if (valid[j] && active) {
            const float dt = __fsub_rn(asv[j], interp_u8f(sB, __fadd_rn(v0, off_r[j]), __fadd_rn(v1, off_c[j])));
            c0 = __fmul_rn(gs0[j], dt);
            c1 = __fmul_rn(gs1[j], dt);
          }
          const int i = sl + LK2_LPK * j;
          buf[0][k][i] = c0;
          buf[1][k][i] = c1;
        }
      }
      const float bsum = ordered_sum(2);
      const float bk0 = __shfl_sync(FULL, bsum, lead), bk1 = __shfl_sync(FULL, bsum, lead + 1);
      if (active) {
        nk0 = __fadd_rn(__fmul_rn(I00, bk0), __fmul_rn(I01, bk1));
        nk1 = __fadd_rn(__fmul_rn(I01, bk0), __fmul_rn(I11, bk1));
        v0 = __fadd_rn(v0, nk0);
        v1 = __fadd_rn(v1, nk1);
        const int iv0 = (int)v0, iv1 = (int)v1;
        if (!finite2(v0, v1) || iv0 < 0 || iv0 >= sB.nrows || iv1 < 0 || iv1 >= sB.ncols) { failed = true; active = false; }
      }
    }
    // ---- matching error for the keypoints that neither were rejected nor left the domain
    const bool want_err = !rejected && !failed;
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      if (have[j]) {
        float e = 0.f;
        if (want_err) {
          const int bi = interp_u8(sB, __fadd_rn(v0, off_r[j]), __fadd_rn(v1, off_c[j]));
          e = fabsf((float)((int)asv[j] - bi));
        }
        const int i = sl + LK2_LPK * j;
        buf[0][k][i] = e;
        buf[1][k][i] = asv[j];
      }
    }
    const float esum = ordered_sum(2);
    const float err = __shfl_sync(FULL, esum, lead), asum = __shfl_sync(FULL, esum, lead + 1);
    const int cpt2 = cpt + npix;
    float stddev = 1.f;
    if (P.err_mode != VPPB_LK_ERR_SAD) {  // warp-uniform
      const float avg = __fdiv_rn(asum, (float)npix);
#pragma unroll
      for (int j = 0; j < PPL; j++)
        if (have[j]) buf[0][k][sl + LK2_LPK * j] = fabsf(__fsub_rn(avg, asv[j]));
      const float dsum = ordered_sum(1);
      stddev = __fdiv_rn(__shfl_sync(FULL, dsum, lead), (float)npix);
    }
    if (rejected) { m0 = -1.f; m1 = -1.f; merr = FLT_MAX; }
    else if (failed) { m0 = 0.f; m1 = 0.f; merr = FLT_MAX; }
    else {
      merr = P.err_mode == VPPB_LK_ERR_SAD ? __fdiv_rn(err, (float)cpt2) : __fdiv_rn(err, __fmul_rn((float)cpt2, stddev));
      m0 = __fsub_rn(v0, p0);
      m1 = __fsub_rn(v1, p1);
    }
    if (!P.gate_on_max_err || merr < P.max_err) { tr0 = m0; tr1 = m1; }
    dist = merr;
  }
  if (sl == 0 && kp_ok) {
    flow_out[kp_idx].r = tr0;
    flow_out[kp_idx].c = tr1;
    err_out[kp_idx] = dist;
  }
}

// Warps per CTA: the whole problem is one wave of resident CTAs and the kernel is issue-bound, so the launch lasts as long as the SM that
// holds the most warps.  10 000 keypoints are 2 500 warps: in CTAs of 4 warps (625 CTAs over 148 SMs) the fullest SM holds 5 x 4 = 20 warps
// against an average of 16.9; single-warp CTAs level that out.  VPPB_LK_WARPS=1|2|4 selects the instantiation (A/B timing).
template <bool GF, int WARPS>
static bool lk_launch_v2w(int winsize, cudaStream_t st, const LkLevels& L, const vppb_lk_params& P, const vppb_float2* kps, const vppb_float2* pred, int n,
                          vppb_float2* flow, float* err) {
  const int per_cta = WARPS * LK2_KPW;
  const int grid = (n + per_cta - 1) / per_cta;
  switch (winsize) {
    case 1: case 3: case 5: k_lk_match_v2<4, GF, WARPS><<<grid, WARPS * 32, 0, st>>>(L, P, kps, pred, n, flow, err); return true;
    case 7: k_lk_match_v2<7, GF, WARPS><<<grid, WARPS * 32, 0, st>>>(L, P, kps, pred, n, flow, err); return true;
    case 9: k_lk_match_v2<11, GF, WARPS><<<grid, WARPS * 32, 0, st>>>(L, P, kps, pred, n, flow, err); return true;
    case 11: k_lk_match_v2<16, GF, WARPS><<<grid, WARPS * 32, 0, st>>>(L, P, kps, pred, n, flow, err); return true;
    default: return false;
  }
}
template <bool GF>
static bool lk_launch_v2(int winsize, cudaStream_t st, const LkLevels& L, const vppb_lk_params& P, const vppb_float2* kps,
                         const vppb_float2* pred, int n, vppb_float2* flow, float* err) {
  const char* e = getenv("VPPB_LK_WARPS");
  const int w = e ? atoi(e) : LK2_WARPS;
  if (w == 1) return lk_launch_v2w<GF, 1>(winsize, st, L, P, kps, pred, n, flow, err);
  if (w == 2) return lk_launch_v2w<GF, 2>(winsize, st, L, P, kps, pred, n, flow, err);
  return lk_launch_v2w<GF, 4>(winsize, st, L, P, kps, pred, n, flow, err);
}

template <bool GF>
static void lk_launch(int ppl, int grid, cudaStream_t st, const LkLevels& L, const vppb_lk_params& P, const vppb_float2* kps,
                      const vppb_float2* pred, int n, vppb_float2* flow, float* err) {
  switch (ppl) {
    case 1: k_lk_match<1, GF><<<grid, 128, 0, st>>>(L, P, kps, pred, n, flow, err); break;
    case 2: k_lk_match<2, GF><<<grid, 128, 0, st>>>(L, P, kps, pred, n, flow, err); break;
    case 3: k_lk_match<3, GF><<<grid, 128, 0, st>>>(L, P, kps, pred, n, flow, err); break;
    case 4: k_lk_match<4, GF><<<grid, 128, 0, st>>>(L, P, kps, pred, n, flow, err); break;
    case 6: k_lk_match<6, GF><<<grid, 128, 0, st>>>(L, P, kps, pred, n, flow, err); break;
    default: k_lk_match<8, GF><<<grid, 128, 0, st>>>(L, P, kps, pred, n, flow, err); break;
  }
}

// ------------------------------------------------------------------------------------------------
// oriented_lk_match_point_square_win<WS> (lk.hh:180-317), one level.  One warp per keypoint, the same ordered float sums as
// k_lk_match.  Differences from the square-window matcher: the gradient matrix is summed over the axis-aligned window but the
// template is sampled on a window rotated to dir1 and the search window is rotated to dir2; the eigenvalue test uses G itself
// (lk.hh:219); steps longer than max_step are shortened to it (lk.hh:276-280); at most max_iter steps (k < max_iter, lk.hh:258);
// the search is confined to B's domain shrunk by 3 pixels (lk.hh:251,283); error = SAD / (cpt * MAD) (lk.hh:288-314).
struct LkOriented {
  Img A, B, Ag;
  int winsize, max_iter;
  float min_ev, delta, max_step;
};

template <int PPL, bool GRAD_FLOAT>
__global__ void __launch_bounds__(128) k_lk_match_oriented(LkOriented P, const vppb_float2* __restrict__ kps, const vppb_float2* __restrict__ prediction,
                                                           const vppb_float2* __restrict__ dir1, const vppb_float2* __restrict__ dir2, int n,
                                                           vppb_float2* __restrict__ flow_out, float* __restrict__ err_out) {
  const int lane = threadIdx.x & 31;
  const int q = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (q >= n) return;  // warp-uniform
  const int ws = P.winsize, hws = ws / 2, npix = ws * ws;
  const Img& A = P.A;
  const Img& B = P.B;
  const Img& Ag = P.Ag;
  const float p0 = kps[q].r, p1 = kps[q].c;
  float off_r[PPL], off_c[PPL];
  bool have[PPL];
#pragma unroll
  for (int j = 0; j < PPL; j++) {
    const int i = lane + 32 * j;
    have[j] = i < npix;
    off_r[j] = (float)(i / ws - hws);
    off_c[j] = (float)(i % ws - hws);
  }
  // gradient matrix over the axis-aligned window (lk.hh:198-216)
  float t00[PPL], t01[PPL], t11[PPL];
  int cpt = 0;
#pragma unroll
  for (int j = 0; j < PPL; j++) {
    const float n0 = __fadd_rn(p0, off_r[j]), n1 = __fadd_rn(p1, off_c[j]);
    const int i0 = (int)n0, i1 = (int)n1;
    const bool ok = have[j] && i0 >= 0 && i0 < A.nrows && i1 >= 0 && i1 < A.ncols;
    float2 g = make_float2(0.f, 0.f);
    if (ok) g = interp_grad<GRAD_FLOAT>(Ag, n0, n1);
    t00[j] = __fmul_rn(g.x, g.x); t01[j] = __fmul_rn(g.x, g.y); t11[j] = __fmul_rn(g.y, g.y);
    cpt += __popc(__ballot_sync(FULL, ok));
  }
  float G00, G01, G11;
  seq_sum2<PPL>(t00, t01, npix, G00, G01);
  G11 = seq_sum<PPL>(t11, npix);
  float m0 = -1.f, m1 = -1.f, merr = FLT_MAX;
  bool done = false;
  {
    const float half = __fmul_rn(__fadd_rn(G00, G11), 0.5f), diff = __fmul_rn(__fsub_rn(G00, G11), 0.5f);
    const float root = __fsqrt_rn(__fadd_rn(__fmul_rn(diff, diff), __fmul_rn(G01, G01)));
    const float e1 = fabsf(__fadd_rn(half, root)), e2 = fabsf(__fsub_rn(half, root));
    float min_ev = 99999.f;
    if (e1 < min_ev) min_ev = e1;
    if (e2 < min_ev) min_ev = e2;
    if (min_ev < P.min_ev) done = true;
  }
  if (!done) {
    const float det = __fsub_rn(__fmul_rn(G00, G11), __fmul_rn(G01, G01));
    const float invdet = __fdiv_rn(1.f, det);
    const float I00 = __fmul_rn(G11, invdet), I01 = __fmul_rn(-G01, invdet), I11 = __fmul_rn(G00, invdet);
    float v0 = __fadd_rn(p0, prediction[q].r), v1 = __fadd_rn(p1, prediction[q].c);
    float nk0 = 1.f, nk1 = 1.f;
    // template on the window rotated to dir1: columns along mx = dir1, rows along my = (-mx[1], mx[0]) (lk.hh:231-249)
    float mx0 = dir1[q].r, mx1 = dir1[q].c, my0 = -mx1, my1 = mx0;
    float gs0[PPL], gs1[PPL], asv[PPL];
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      const float n0 = __fadd_rn(p0, __fadd_rn(__fmul_rn(off_r[j], my0), __fmul_rn(off_c[j], mx0)));
      const float n1 = __fadd_rn(p1, __fadd_rn(__fmul_rn(off_r[j], my1), __fmul_rn(off_c[j], mx1)));
      const int i0 = (int)n0, i1 = (int)n1;
      gs0[j] = 0.f; gs1[j] = 0.f; asv[j] = 0.f;
      if (have[j] && i0 >= 0 && i0 < Ag.nrows && i1 >= 0 && i1 < Ag.ncols) {
        const float2 g = interp_grad<GRAD_FLOAT>(Ag, n0, n1);
        gs0[j] = g.x; gs1[j] = g.y;
        asv[j] = (float)interp_u8(A, n0, n1);
      }
    }
    mx0 = dir2[q].r; mx1 = dir2[q].c; my0 = -mx1; my1 = mx0;
    bool failed = false;
    for (int k = 0; k < P.max_iter; k++) {
      const float nrm = __fsqrt_rn(__fadd_rn(__fmul_rn(nk0, nk0), __fmul_rn(nk1, nk1)));
      if (!(nrm >= P.delta)) break;
      float c0[PPL], c1[PPL];
#pragma unroll
      for (int j = 0; j < PPL; j++) {
        c0[j] = 0.f; c1[j] = 0.f;
        if (have[j]) {
          const float n0 = __fadd_rn(v0, __fadd_rn(__fmul_rn(off_r[j], my0), __fmul_rn(off_c[j], mx0)));
          const float n1 = __fadd_rn(v1, __fadd_rn(__fmul_rn(off_r[j], my1), __fmul_rn(off_c[j], mx1)));
          const float dt = __fsub_rn(asv[j], (float)interp_u8(B, n0, n1));
          c0[j] = __fmul_rn(gs0[j], dt);
          c1[j] = __fmul_rn(gs1[j], dt);
        }
      }
      float bk0, bk1;
      seq_sum2<PPL>(c0, c1, npix, bk0, bk1);
      nk0 = __fadd_rn(__fmul_rn(I00, bk0), __fmul_rn(I01, bk1));
      nk1 = __fadd_rn(__fmul_rn(I01, bk0), __fmul_rn(I11, bk1));
      const float nn = __fsqrt_rn(__fadd_rn(__fmul_rn(nk0, nk0), __fmul_rn(nk1, nk1)));
      if (nn > P.max_step) {
        nk0 = __fmul_rn(__fdiv_rn(nk0, nn), P.max_step);
        nk1 = __fmul_rn(__fdiv_rn(nk1, nn), P.max_step);
      }
      v0 = __fadd_rn(v0, nk0);
      v1 = __fadd_rn(v1, nk1);
      const int iv0 = (int)v0, iv1 = (int)v1;
      if (!finite2(v0, v1) || iv0 < 3 || iv0 > B.nrows - 4 || iv1 < 3 || iv1 > B.ncols - 4) { failed = true; break; }
    }
    if (failed) {
      m0 = 0.f; m1 = 0.f; merr = FLT_MAX;
    } else {
      const float wsq = (float)npix;
      const float avg = __fdiv_rn(seq_sum<PPL>(asv, npix), wsq);
      float dv[PPL], e[PPL];
#pragma unroll
      for (int j = 0; j < PPL; j++) {
        dv[j] = have[j] ? fabsf(__fsub_rn(avg, asv[j])) : 0.f;
        e[j] = 0.f;
        if (have[j]) {
          const float n0 = __fadd_rn(v0, __fadd_rn(__fmul_rn(off_r[j], my0), __fmul_rn(off_c[j], mx0)));
          const float n1 = __fadd_rn(v1, __fadd_rn(__fmul_rn(off_r[j], my1), __fmul_rn(off_c[j], mx1)));
          e[j] = fabsf((float)((int)asv[j] - interp_u8(B, n0, n1)));
        }
      }
      const float stddev = __fdiv_rn(seq_sum<PPL>(dv, npix), wsq);
      const float err = seq_sum<PPL>(e, npix);
      merr = __fdiv_rn(err, __fmul_rn((float)(cpt + npix), stddev));
      m0 = __fsub_rn(v0, p0);
      m1 = __fsub_rn(v1, p1);
    }
  }
  if (lane == 0) {
    flow_out[q].r = m0;
    flow_out[q].c = m1;
    err_out[q] = merr;
  }
}

template <bool GF>
static void lk_oriented_launch(int ppl, int grid, cudaStream_t st, const LkOriented& P, const vppb_float2* kps, const vppb_float2* pred,
                               const vppb_float2* d1, const vppb_float2* d2, int n, vppb_float2* flow, float* err) {
  switch (ppl) {
    case 1: k_lk_match_oriented<1, GF><<<grid, 128, 0, st>>>(P, kps, pred, d1, d2, n, flow, err); break;
    case 2: k_lk_match_oriented<2, GF><<<grid, 128, 0, st>>>(P, kps, pred, d1, d2, n, flow, err); break;
    case 3: k_lk_match_oriented<3, GF><<<grid, 128, 0, st>>>(P, kps, pred, d1, d2, n, flow, err); break;
    case 4: k_lk_match_oriented<4, GF><<<grid, 128, 0, st>>>(P, kps, pred, d1, d2, n, flow, err); break;
    case 6: k_lk_match_oriented<6, GF><<<grid, 128, 0, st>>>(P, kps, pred, d1, d2, n, flow, err); break;
    default: k_lk_match_oriented<8, GF><<<grid, 128, 0, st>>>(P, kps, pred, d1, d2, n, flow, err); break;
  }
}

}  // namespace vppb

using namespace vppb;

extern "C" {

int vppb_lk_match_u8(const vppb_img* prev, const vppb_img* next, const vppb_img* grad, const vppb_lk_params* params,
                     const vppb_float2* kps, const vppb_float2* prediction, int32_t n, vppb_float2* flow_out, float* err_out,
                     void* stream) {
  VPPB_REQUIRE(prev && next && grad && params, VPPB_E_ARG, "vppb_lk_match_u8: NULL argument");
  VPPB_REQUIRE(n == 0 || (kps && flow_out && err_out), VPPB_E_ARG, "vppb_lk_match_u8: NULL keypoint/output array");
  const vppb_lk_params& P = *params;
  VPPB_REQUIRE(P.nlevels >= 1 && P.nlevels <= LK_MAX_LEVELS && P.min_scale >= 0 && P.min_scale < P.nlevels, VPPB_E_ARG,
               "vppb_lk_match_u8: nlevels %d / min_scale %d out of range", P.nlevels, P.min_scale);
  VPPB_REQUIRE(P.winsize >= 1 && P.winsize <= 15 && (P.winsize & 1), VPPB_E_ARG, "vppb_lk_match_u8: winsize %d must be odd and <= 15", P.winsize);
  LkLevels L;
  for (int s = 0; s < P.nlevels; s++) {
    VPPB_REQUIRE(prev[s].base && next[s].base && grad[s].base, VPPB_E_ARG, "vppb_lk_match_u8: level %d has a NULL image", s);
    VPPB_REQUIRE(prev[s].elem_bytes == 1 && next[s].elem_bytes == 1 && grad[s].elem_bytes == 8, VPPB_E_ARG,
                 "vppb_lk_match_u8: level %d element sizes must be 1/1/8", s);
    VPPB_REQUIRE(same_domain(&prev[s], &next[s]) && same_domain(&prev[s], &grad[s]), VPPB_E_ARG, "vppb_lk_match_u8: level %d domains differ", s);
    VPPB_REQUIRE(prev[s].border >= 1 && next[s].border >= 1 && grad[s].border >= 1, VPPB_E_BORDER, "vppb_lk_match_u8: level %d needs border >= 1", s);
    VPPB_REQUIRE(((uintptr_t)grad[s].base % 8) == 0 && (grad[s].pitch % 8) == 0, VPPB_E_ARG, "vppb_lk_match_u8: gradient level %d not 8-byte aligned", s);
    L.prev[s] = view(&prev[s]); L.next[s] = view(&next[s]); L.grad[s] = view(&grad[s]);
  }
  if (n == 0) return VPPB_OK;
  const int npix = P.winsize * P.winsize;
  int ppl = (npix + 31) / 32;
  if (ppl == 5) ppl = 6;
  if (ppl == 7) ppl = 8;
  const int grid = (n + 3) / 4;  // 4 warps (keypoints) per CTA
  static int use_v1 = -1;  // VPPB_LK_V1=1 forces the one-keypoint-per-warp kernel (A/B comparisons)
  if (use_v1 < 0) { const char* e = getenv("VPPB_LK_V1"); use_v1 = (e && atoi(e)) ? 1 : 0; }
  bool done = false;
  bool small = true;  // the v2 kernel addresses samples with 32-bit offsets from pixel (0,0)
  for (int s = 0; s < P.nlevels; s++)
    small = small && ((long long)(prev[s].nrows + 2LL * prev[s].border) * prev[s].pitch < (1LL << 31)) && ((long long)(next[s].nrows + 2LL * next[s].border) * next[s].pitch < (1LL << 31));
  if (!use_v1 && small)
    done = P.grad_is_float ? lk_launch_v2<true>(P.winsize, as_stream(stream), L, P, kps, prediction, n, flow_out, err_out)
                           : lk_launch_v2<false>(P.winsize, as_stream(stream), L, P, kps, prediction, n, flow_out, err_out);
  if (!done) {
    if (P.grad_is_float) lk_launch<true>(ppl, grid, as_stream(stream), L, P, kps, prediction, n, flow_out, err_out);
    else lk_launch<false>(ppl, grid, as_stream(stream), L, P, kps, prediction, n, flow_out, err_out);
  }
  VPPB_LAUNCH_CHECK("vppb_lk_match_u8");
  return VPPB_OK;
}

int vppb_lk_match_oriented_u8(const vppb_img* a, const vppb_img* b, const vppb_img* grad, int32_t grad_is_float, int32_t winsize, float min_ev,
                              int32_t max_iter, float delta, float max_step_norm, const vppb_float2* kps, const vppb_float2* prediction,
                              const vppb_float2* dir1, const vppb_float2* dir2, int32_t n, vppb_float2* flow_out, float* err_out, void* stream) {
  VPPB_REQUIRE(a && b && grad && a->base && b->base && grad->base, VPPB_E_ARG, "vppb_lk_match_oriented_u8: NULL argument");
  VPPB_REQUIRE(n == 0 || (kps && prediction && dir1 && dir2 && flow_out && err_out), VPPB_E_ARG, "vppb_lk_match_oriented_u8: NULL keypoint/output array");
  VPPB_REQUIRE(winsize >= 1 && winsize <= 15 && (winsize & 1), VPPB_E_ARG, "vppb_lk_match_oriented_u8: winsize %d must be odd and <= 15", winsize);
  VPPB_REQUIRE(a->elem_bytes == 1 && b->elem_bytes == 1 && grad->elem_bytes == 8, VPPB_E_ARG, "vppb_lk_match_oriented_u8: element sizes must be 1/1/8");
  VPPB_REQUIRE(same_domain(a, b) && same_domain(a, grad), VPPB_E_ARG, "vppb_lk_match_oriented_u8: domains differ");
  VPPB_REQUIRE(a->border >= 1 && b->border >= 1 && grad->border >= 1, VPPB_E_BORDER, "vppb_lk_match_oriented_u8: border >= 1 needed");
  VPPB_REQUIRE(((uintptr_t)grad->base % 8) == 0 && (grad->pitch % 8) == 0, VPPB_E_ARG, "vppb_lk_match_oriented_u8: gradient not 8-byte aligned");
  if (n == 0) return VPPB_OK;
  LkOriented P;
  P.A = view(a); P.B = view(b); P.Ag = view(grad);
  P.winsize = winsize; P.max_iter = max_iter; P.min_ev = min_ev; P.delta = delta; P.max_step = max_step_norm;
  int ppl = (winsize * winsize + 31) / 32;
  if (ppl == 5) ppl = 6;
  if (ppl == 7) ppl = 8;
  const int grid = (n + 3) / 4;
  if (grad_is_float) lk_oriented_launch<true>(ppl, grid, as_stream(stream), P, kps, prediction, dir1, dir2, n, flow_out, err_out);
  else lk_oriented_launch<false>(ppl, grid, as_stream(stream), P, kps, prediction, dir1, dir2, n, flow_out, err_out);
  VPPB_LAUNCH_CHECK("vppb_lk_match_oriented_u8");
  return VPPB_OK;
}

}  // extern "C"
