/*
 * vpp_oracle_algos.c — CPU restatement of FAST9 and (pyramidal) Lucas-Kanade.
 * TEST INFRASTRUCTURE ONLY — see vpp_oracle.h.  Citations are reference file:line.
 */
#include "vpp_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ROW(img, r) ((img)->base + (int64_t)(r) * (img)->pitch)

/* ------------------------------------------------------------------------------------------
 * FAST9.  fast.hpp:253-508 evaluates, for every pixel of the domain, a pruning tree over the 16
 * ring flags a0..a15 (bit 4 = brighter than v+th, bit 0 = darker than v-th, both saturating u8,
 * fast.hpp:120-126,319-324) ANDed into `possible`, which starts at the mask byte (fast.hpp:310-317).
 * The tree is exactly "9 circularly contiguous flags" per polarity bit (checked exhaustively over
 * all 2^16 patterns by tests/test_oracle_kats.py against oracle/_ref, which runs the tree itself).
 * Ring slots as loaded by the reference: slots 4 and 12 come from a_row1 (row r-3) at c+3 / c-3
 * (fast.hpp:367-368) instead of row r; ring = 1 selects the true ring used by fast.hpp:79-112.
 */
static const signed char RING[2][16][2] = {
    {{-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}, {-3, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3}, {-3, -3}, {-1, -3}, {-2, -2}, {-3, -1}},
    {{-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}, {0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3}, {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}}};

static int arc9(unsigned m16) {
  for (int s = 0; s < 16; s++) {
    int ok = 1;
    for (int k = 0; k < 9 && ok; k++) ok = (m16 >> ((s + k) & 15)) & 1;
    if (ok) return 1;
  }
  return 0;
}

/* fast.hpp:36-77 (true ring) */
int vo_fast9_score(const vo_img* img, int th, int r, int c) {
  const unsigned char* p = ROW(img, r) + c;
  int v = *p, sum_inf = 0, sum_sup = 0;
  for (int i = 0; i < 16; i++) {
    int a = p[(int64_t)RING[1][i][0] * img->pitch + RING[1][i][1]];
    int diff = v - a;
    if (diff < -th) sum_inf -= diff;
    else if (diff > th) sum_sup += diff;
  }
  return sum_sup > sum_inf ? sum_sup : sum_inf;
}

static int is_corner(const vo_img* img, int th, int ring, int mask_byte, int r, int c) {
  if (mask_byte == 0) return 0;
  const unsigned char* p = ROW(img, r) + c;
  int v = *p;
  int thb = th & 255;
  int hi = v + thb > 255 ? 255 : v + thb; /* u_adds */
  int lo = v - thb < 0 ? 0 : v - thb;     /* u_subs */
  unsigned mb = 0, md = 0;
  for (int i = 0; i < 16; i++) {
    int a = p[(int64_t)RING[ring][i][0] * img->pitch + RING[ring][i][1]];
    mb |= (unsigned)(a > hi) << i;
    md |= (unsigned)(a < lo) << i;
  }
  return ((mask_byte & 0x10) && arc9(mb)) || ((mask_byte & 0x01) && arc9(md));
}

int vo_fast9_u8(const vo_img* img, int th, const vo_img* mask, int mode, int block_size, int ring, vo_int2* kps, int32_t* scores,
                int capacity) {
  const int nr = img->nrows, nc = img->ncols;
  unsigned char* det = (unsigned char*)calloc((size_t)nr * nc, 1);
  /* fast.hpp:287-499 */
#pragma omp parallel for
  for (int r = 0; r < nr; r++)
    for (int c = 0; c < nc; c++) {
      int m = (mask && mask->base) ? ROW(mask, r)[c] : 255;
      det[(size_t)r * nc + c] = (unsigned char)is_corner(img, th, ring, m, r, c);
    }
  unsigned char* keep = det;
  unsigned char* sc = NULL;
  if (mode != 0) {
    /* fast.hpp:685-694: u8 score image, border 1, zero; scores_img(p) = score / 16 */
    sc = (unsigned char*)calloc((size_t)(nr + 2) * (nc + 2), 1);
#define SC(r, c) sc[(size_t)((r) + 1) * (nc + 2) + (c) + 1]
    for (int r = 0; r < nr; r++)
      for (int c = 0; c < nc; c++)
        if (det[(size_t)r * nc + c]) SC(r, c) = (unsigned char)(vo_fast9_score(img, th, r, c) / 16);
    keep = (unsigned char*)calloc((size_t)nr * nc, 1);
    if (mode == 1) {
      /* fast.hpp:896-927: strictly greater than the 8 neighbours */
      for (int r = 0; r < nr; r++)
        for (int c = 0; c < nc; c++)
          if (det[(size_t)r * nc + c]) {
            unsigned a = SC(r, c);
            int is_max = 1;
            for (int dr = -1; dr <= 1; dr++)
              for (int dc = -1; dc <= 1; dc++)
                if (dr || dc) is_max &= a > SC(r + dr, c + dc);
            keep[(size_t)r * nc + c] = (unsigned char)is_max;
          }
    } else {
      /* fast.hpp:763-790: raster arg-max per block_size^2 cell, strict '>', kept if > 0 */
      for (int r = 0; r < nr; r += block_size)
        for (int c = 0; c < nc; c += block_size) {
          unsigned vmax = 0;
          int pr = 0, pc = 0;
          for (int br = 0; br < block_size; br++)
            for (int bc = c; bc < c + block_size; bc++)
              if (r + br < nr && bc < nc) {
                unsigned v = SC(r + br, bc);
                if (v > vmax) { vmax = v; pr = r + br; pc = bc; }
              }
          if (vmax > 0) keep[(size_t)pr * nc + pc] = 1;
        }
    }
  }
  int n = 0;
  if (mode == 2) {
    /* serial order of fast.hpp:763-790: cells in raster order, at most one keypoint per cell */
    for (int r0 = 0; r0 < nr; r0 += block_size)
      for (int c0 = 0; c0 < nc; c0 += block_size)
        for (int r = r0; r < r0 + block_size && r < nr; r++)
          for (int c = c0; c < c0 + block_size && c < nc; c++)
            if (keep[(size_t)r * nc + c]) {
              if (n < capacity) { kps[n].r = r; kps[n].c = c; if (scores) scores[n] = (int)SC(r, c); }
              n++;
            }
  } else {
    for (int r = 0; r < nr; r++)
      for (int c = 0; c < nc; c++)
        if (keep[(size_t)r * nc + c]) {
          if (n < capacity) {
            kps[n].r = r;
            kps[n].c = c;
            /* fast.hpp:670-671 raw score; :698-704 the u8 score image entry for the maxima modes */
            if (scores) scores[n] = mode == 0 ? vo_fast9_score(img, th, r, c) : (int)SC(r, c);
          }
          n++;
        }
  }
#undef SC
  if (keep != det) free(keep);
  free(det);
  free(sc);
  return n <= capacity ? n : -n;
}

/* ------------------------------------------------------------------------------------------
 * imageNd.hpp:280-300 linear_interpolate: x = int(p) (truncation), a = p - x, four taps weighted
 * in float and summed left to right, result cast back to the pixel type (truncation).
 * The 2x2 footprint is clamped into the allocated frame so the oracle never reads outside its
 * buffers (the reference reads whatever is there; only reachable for out-of-range samples). */
static int clamp_tap(int x, int n, int border) {
  int lo = -border, hi = n + border - 2;
  return x < lo ? lo : (x > hi ? hi : x);
}

int vo_interp_u8(const vo_img* im, float p0, float p1) {
  int x0 = (int)p0, x1 = (int)p1;
  float a0 = p0 - (float)x0, a1 = p1 - (float)x1;
  const unsigned char* l1 = ROW(im, clamp_tap(x0, im->nrows, im->border)) + clamp_tap(x1, im->ncols, im->border);
  const unsigned char* l2 = l1 + im->pitch;
  float res = (1 - a0) * (1 - a1) * (float)l1[0] + a0 * (1 - a1) * (float)l2[0] + (1 - a0) * a1 * (float)l1[1] + a0 * a1 * (float)l2[1];
  return (int)(unsigned char)(int)res; /* cvttss2si, then the low byte */
}

static void interp_grad(const vo_img* im, int is_float, float p0, float p1, float* gx, float* gy) {
  int x0 = (int)p0, x1 = (int)p1;
  float a0 = p0 - (float)x0, a1 = p1 - (float)x1;
  const unsigned char* l1 = ROW(im, clamp_tap(x0, im->nrows, im->border)) + (int64_t)clamp_tap(x1, im->ncols, im->border) * 8;
  const unsigned char* l2 = l1 + im->pitch;
  float v[4][2];
  for (int k = 0; k < 2; k++) {
    if (is_float) {
      v[0][k] = ((const float*)l1)[k]; v[1][k] = ((const float*)l2)[k];
      v[2][k] = ((const float*)l1)[2 + k]; v[3][k] = ((const float*)l2)[2 + k];
    } else {
      v[0][k] = (float)((const int32_t*)l1)[k]; v[1][k] = (float)((const int32_t*)l2)[k];
      v[2][k] = (float)((const int32_t*)l1)[2 + k]; v[3][k] = (float)((const int32_t*)l2)[2 + k];
    }
  }
  float r[2];
  for (int k = 0; k < 2; k++) {
    r[k] = (1 - a0) * (1 - a1) * v[0][k] + a0 * (1 - a1) * v[1][k] + (1 - a0) * a1 * v[2][k] + a0 * a1 * v[3][k];
    if (!is_float) r[k] = (float)(int)r[k]; /* cast<vint2>, then read back as float */
  }
  *gx = r[0];
  *gy = r[1];
}

/* lucas_kanade.hpp:12-131 / lk.hh:42-175.  Returns the match (m0, m1, err). */
static void lk_match(float p0, float p1, float tr0, float tr1, const vo_img* A, const vo_img* B, const vo_img* Ag, const vo_lk_params* P,
                     float* m0, float* m1, float* merr) {
  const int ws = P->winsize, hws = ws / 2, npix = ws * ws;
  float gs0[225], gs1[225];
  int as[225];
  unsigned char valid[225];
  /* gradient matrix (lucas_kanade.hpp:24-43) */
  float G00 = 0, G01 = 0, G11 = 0;
  int cpt = 0, i = 0;
  for (int r = -hws; r <= hws; r++)
    for (int c = -hws; c <= hws; c++, i++) {
      float n0 = p0 + (float)r, n1 = p1 + (float)c;
      int i0 = (int)n0, i1 = (int)n1;
      valid[i] = (unsigned char)(i0 >= 0 && i0 < A->nrows && i1 >= 0 && i1 < A->ncols);
      gs0[i] = 0; gs1[i] = 0; as[i] = 0; /* reference: uninitialised when out of domain */
      if (valid[i]) {
        float gx, gy;
        interp_grad(Ag, P->grad_is_float, n0, n1, &gx, &gy);
        G00 += gx * gx; G01 += gx * gy; G11 += gy * gy;
        cpt++;
        gs0[i] = gx; gs1[i] = gy;                 /* lucas_kanade.hpp:78 */
        as[i] = vo_interp_u8(A, n0, n1);          /* lucas_kanade.hpp:79 */
      }
    }
  /* minimum |eigenvalue| of G / cpt (lucas_kanade.hpp:45-52), symmetric 2x2 closed form */
  {
    float cf = (float)cpt;
    float a = G00 / cf, b = G01 / cf, d = G11 / cf;
    float half = (a + d) * 0.5f, diff = (a - d) * 0.5f;
    float root = sqrtf(diff * diff + b * b);
    float e1 = fabsf(half + root), e2 = fabsf(half - root);
    float min_ev = 99999.f;
    if (e1 < min_ev) min_ev = e1;
    if (e2 < min_ev) min_ev = e2;
    if (min_ev < P->min_ev) { *m0 = -1.f; *m1 = -1.f; *merr = FLT_MAX; return; }
  }
  /* G^-1, Eigen's 2x2 closed form (lucas_kanade.hpp:54) */
  float det = G00 * G11 - G01 * G01;
  float invdet = 1.f / det;
  float I00 = G11 * invdet, I01 = -G01 * invdet, I11 = G00 * invdet;

  float v0 = p0 + tr0, v1 = p1 + tr1;
  float nk0 = 1.f, nk1 = 1.f;
  /* gradient descent (lucas_kanade.hpp:87-113) */
  for (int k = 0; k <= P->max_iter && sqrtf(nk0 * nk0 + nk1 * nk1) >= P->delta; k++) {
    float bk0 = 0, bk1 = 0;
    i = 0;
    for (int r = -hws; r <= hws; r++)
      for (int c = -hws; c <= hws; c++, i++)
        if (valid[i]) {
          float dt = (float)as[i] - (float)vo_interp_u8(B, v0 + (float)r, v1 + (float)c);
          bk0 += gs0[i] * dt;
          bk1 += gs1[i] * dt;
        }
    nk0 = I00 * bk0 + I01 * bk1;
    nk1 = I01 * bk0 + I11 * bk1;
    v0 += nk0;
    v1 += nk1;
    int iv0 = (int)v0, iv1 = (int)v1;
    if (!isfinite(v0) || !isfinite(v1) || iv0 < 0 || iv0 >= B->nrows || iv1 < 0 || iv1 >= B->ncols) {
      *m0 = 0.f; *m1 = 0.f; *merr = FLT_MAX;
      return;
    }
  }
  /* error (lucas_kanade.hpp:116-128; lk.hh:151-173) */
  float err = 0;
  i = 0;
  for (int r = -hws; r <= hws; r++)
    for (int c = -hws; c <= hws; c++, i++) {
      err += fabsf((float)(as[i] - vo_interp_u8(B, v0 + (float)r, v1 + (float)c)));
      cpt++;
    }
  if (P->err_mode == 0) {
    *merr = err / (float)cpt;
  } else {
    float avg = 0, stddev = 0;
    for (i = 0; i < npix; i++) avg += (float)as[i];
    avg /= (float)npix;
    for (i = 0; i < npix; i++) stddev += fabsf(avg - (float)as[i]);
    stddev /= (float)npix;
    *merr = err / ((float)cpt * stddev);
  }
  *m0 = v0 - p0;
  *m1 = v1 - p1;
}

/* lucas_kanade.hpp:159-181 (serial over keypoints) / pyrlk_match.hh:24-41 (omp parallel for) */
void vo_lk_match_u8(const vo_img* prev, const vo_img* next, const vo_img* grad, const vo_lk_params* P, const vo_float2* kps,
                    const vo_float2* prediction, int n, vo_float2* flow_out, float* err_out) {
#pragma omp parallel for schedule(dynamic, 16)
  for (int k = 0; k < n; k++) {
    float tr0 = 0.f, tr1 = 0.f, dist = 0.f;
    if (prediction) {
      tr0 = prediction[k].r / P->pred_div;
      tr1 = prediction[k].c / P->pred_div;
    }
    for (int S = P->nlevels - 1; S >= P->min_scale; S--) {
      tr0 *= P->factor;
      tr1 *= P->factor;
      float scale = (float)(1 << S);
      float m0, m1, merr;
      lk_match(kps[k].r / scale, kps[k].c / scale, tr0, tr1, &prev[S], &next[S], &grad[S], P, &m0, &m1, &merr);
      if (!P->gate_on_max_err || merr < P->max_err) { tr0 = m0; tr1 = m1; }
      dist = merr;
    }
    flow_out[k].r = tr0;
    flow_out[k].c = tr1;
    err_out[k] = dist;
  }
}

/* kitti::flow_error_stats (evaluation/utils/kitti.hh:75-134): flow / ref are (nrows x ncols) vfloat3 images (12-byte pixels, channel 2 > 0 =
 * defined).  out[0..5] = n1, n3, n5, n10 (percent), avg end-point error, density (percent); returns the number of compared vectors.
 * errors_map (u8, may be NULL) = min(err * 20, 255). */
int vo_flow_error_stats(const vo_img* flow, const vo_img* ref, float* out, const vo_img* errors_map) {
  int n = 0, cpt = 0, n1 = 0, n3 = 0, n5 = 0, n10 = 0;
  float error_sum = 0.f;
  for (int r = 0; r < flow->nrows; r++)
    for (int c = 0; c < flow->ncols; c++) {
      const float* f = (const float*)((const unsigned char*)flow->base + (long long)r * flow->pitch + (long long)c * 12);
      const float* g = (const float*)((const unsigned char*)ref->base + (long long)r * ref->pitch + (long long)c * 12);
      if (errors_map) ((unsigned char*)errors_map->base)[(long long)r * errors_map->pitch + c] = 0;
      if (f[2] > 0.f) cpt++;
      if (f[2] > 0.f && g[2] > 0.f) {
        n++;
        const float d0 = f[0] - g[0], d1 = f[1] - g[1];
        const float err = sqrtf(d0 * d0 + d1 * d1);
        error_sum += err;
        if (err > 1.f) n1++;
        if (err > 3.f) n3++;
        if (err > 5.f) n5++;
        if (err > 10.f) n10++;
        if (errors_map) {
          const float m = err * 20.f < 255.f ? err * 20.f : 255.f;
          ((unsigned char*)errors_map->base)[(long long)r * errors_map->pitch + c] = (unsigned char)m;
        }
      }
    }
  out[0] = n1 ? 100 * (float)n1 / n : 0;
  out[1] = n3 ? 100 * (float)n3 / n : 0;
  out[2] = n5 ? 100 * (float)n5 / n : 0;
  out[3] = n10 ? 100 * (float)n10 / n : 0;
  out[4] = error_sum / (n ? n : 1);
  out[5] = 100.f * (float)cpt / (flow->nrows * flow->ncols);
  return n;
}
