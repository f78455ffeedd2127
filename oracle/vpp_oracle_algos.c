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

/* ==========================================================================================
 * the remaining stencils on the same path.
 * ========================================================================================== */

/* lbp_transform (vpp/algorithms/lbp/lbp_transform.hh:7-38), V = U = unsigned char: bit k of out(r, c) = neighbour k > centre, neighbours
 * in the order (-1,-1) (-1,0) (-1,1) (0,-1) (0,1) (1,-1) (1,0) (1,1).  `in` needs a border >= 1 (read as it is: the caller fills it). */
void vo_lbp_u8(const vo_img* in, const vo_img* out) {
#pragma omp parallel for
  for (int r = 0; r < in->nrows; r++) {
    const unsigned char *r0 = ROW(in, r - 1), *r1 = ROW(in, r), *r2 = ROW(in, r + 1);
    unsigned char* o = ROW(out, r);
    for (int i = 0; i < in->ncols; i++)
      o[i] = (unsigned char)(((r0[i - 1] > r1[i]) << 0) + ((r0[i] > r1[i]) << 1) + ((r0[i + 1] > r1[i]) << 2) + ((r1[i - 1] > r1[i]) << 3) +
                             ((r1[i + 1] > r1[i]) << 4) + ((r2[i - 1] > r1[i]) << 5) + ((r2[i] > r1[i]) << 6) + ((r2[i + 1] > r1[i]) << 7));
  }
}

/* local_maxima_filter (fast.hpp:555-575): IN PLACE, a pixel that is not strictly greater than its 8 neighbours becomes 0 - and the
 * neighbours above / to the left have already been filtered when it is looked at.  The reference runs it through pixel_wise (rows in
 * parallel under OpenMP: racy); this is the serial order (row by row, left to right), which is what a build without OpenMP computes.
 * elem 1 = unsigned char, elem 4 = int32 / unsigned int values below 2^31.  Needs a border >= 1 (read as it is, never written). */
void vo_local_maxima_filter(const vo_img* img) {
  for (int r = 0; r < img->nrows; r++)
    for (int c = 0; c < img->ncols; c++) {
      int is_max = 1;
      if (img->elem == 1) {
        const unsigned char *r0 = ROW(img, r - 1) + c, *r1 = ROW(img, r) + c, *r2 = ROW(img, r + 1) + c;
        const unsigned char a = *r1;
        is_max &= a > r0[-1]; is_max &= a > r0[0]; is_max &= a > r0[1]; is_max &= a > r1[-1]; is_max &= a > r1[1];
        is_max &= a > r2[-1]; is_max &= a > r2[0]; is_max &= a > r2[1];
        if (!is_max) ROW(img, r)[c] = 0;
      } else {
        const int32_t *r0 = (const int32_t*)ROW(img, r - 1) + c, *r1 = (const int32_t*)ROW(img, r) + c, *r2 = (const int32_t*)ROW(img, r + 1) + c;
        const int32_t a = *r1;
        is_max &= a > r0[-1]; is_max &= a > r0[0]; is_max &= a > r0[1]; is_max &= a > r1[-1]; is_max &= a > r1[1];
        is_max &= a > r2[-1]; is_max &= a > r2[0]; is_max &= a > r2[1];
        if (!is_max) ((int32_t*)ROW(img, r))[c] = 0;
      }
    }
}

/* fast_detector9_blockwise_rank (fast.hpp:801-886) on fast_detector9_maxima2 (fast.hpp:710-740): raw fast9_score at every detected corner
 * (0 elsewhere, border 1 of zeros); per block_size x block_size block the strict 3x3 maxima of that image enter a table of
 * max_points slots by the reference's rule - a candidate REPLACES the first slot whose score is smaller (it is not an insertion: the
 * slot's previous point is lost) -, the table is sorted by decreasing score and every non-empty slot k is reported as (row, col, k).
 * The reference never instantiates this template (fast.hpp:949-952 is commented out, and its mask parameter types do not match), so this
 * restatement is NOT pinned to a build of the reference; std::sort on <= 16 elements is libstdc++'s insertion sort, i.e. stable.
 * Blocks in raster order (the serial order of the reference's loops).  kps3: (row, col, k) triples; scores (may be NULL): raw scores.
 * Returns the number of records, or -(needed) if capacity is too small. */
int vo_fast9_blockwise_rank(const vo_img* img, int th, const vo_img* mask, int block_size, int max_points, int ring, int32_t* kps3, int32_t* scores,
                            int capacity) {
  const int nr = img->nrows, nc = img->ncols;
  if (max_points < 1 || max_points > 16 || block_size < 1) return 0;
  int32_t* S = (int32_t*)calloc((size_t)(nr + 2) * (nc + 2), sizeof(int32_t));
#define S_(r, c) S[(size_t)((r) + 1) * (nc + 2) + (c) + 1]
#pragma omp parallel for
  for (int r = 0; r < nr; r++)
    for (int c = 0; c < nc; c++) {
      const int m = (mask && mask->base) ? ROW(mask, r)[c] : 255;
      if (is_corner(img, th, ring, m, r, c)) S_(r, c) = vo_fast9_score(img, th, r, c);
    }
  int n = 0;
  for (int r = 0; r < nr; r += block_size)
    for (int c = 0; c < nc; c += block_size) {
      int pv[16], pr[16], pc[16];
      for (int k = 0; k < max_points; k++) { pv[k] = 0; pr[k] = 0; pc[k] = 0; }
      for (int br = 0; br < block_size; br++)
        for (int bc = c; bc < c + block_size; bc++)
          if (r + br < nr && bc < nc) {
            const int v = S_(r + br, bc);
            if (v > 0) {
              int is_max = 1;
              for (int dr = -1; dr <= 1; dr++)
                for (int dc = -1; dc <= 1; dc++)
                  if (dr || dc) is_max &= v > S_(r + br + dr, bc + dc);
              if (is_max)
                for (int k = 0; k < max_points; k++)
                  if (pv[k] < v) { pv[k] = v; pr[k] = br; pc[k] = bc; break; }
            }
          }
      /* stable insertion sort, decreasing score */
      for (int i = 1; i < max_points; i++) {
        const int v = pv[i], a = pr[i], b = pc[i];
        int j = i - 1;
        while (j >= 0 && pv[j] < v) { pv[j + 1] = pv[j]; pr[j + 1] = pr[j]; pc[j + 1] = pc[j]; j--; }
        pv[j + 1] = v; pr[j + 1] = a; pc[j + 1] = b;
      }
      for (int k = 0; k < max_points; k++)
        if (pv[k] > 0) {
          if (n < capacity) {
            kps3[3 * n] = r + pr[k]; kps3[3 * n + 1] = pc[k]; kps3[3 * n + 2] = k;
            if (scores) scores[n] = pv[k];
          }
          n++;
        }
    }
#undef S_
  free(S);
  return n <= capacity ? n : -n;
}

/* oriented_lk_match_point_square_win<WS>::operator() (lk.hh:180-317), one level, per keypoint: the gradient matrix over the axis-aligned
 * window, the template sampled on a window rotated to `dir1` (columns along dir1, rows along its normal), the search window rotated
 * to `dir2`; min |eigenvalue| of G itself (not G / cpt, lk.hh:219), steps clamped to max_step_norm (lk.hh:276-280), k < max_iter
 * (lk.hh:258), the search confined to B's domain shrunk by 3 (lk.hh:251,283), error = SAD / (cpt * MAD) (lk.hh:288-314).
 * Same deviations as vo_lk_match_u8: as[] / gs[] of samples outside the domain are zero, bilinear footprints are clamped into the frame. */
void vo_lk_match_oriented_u8(const vo_img* A, const vo_img* B, const vo_img* Ag, int grad_is_float, int winsize, float min_ev_th, int max_iter,
                             float delta, float max_step_norm, const vo_float2* kps, const vo_float2* prediction, const vo_float2* dir1,
                             const vo_float2* dir2, int n, vo_float2* flow_out, float* err_out) {
#pragma omp parallel for schedule(dynamic, 16)
  for (int q = 0; q < n; q++) {
    const int ws = winsize, hws = ws / 2, npix = ws * ws;
    const float p0 = kps[q].r, p1 = kps[q].c;
    float gs0[225], gs1[225];
    int as[225];
    float G00 = 0, G01 = 0, G11 = 0;
    int cpt = 0;
    for (int r = -hws; r <= hws; r++)
      for (int c = -hws; c <= hws; c++) {
        const float n0 = p0 + (float)r * 1.f, n1 = p1 + (float)c * 1.f;
        const int i0 = (int)n0, i1 = (int)n1;
        if (i0 >= 0 && i0 < A->nrows && i1 >= 0 && i1 < A->ncols) {
          float gx, gy;
          interp_grad(Ag, grad_is_float, n0, n1, &gx, &gy);
          G00 += gx * gx; G01 += gx * gy; G11 += gy * gy;
          cpt++;
        }
      }
    {
      const float half = (G00 + G11) * 0.5f, diff = (G00 - G11) * 0.5f;
      const float root = sqrtf(diff * diff + G01 * G01);
      const float e1 = fabsf(half + root), e2 = fabsf(half - root);
      float min_ev = 99999.f;
      if (e1 < min_ev) min_ev = e1;
      if (e2 < min_ev) min_ev = e2;
      if (min_ev < min_ev_th) { flow_out[q].r = -1.f; flow_out[q].c = -1.f; err_out[q] = FLT_MAX; continue; }
    }
    const float det = G00 * G11 - G01 * G01, invdet = 1.f / det;
    const float I00 = G11 * invdet, I01 = -G01 * invdet, I11 = G00 * invdet;
    float v0 = p0 + prediction[q].r, v1 = p1 + prediction[q].c;
    float nk0 = 1.f, nk1 = 1.f;
    float mx0 = dir1[q].r, mx1 = dir1[q].c, my0 = -mx1, my1 = mx0;
    int i = 0;
    for (int r = -hws; r <= hws; r++)
      for (int c = -hws; c <= hws; c++, i++) {
        /* p + (r * my + c * mx) * factor, factor = 1: the products are rounded before they are added */
        const float n0 = p0 + ((float)r * my0 + (float)c * mx0) * 1.f, n1 = p1 + ((float)r * my1 + (float)c * mx1) * 1.f;
        const int i0 = (int)n0, i1 = (int)n1;
        gs0[i] = 0; gs1[i] = 0; as[i] = 0;
        if (i0 >= 0 && i0 < Ag->nrows && i1 >= 0 && i1 < Ag->ncols) {
          interp_grad(Ag, grad_is_float, n0, n1, &gs0[i], &gs1[i]);
          as[i] = vo_interp_u8(A, n0, n1);
        }
      }
    mx0 = dir2[q].r; mx1 = dir2[q].c; my0 = -mx1; my1 = mx0;
    int failed = 0;
    for (int k = 0; k < max_iter && sqrtf(nk0 * nk0 + nk1 * nk1) >= delta; k++) {
      float bk0 = 0, bk1 = 0;
      i = 0;
      for (int r = -hws; r <= hws; r++)
        for (int c = -hws; c <= hws; c++, i++) {
          const float n0 = v0 + ((float)r * my0 + (float)c * mx0) * 1.f, n1 = v1 + ((float)r * my1 + (float)c * mx1) * 1.f;
          const float dt = (float)as[i] - (float)vo_interp_u8(B, n0, n1);
          bk0 += gs0[i] * dt;
          bk1 += gs1[i] * dt;
        }
      nk0 = I00 * bk0 + I01 * bk1;
      nk1 = I01 * bk0 + I11 * bk1;
      const float nn = sqrtf(nk0 * nk0 + nk1 * nk1);
      if (nn > max_step_norm) { nk0 /= nn; nk1 /= nn; nk0 *= max_step_norm; nk1 *= max_step_norm; }
      v0 += nk0;
      v1 += nk1;
      const int iv0 = (int)v0, iv1 = (int)v1;
      if (!isfinite(v0) || !isfinite(v1) || iv0 < 3 || iv0 > B->nrows - 1 - 3 || iv1 < 3 || iv1 > B->ncols - 1 - 3) { failed = 1; break; }
    }
    if (failed) { flow_out[q].r = 0.f; flow_out[q].c = 0.f; err_out[q] = FLT_MAX; continue; }
    float avg = 0, stddev = 0;
    for (i = 0; i < npix; i++) avg += (float)as[i];
    avg /= (float)npix;
    for (i = 0; i < npix; i++) stddev += fabsf(avg - (float)as[i]);
    stddev /= (float)npix;
    float err = 0;
    i = 0;
    for (int r = -hws; r <= hws; r++)
      for (int c = -hws; c <= hws; c++, i++) {
        const float n0 = v0 + ((float)r * my0 + (float)c * mx0) * 1.f, n1 = v1 + ((float)r * my1 + (float)c * mx1) * 1.f;
        err += fabsf((float)(as[i] - vo_interp_u8(B, n0, n1)));
        cpt++;
      }
    flow_out[q].r = v0 - p0;
    flow_out[q].c = v1 - p1;
    err_out[q] = err / ((float)cpt * stddev);
  }
}
