/*
 * vpp_oracle_sdof.c — CPU restatement of the semi-dense optical flow used by video_extruder
 * (reference: vpp/algorithms/optical_flow/semi_dense_optical_flow.hpp:17-214,
 *  vpp/algorithms/optical_flow/gradient_descent.hh:10-89), SERIAL semantics: keypoints in index
 * order (first keypoint in a cell claims it), Gauss-Seidel propagation sweeps in the reference's raster
 * orders.  TEST INFRASTRUCTURE ONLY — see vpp_oracle.h.
 */
#include "vpp_oracle.h"

#include <limits.h>
#include <stdlib.h>
#include <string.h>

#define ROW(img, r) ((img)->base + (int64_t)(r) * (img)->pitch)

typedef struct { vo_img img; unsigned char* buf; } owned_img;

static owned_img make_img(int nr, int nc, int elem, int border) {
  owned_img o;
  int pitch; int64_t total, origin;
  vo_layout(nr, nc, elem, border, 16, &pitch, &total, &origin);
  o.buf = (unsigned char*)calloc((size_t)total + 16, 1);
  o.img.base = o.buf + origin; o.img.nrows = nr; o.img.ncols = nc; o.img.pitch = pitch; o.img.border = border; o.img.elem = elem;
  return o;
}

/* semi_dense_optical_flow.hpp:17-42.  The row-wise early exit (err <= th) only short-cuts sums that are
 * already above the threshold, so every comparison `d < match_distance` sees the same outcome. */
static int sad_distance(const vo_img* i1, const vo_img* i2, int ar, int ac, int br, int bc, int ws, int th) {
  int err = 0;
  const unsigned char* row1 = ROW(i1, ar - ws / 2) + (ac - ws / 2);
  const unsigned char* row2 = ROW(i2, br - ws / 2) + (bc - ws / 2);
  for (int r = 0; r < ws && err <= th; r++) {
    int err2 = 0;
    for (int c = 0; c < ws; c++) err2 += abs((int)row1[c] - (int)row2[c]);
    err += err2;
    row1 += i1->pitch;
    row2 += i2->pitch;
  }
  return err;
}

typedef struct { const vo_img *i1, *i2; int ws; } dist_ctx;

/* semi_dense_optical_flow.hpp:102-108 */
static int distance(const dist_ctx* c, int ar, int ac, int br, int bc, int max_distance) {
  if (ar >= 0 && ar < c->i1->nrows && ac >= 0 && ac < c->i1->ncols && br >= 0 && br < c->i2->nrows && bc >= 0 && bc < c->i2->ncols)
    return sad_distance(c->i1, c->i2, ar, ac, br, bc, c->ws, max_distance);
  return INT_MAX;
}

/* gradient_descent.hh:10-89, tables verbatim */
static void gradient_descent_match(const dist_ctx* ctx, int pr, int pc, int predr, int predc, int max_iteration, int* flow_r, int* flow_c,
                                   int* dist_out) {
  static const int c8_it[9][2] = {{6, 3}, {0, 3}, {0, 5}, {2, 5}, {2, 7}, {4, 7}, {4, 1}, {6, 1}, {0, 0}};
  static const int c8[8][2] = {{-1, 1}, {0, 1}, {1, 1}, {-1, 0}, {1, 0}, {-1, -1}, {0, -1}, {1, -1}};
  int mr = predr, mc = predc;
  int match_distance = distance(ctx, pr, pc, predr, predc, INT_MAX);
  unsigned match_i = 8;
  for (int search = 0; search < max_iteration; search++) {
    int i = c8_it[match_i][0];
    const int end = c8_it[match_i][1];
    {
      int nr = predr + c8[i][0], nc = predc + c8[i][1];
      int d = distance(ctx, pr, pc, nr, nc, match_distance);
      if (d < match_distance) { mr = nr; mc = nc; match_i = (unsigned)i; match_distance = d; }
      i = (i + 1) & 7;
    }
    for (; i != end; i = (i + 1) & 7) {
      int nr = predr + c8[i][0], nc = predc + c8[i][1];
      int d = distance(ctx, pr, pc, nr, nc, match_distance);
      if (d < match_distance) { mr = nr; mc = nc; match_i = (unsigned)i; match_distance = d; }
    }
    if (predr == mr && predc == mc) break;
    predr = mr; predc = mc;
  }
  *flow_r = mr - pr; *flow_c = mc - pc; *dist_out = match_distance;
}

#define FLOW(m, r, c) ((int32_t*)(ROW(&(m).img, r) + (int64_t)(c) * 8))
#define MARK(m, r, c) (ROW(&(m).img, r)[c])
#define DIST(m, r, c) (((int32_t*)ROW(&(m).img, r))[c])

/* semi_dense_optical_flow.hpp:46-214.  kps: (row, col) ints; out_pos[i] = matched position, out_dist[i], out_valid[i]
 * (the arguments the reference hands to match_callback; valid = 0 where it is not called). */
void vo_semi_dense_flow(const vo_img* i1, const vo_img* i2, const vo_int2* kps, int n, int winsize, int nscales, int min_scale,
                        int propagation, int patchsize, vo_int2* out_pos, int32_t* out_dist, unsigned char* out_valid) {
  owned_img* p1 = (owned_img*)malloc(sizeof(owned_img) * nscales);
  owned_img* p2 = (owned_img*)malloc(sizeof(owned_img) * nscales);
  owned_img* fm = (owned_img*)malloc(sizeof(owned_img) * nscales);
  owned_img* mk = (owned_img*)malloc(sizeof(owned_img) * nscales);
  owned_img* dm = (owned_img*)malloc(sizeof(owned_img) * nscales);
  /* pyramids (:70-74): images border 2*winsize, cell maps border nscales, level size 1 + n/2 */
  int nr = i1->nrows, nc = i1->ncols, cr = i1->nrows / patchsize, cc = i1->ncols / patchsize;
  for (int s = 0; s < nscales; s++) {
    p1[s] = make_img(nr, nc, 1, 2 * winsize);
    p2[s] = make_img(nr, nc, 1, 2 * winsize);
    fm[s] = make_img(cr, cc, 8, nscales);
    mk[s] = make_img(cr, cc, 1, nscales);
    dm[s] = make_img(cr, cc, 4, nscales);
    nr = (int)(1 + nr / 2.f); nc = (int)(1 + nc / 2.f); cr = (int)(1 + cr / 2.f); cc = (int)(1 + cc / 2.f);
  }
  vo_copy(i1, &p1[0].img, 0); vo_copy(i2, &p2[0].img, 0);
  vo_fill_border_mirror(&p1[0].img); vo_fill_border_mirror(&p2[0].img);
  for (int s = 1; s < nscales; s++) {
    vo_lowpass_sub2(&p1[s - 1].img, &p1[s].img, 0); vo_fill_border_mirror(&p1[s].img);
    vo_lowpass_sub2(&p2[s - 1].img, &p2[s].img, 0); vo_fill_border_mirror(&p2[s].img);
  }
  for (int scale = nscales - 1; scale >= min_scale; scale--) {
    const int scale_div = 1 << scale;
    dist_ctx ctx = {&p1[scale].img, &p2[scale].img, winsize};
    /* mark <- 0 incl. border (:110-111); the images' borders are already mirror-filled (:98-99) */
    for (int r = -nscales; r < mk[scale].img.nrows + nscales; r++) memset(ROW(&mk[scale].img, r) - nscales, 0, (size_t)mk[scale].img.ncols + 2 * nscales);
    /* gradient descent per keypoint, first keypoint of a cell claims it (:114-143) */
    for (int i = 0; i < n; i++) {
      const int pr = kps[i].r / scale_div, pc = kps[i].c / scale_div;
      const int fr = pr / patchsize, fc = pc / patchsize;
      if (!MARK(mk[scale], fr, fc)) {
        int predr = pr, predc = pc;
        const int mr_ = pr / (2 * patchsize), mc_ = pc / (2 * patchsize);
        if (scale < nscales - 1 && MARK(mk[scale + 1], mr_, mc_)) {
          predr = pr + FLOW(fm[scale + 1], mr_, mc_)[0] * 2;
          predc = pc + FLOW(fm[scale + 1], mr_, mc_)[1] * 2;
        }
        int flr, flc, d;
        gradient_descent_match(&ctx, pr, pc, predr, predc, 5, &flr, &flc, &d);
        FLOW(fm[scale], fr, fc)[0] = flr; FLOW(fm[scale], fr, fc)[1] = flc;
        DIST(dm[scale], fr, fc) = d;
        MARK(mk[scale], fr, fc) = 2;
      }
    }
    /* propagation sweeps (:146-201) */
    const int inr = p1[scale].img.nrows, inc = p1[scale].img.ncols;
    for (int Ki = 0; Ki < propagation; Ki++) {
      const int fwd = Ki % 2;
      for (int r = fwd ? 0 : inr - 1; fwd ? r < inr : r >= 0; r += fwd ? patchsize : -patchsize)
        for (int c = fwd ? 0 : inc - 1; fwd ? c < inc : c >= 0; c += fwd ? patchsize : -patchsize) {
          const int fr = r / patchsize, fc = c / patchsize;
          if (!MARK(mk[scale], fr, fc)) continue;
          const int prev0 = FLOW(fm[scale], fr, fc)[0], prev1 = FLOW(fm[scale], fr, fc)[1];
          for (int dr = -1; dr <= 1; dr++)
            for (int dc = -1; dc <= 1; dc++) {
              if (!dr && !dc) continue;
              const int nr_ = fr + dr, nc_ = fc + dc;
              if (nr_ < 0 || nr_ >= fm[scale].img.nrows || nc_ < 0 || nc_ >= fm[scale].img.ncols || !MARK(mk[scale], nr_, nc_)) continue;
              const int n0 = FLOW(fm[scale], nr_, nc_)[0], n1 = FLOW(fm[scale], nr_, nc_)[1];
              const int a0 = FLOW(fm[scale], fr, fc)[0] - n0, a1 = FLOW(fm[scale], fr, fc)[1] - n1;
              const int b0 = prev0 - n0, b1 = prev1 - n1;
              /* Eigen integer norm() = (int)sqrt(squaredNorm): "> 2" <=> squared norm >= 9 */
              if (a0 * a0 + a1 * a1 < 9 || b0 * b0 + b1 * b1 < 9) continue;
              const int d1 = DIST(dm[scale], fr, fc);
              const int d2 = distance(&ctx, r, c, r + n0, c + n1, INT_MAX);
              if (d2 < d1) {
                int flr, flc, d;
                gradient_descent_match(&ctx, r, c, r + n0, c + n1, 5, &flr, &flc, &d);
                if (d < d1) {
                  MARK(mk[scale], fr, fc) = 1;
                  FLOW(fm[scale], fr, fc)[0] = flr; FLOW(fm[scale], fr, fc)[1] = flc;
                  DIST(dm[scale], fr, fc) = d;
                }
              }
            }
        }
    }
  }
  /* results (:205-212) */
  const int div = patchsize * (1 << min_scale);
  for (int i = 0; i < n; i++) {
    const int fr = kps[i].r / div, fc = kps[i].c / div;
    out_valid[i] = 0; out_pos[i].r = 0; out_pos[i].c = 0; out_dist[i] = 0;
    if (fr >= 0 && fr < mk[min_scale].img.nrows && fc >= 0 && fc < mk[min_scale].img.ncols && MARK(mk[min_scale], fr, fc)) {
      out_valid[i] = 1;
      out_pos[i].r = kps[i].r + FLOW(fm[min_scale], fr, fc)[0] * (1 << min_scale);
      out_pos[i].c = kps[i].c + FLOW(fm[min_scale], fr, fc)[1] * (1 << min_scale);
      out_dist[i] = DIST(dm[min_scale], fr, fc);
    }
  }
  for (int s = 0; s < nscales; s++) { free(p1[s].buf); free(p2[s].buf); free(fm[s].buf); free(mk[s].buf); free(dm[s].buf); }
  free(p1); free(p2); free(fm); free(mk); free(dm);
}
