/*
 * vpp_oracle.c — CPU restatement of the Video++ dense-pixel hot path (containers, pixel_wise
 * named kernels, border fills, 5x5 box stencil, Scharr, pyramid low-pass).
 * TEST INFRASTRUCTURE ONLY — see vpp_oracle.h.  Citations are reference file:line.
 *
 * `#pragma omp parallel for` sits exactly where the reference has it, so the timing build
 * (-fopenmp) exercises the reference's own parallel structure; the parity build ignores them.
 */
#include "vpp_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int vo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void vo_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

#define ROW(img, r) ((img)->base + (int64_t)(r) * (img)->pitch)

/* imageNd.hpp:151-196: border bytes and pitch rounded up to the alignment; begin_ =
 * data_ + border_padding + offset(border, border). */
int vo_layout(int nrows, int ncols, int elem, int border, int align, int* pitch, int64_t* total, int64_t* origin) {
  if (nrows <= 0 || ncols <= 0 || elem <= 0 || border < 0 || align <= 0) return -2;
  int64_t border_size = (int64_t)border * elem, border_padding = 0;
  if (border_size % align) {
    border_padding = align - (border_size % align);
    border_size += border_padding;
  }
  int64_t p = (int64_t)ncols * elem + border_size * 2;
  if (p % align) p += align - (p % align);
  if (pitch) *pitch = (int)p;
  if (total) *total = ((int64_t)nrows + 2 * border) * p;
  if (origin) *origin = border_padding + (int64_t)border * p + (int64_t)border * elem;
  return 0;
}

/* pixel_wise(A,B,C) | a = b + c  (pixel_wise.hpp:85-93 row-parallel, :69-73 process_row;
 * kernel benchmarks/image_add.cc:51-57).  Domain = first argument's. */
void vo_pw_add_i32(const vo_img* a, const vo_img* b, const vo_img* c) {
  const int nr = a->nrows, nc = a->ncols;
#pragma omp parallel for
  for (int r = 0; r < nr; r++) {
    int32_t* pa = (int32_t*)ROW(a, r);
    const int32_t* pb = (const int32_t*)ROW(b, r);
    const int32_t* pc = (const int32_t*)ROW(c, r);
    for (int col = 0; col < nc; col++) pa[col] = (int32_t)((uint32_t)pb[col] + (uint32_t)pc[col]);
  }
}

/* fill.hh:12-15 (domain) / :24-28 (domain_with_border) */
void vo_fill(const vo_img* img, const void* value, int with_border) {
  const int b = with_border ? img->border : 0, e = img->elem;
#pragma omp parallel for
  for (int r = -b; r < img->nrows + b; r++) {
    unsigned char* p = ROW(img, r);
    for (int c = -b; c < img->ncols + b; c++) memcpy(p + (int64_t)c * e, value, e);
  }
}

/* copy.hh:10-13 (domain of src) / :22-27 (src.domain_with_border) */
void vo_copy(const vo_img* src, const vo_img* dst, int with_border) {
  const int b = with_border ? src->border : 0, e = src->elem;
#pragma omp parallel for
  for (int r = -b; r < src->nrows + b; r++)
    memcpy(ROW(dst, r) - (int64_t)b * e, ROW(src, r) - (int64_t)b * e, (size_t)(src->ncols + 2 * b) * e);
}

/* colorspace_conversions.hh:10-47: o = (i[0] + i[1] + i[2]) / 3 (int arithmetic, truncating; a 4th channel is ignored),
 * for every pixel of in.domain_with_border(); `out` is built with the same border.  Here the frame converted is
 * out's domain + min(in, out) border. */
void vo_rgb_to_graylevel(const vo_img* in, const vo_img* out) {
  const int b = in->border < out->border ? in->border : out->border, e = in->elem;
#pragma omp parallel for
  for (int r = -b; r < out->nrows + b; r++) {
    const unsigned char* i = ROW(in, r);
    unsigned char* o = ROW(out, r);
    for (int c = -b; c < out->ncols + b; c++) o[c] = (unsigned char)(((int)i[(int64_t)c * e] + (int)i[(int64_t)c * e + 1] + (int)i[(int64_t)c * e + 2]) / 3);
  }
}

static inline unsigned char* px(const vo_img* img, int r, int c) { return ROW(img, r) + (int64_t)c * img->elem; }

/* fill.hh:32-45: four strips top / bottom / left / right */
void vo_fill_border_value(const vo_img* img, const void* value) {
  const int b = img->border, nr = img->nrows, nc = img->ncols, e = img->elem;
  for (int r = -b; r < nr + b; r++)
    for (int c = -b; c < nc + b; c++)
      if (r < 0 || r >= nr || c < 0 || c >= nc) memcpy(px(img, r, c), value, e);
}

/* fill.hh:48-83: img(-1-k, .) = img(k, .), img(nr+k, .) = img(nr-1-k, .), same for columns,
 * corners from the diagonally mirrored pixel; every source pixel lies in the domain. */
void vo_fill_border_mirror(const vo_img* img) {
  const int b = img->border, nr = img->nrows, nc = img->ncols, e = img->elem;
  for (int r = -b; r < nr + b; r++)
    for (int c = -b; c < nc + b; c++) {
      if (r >= 0 && r < nr && c >= 0 && c < nc) continue;
      int sr = r < 0 ? -r - 1 : (r >= nr ? 2 * nr - r - 1 : r);
      int sc = c < 0 ? -c - 1 : (c >= nc ? 2 * nc - c - 1 : c);
      memcpy(px(img, r, c), px(img, sr, sc), e);
    }
}

/* fill.hh:86-121: clamp to the nearest domain pixel */
void vo_fill_border_closest(const vo_img* img) {
  const int b = img->border, nr = img->nrows, nc = img->ncols, e = img->elem;
  for (int r = -b; r < nr + b; r++)
    for (int c = -b; c < nc + b; c++) {
      if (r >= 0 && r < nr && c >= 0 && c < nc) continue;
      int sr = r < 0 ? 0 : (r >= nr ? nr - 1 : r);
      int sc = c < 0 ? 0 : (c >= nc ? nc - 1 : c);
      memcpy(px(img, r, c), px(img, sr, sc), e);
    }
}

/* sum.hh:12-19: serial, accumulator plus_promotion<V> (= int for char / uchar / int) */
int64_t vo_sum_i32(const vo_img* img, int is_signed) {
  uint32_t acc = 0; /* int arithmetic modulo 2^32 */
  for (int r = 0; r < img->nrows; r++) {
    const unsigned char* p = ROW(img, r);
    for (int c = 0; c < img->ncols; c++) {
      if (img->elem == 4) acc += (uint32_t)((const int32_t*)p)[c];
      else if (is_signed) acc += (uint32_t)(int32_t)((const signed char*)p)[c];
      else acc += p[c];
    }
  }
  return (int64_t)(int32_t)acc;
}

/* 5x5 box on u8 channels: vint3 accumulator, `/ 25` integer division, cast back to uchar
 * (benchmarks/box_5x5_filter2.cc:71-81 kernel body on the vuchar3 image of BASELINE config 2;
 * accumulate-then-divide form of examples/box_filter.cc:23-32). */
void vo_box5x5_u8(const vo_img* in, const vo_img* out, int channels) {
  const int nr = out->nrows, nc = out->ncols, ch = channels;
#pragma omp parallel for
  for (int r = 0; r < nr; r++) {
    unsigned char* o = ROW(out, r);
    const unsigned char* rows[5];
    for (int i = -2; i <= 2; i++) rows[i + 2] = ROW(in, r + i);
    for (int x = 0; x < nc * ch; x++) {
      int sum = 0;
      for (int d = 0; d < 5; d++)
        for (int e = -2; e <= 2; e++) sum += rows[d][x + e * ch];
      o[x] = (unsigned char)(sum / 25);
    }
  }
}

/* benchmarks/box_5x5_filter2.cc:71-81 on image2d<int> (int sum wraps on two's complement) */
void vo_box5x5_i32(const vo_img* in, const vo_img* out) {
  const int nr = out->nrows, nc = out->ncols;
#pragma omp parallel for
  for (int r = 0; r < nr; r++) {
    int32_t* o = (int32_t*)ROW(out, r);
    const int32_t* rows[5];
    for (int i = -2; i <= 2; i++) rows[i + 2] = (const int32_t*)ROW(in, r + i);
    for (int c = 0; c < nc; c++) {
      uint32_t sum = 0;
      for (int d = 0; d < 5; d++)
        for (int e = -2; e <= 2; e++) sum += (uint32_t)rows[d][c + e];
      o[c] = (int32_t)sum / 25;
    }
  }
}

/* scharr.hh:46-87.  component 0 = d/d(row), component 1 = d/d(col); arithmetic in the output
 * component type Vt (int or float), then `/ 32.f` in float, then converted to Vt (truncation
 * toward zero for int). */
void vo_scharr_u8(const vo_img* in, const vo_img* out, int as_float) {
  const int nr = out->nrows, nc = out->ncols;
#pragma omp parallel for
  for (int r = 0; r < nr; r++) {
    const unsigned char* row1 = ROW(in, r - 1);
    const unsigned char* row2 = ROW(in, r);
    const unsigned char* row3 = ROW(in, r + 1);
    if (as_float) {
      float* o = (float*)ROW(out, r);
      for (int c = 0; c < nc; c++) {
        float a = (3 * (float)row3[c - 1] + 10 * (float)row3[c] + 3 * (float)row3[c + 1] - 3 * (float)row1[c - 1] -
                   10 * (float)row1[c] - 3 * (float)row1[c + 1]) / 32.f;
        float b = (3 * (float)row1[c + 1] + 10 * (float)row2[c + 1] + 3 * (float)row3[c + 1] - 3 * (float)row1[c - 1] -
                   10 * (float)row2[c - 1] - 3 * (float)row3[c - 1]) / 32.f;
        o[2 * c] = a;
        o[2 * c + 1] = b;
      }
    } else {
      int32_t* o = (int32_t*)ROW(out, r);
      for (int c = 0; c < nc; c++) {
        int a = 3 * (int)row3[c - 1] + 10 * (int)row3[c] + 3 * (int)row3[c + 1] - 3 * (int)row1[c - 1] - 10 * (int)row1[c] -
                3 * (int)row1[c + 1];
        int b = 3 * (int)row1[c + 1] + 10 * (int)row2[c + 1] + 3 * (int)row3[c + 1] - 3 * (int)row1[c - 1] -
                10 * (int)row2[c - 1] - 3 * (int)row3[c - 1];
        o[2 * c] = (int32_t)((float)a / 32.f);
        o[2 * c + 1] = (int32_t)((float)b / 32.f);
      }
    }
  }
}

/* ---- pyramid.hh:12-59 antialiasing_lowpass_filter ------------------------------------------
 * H pass over the domain rows reading the (filled) column border of `in`; temp gets border 2,
 * mirror-filled (:36); V pass reads temp rows r-2..r+2.  kind 0: u8 (S = int, /16 integer),
 * 1: vint2 (integer /16 per component, truncating toward 0), 2: vfloat2 (float, left-to-right). */
static void lowpass_full(const vo_img* in, vo_img* tmp2 /* nr x nc result, border 2, mirror-filled on return */, int kind) {
  const int nr = in->nrows, nc = in->ncols;
  const int comps = kind == 0 ? 1 : 2;
  const int e = in->elem;
  int pitch; int64_t total, origin;
  vo_layout(nr, nc, e, 2, 16, &pitch, &total, &origin);
  unsigned char* tbuf = (unsigned char*)malloc((size_t)total + 16);
  vo_img tmp = {tbuf + origin, nr, nc, pitch, 2, e};
#pragma omp parallel for
  for (int r = 0; r < nr; r++) {
    for (int c = 0; c < nc; c++) {
      for (int k = 0; k < comps; k++) {
        if (kind == 0) {
          const unsigned char* i = ROW(in, r);
          int s = (1 * (int)i[c - 2] + 4 * (int)i[c - 1] + 6 * (int)i[c] + 4 * (int)i[c + 1] + 1 * (int)i[c + 2]) / 16;
          ROW(&tmp, r)[c] = (unsigned char)s;
        } else if (kind == 1) {
          const int32_t* i = (const int32_t*)ROW(in, r);
          int s = (1 * i[2 * (c - 2) + k] + 4 * i[2 * (c - 1) + k] + 6 * i[2 * c + k] + 4 * i[2 * (c + 1) + k] +
                   1 * i[2 * (c + 2) + k]) / 16;
          ((int32_t*)ROW(&tmp, r))[2 * c + k] = s;
        } else {
          const float* i = (const float*)ROW(in, r);
          float s = (1 * i[2 * (c - 2) + k] + 4 * i[2 * (c - 1) + k] + 6 * i[2 * c + k] + 4 * i[2 * (c + 1) + k] +
                     1 * i[2 * (c + 2) + k]) / 16;
          ((float*)ROW(&tmp, r))[2 * c + k] = s;
        }
      }
    }
  }
  vo_fill_border_mirror(&tmp);
#pragma omp parallel for
  for (int r = 0; r < nr; r++) {
    for (int c = 0; c < nc; c++) {
      for (int k = 0; k < comps; k++) {
        if (kind == 0) {
          int s = (1 * (int)ROW(&tmp, r - 2)[c] + 4 * (int)ROW(&tmp, r - 1)[c] + 6 * (int)ROW(&tmp, r)[c] +
                   4 * (int)ROW(&tmp, r + 1)[c] + 1 * (int)ROW(&tmp, r + 2)[c]) / 16;
          ROW(tmp2, r)[c] = (unsigned char)s;
        } else if (kind == 1) {
          int j = 2 * c + k;
          int s = (1 * ((int32_t*)ROW(&tmp, r - 2))[j] + 4 * ((int32_t*)ROW(&tmp, r - 1))[j] + 6 * ((int32_t*)ROW(&tmp, r))[j] +
                   4 * ((int32_t*)ROW(&tmp, r + 1))[j] + 1 * ((int32_t*)ROW(&tmp, r + 2))[j]) / 16;
          ((int32_t*)ROW(tmp2, r))[j] = s;
        } else {
          int j = 2 * c + k;
          float s = (1 * ((float*)ROW(&tmp, r - 2))[j] + 4 * ((float*)ROW(&tmp, r - 1))[j] + 6 * ((float*)ROW(&tmp, r))[j] +
                     4 * ((float*)ROW(&tmp, r + 1))[j] + 1 * ((float*)ROW(&tmp, r + 2))[j]) / 16;
          ((float*)ROW(tmp2, r))[j] = s;
        }
      }
    }
  }
  free(tbuf);
}

void vo_lowpass(const vo_img* in, const vo_img* out, int kind) {
  vo_img o = *out;
  lowpass_full(in, &o, kind);
}

/* pyramid.hh:179-181 (temp with border 3) + :62-81 subsample2: out(r,c) = temp(2r, 2c).
 * DEVIATION: the reference never fills the temp border, so for an even parent size the last
 * output row/col reads uninitialised memory; here the temp border is mirror-filled. */
void vo_lowpass_sub2(const vo_img* in, const vo_img* out, int kind) {
  const int nr = in->nrows, nc = in->ncols, e = in->elem;
  int pitch; int64_t total, origin;
  vo_layout(nr, nc, e, 3, 16, &pitch, &total, &origin);
  unsigned char* buf = (unsigned char*)malloc((size_t)total + 16);
  vo_img tmp = {buf + origin, nr, nc, pitch, 3, e};
  lowpass_full(in, &tmp, kind);
  vo_fill_border_mirror(&tmp);
#pragma omp parallel for
  for (int r = 0; r < out->nrows; r++)
    for (int c = 0; c < out->ncols; c++) memcpy(px(out, r, c), px(&tmp, 2 * r, 2 * c), e);
  free(buf);
}
