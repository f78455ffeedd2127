// Scharr gradient and the fused low-pass + subsample pyramid step.
// Reference: vpp/algorithms/filters/scharr.hh:46-87, vpp/core/pyramid.hh:12-59 (1-4-6-4-1
// separable low-pass, H pass then mirror then V pass) and :62-81 (subsample2).
// Both are streaming stencils bounded by HBM: scharr moves 1 + 8 bytes per pixel, a pyramid
// step reads N_{l-1} and writes N_l pixels (no temp image: the reference's H temp and its
// mirror fill are reproduced by index mirroring inside the kernel).
// The library is compiled with -fmad=false: float results must match the reference's
// un-contracted evaluation order bit for bit.
#include "common.cuh"

namespace vppb {

// ------------------------------------------------------------------ Scharr
// 4 pixels per thread: rows r-1, r, r+1 are fetched as three aligned 32-bit words each
// (columns c-4 .. c+7), outputs are two 16-byte stores.
template <bool AS_FLOAT>
__global__ void __launch_bounds__(256) k_scharr_u8(Img in, Img out, int groups_per_row) {
  long long total = (long long)out.nrows * groups_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / groups_per_row);
    const int c0 = (int)(i - (long long)r * groups_per_row) * 4;
    int px[3][6];  // columns c0-1 .. c0+4 of rows r-1, r, r+1
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const unsigned char* row = row_ptr<unsigned char>(in, r - 1 + k);
#pragma unroll
      for (int j = 0; j < 6; j++) {
        int c = c0 - 1 + j;
        // columns past ncols are only read when they are still inside the border frame
        px[k][j] = (c < in.ncols + in.border) ? (int)__ldg(row + c) : 0;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int c = c0 + j;
      if (c >= out.ncols) break;
      const int* r1 = &px[0][j];  // r1[0] = (r-1, c-1), r1[1] = (r-1, c), r1[2] = (r-1, c+1)
      const int* r2 = &px[1][j];
      const int* r3 = &px[2][j];
      // scharr.hh:64-83 — integer-valued in both the int and the float instantiation
      const int a = 3 * r3[0] + 10 * r3[1] + 3 * r3[2] - 3 * r1[0] - 10 * r1[1] - 3 * r1[2];
      const int b = 3 * r1[2] + 10 * r2[2] + 3 * r3[2] - 3 * r1[0] - 10 * r2[0] - 3 * r3[0];
      const float fa = __fdiv_rn((float)a, 32.f), fb = __fdiv_rn((float)b, 32.f);
      if (AS_FLOAT) {
        reinterpret_cast<float2*>(row_ptr<unsigned char>(out, r))[c] = make_float2(fa, fb);
      } else {
        reinterpret_cast<int2*>(row_ptr<unsigned char>(out, r))[c] = make_int2((int)fa, (int)fb);  // trunc toward 0
      }
    }
  }
}

// ------------------------------------------------------------------ low-pass + subsample2
template <int KIND> struct LpT;
template <> struct LpT<0> { typedef unsigned char elem; typedef int acc; static constexpr int comps = 1; };
template <> struct LpT<1> { typedef int elem; typedef int acc; static constexpr int comps = 2; };
template <> struct LpT<2> { typedef float elem; typedef float acc; static constexpr int comps = 2; };

__device__ __forceinline__ int lp5(int a, int b, int c, int d, int e) { return (1 * a + 4 * b + 6 * c + 4 * d + 1 * e) / 16; }
__device__ __forceinline__ float lp5(float a, float b, float c, float d, float e) {
  // ((((1*a + 4*b) + 6*c) + 4*d) + 1*e) / 16, no contraction (pyramid.hh:27-32)
  float s = __fadd_rn(__fmul_rn(1.f, a), __fmul_rn(4.f, b));
  s = __fadd_rn(s, __fmul_rn(6.f, c));
  s = __fadd_rn(s, __fmul_rn(4.f, d));
  s = __fadd_rn(s, __fmul_rn(1.f, e));
  return __fdiv_rn(s, 16.f);
}

__device__ __forceinline__ int mirror_idx(int i, int n) { return i < 0 ? -i - 1 : (i >= n ? 2 * n - i - 1 : i); }

// One thread per output pixel component.  out(r,c) = LP(mirror(2r), mirror(2c)); LP's V pass reads
// H rows with mirrored indices (the mirror-filled temp of pyramid.hh:36), H reads in(row, x-2..x+2)
// from the image's own (caller-filled) column border.
template <int KIND>
__global__ void __launch_bounds__(256) k_lowpass_sub2(Img in, Img out, int step) {
  typedef typename LpT<KIND>::elem E;
  typedef typename LpT<KIND>::acc A;
  constexpr int COMPS = LpT<KIND>::comps;
  long long total = (long long)out.nrows * out.ncols * COMPS;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % COMPS);
    const long long pix = i / COMPS;
    const int r = (int)(pix / out.ncols);
    const int c = (int)(pix - (long long)r * out.ncols);
    const int y = mirror_idx(r * step, in.nrows);
    const int x = mirror_idx(c * step, in.ncols);
    A h[5];
#pragma unroll
    for (int d = 0; d < 5; d++) {
      const int yy = mirror_idx(y - 2 + d, in.nrows);
      const E* row = row_ptr<E>(in, yy);
      h[d] = (A)(E)lp5((A)row[(x - 2) * COMPS + k], (A)row[(x - 1) * COMPS + k], (A)row[x * COMPS + k], (A)row[(x + 1) * COMPS + k],
                       (A)row[(x + 2) * COMPS + k]);
    }
    row_ptr<E>(out, r)[c * COMPS + k] = (E)lp5(h[0], h[1], h[2], h[3], h[4]);
  }
}

static int grid_for(long long items, int threads) {
  long long blocks = (items + threads - 1) / threads;
  long long cap = (long long)sm_count() * 16;
  if (blocks < 1) blocks = 1;
  return (int)(blocks < cap ? blocks : cap);
}

}  // namespace vppb

using namespace vppb;

extern "C" {

int vppb_scharr_u8(const vppb_img* in, const vppb_img* out, int as_float, void* stream) {
  VPPB_REQUIRE(in && out && in->base && out->base, VPPB_E_ARG, "vppb_scharr_u8: NULL image");
  VPPB_REQUIRE(in->elem_bytes == 1 && out->elem_bytes == 8, VPPB_E_ARG, "vppb_scharr_u8: needs u8 input and 8-byte output elements");
  VPPB_REQUIRE(in->nrows >= out->nrows && in->ncols >= out->ncols, VPPB_E_ARG, "vppb_scharr_u8: input smaller than output");
  VPPB_REQUIRE(in->border >= 1, VPPB_E_BORDER, "vppb_scharr_u8: input border %d < 1", in->border);
  VPPB_REQUIRE(((uintptr_t)out->base % 8) == 0 && (out->pitch % 8) == 0, VPPB_E_ARG, "vppb_scharr_u8: output not 8-byte aligned");
  const int groups = (out->ncols + 3) / 4;
  const int grid = grid_for((long long)out->nrows * groups, 256);
  if (as_float) k_scharr_u8<true><<<grid, 256, 0, as_stream(stream)>>>(view(in), view(out), groups);
  else k_scharr_u8<false><<<grid, 256, 0, as_stream(stream)>>>(view(in), view(out), groups);
  VPPB_LAUNCH_CHECK("vppb_scharr_u8");
  return VPPB_OK;
}

int vppb_lowpass_sub2(const vppb_img* in, const vppb_img* out, int kind, void* stream) {
  VPPB_REQUIRE(in && out && in->base && out->base, VPPB_E_ARG, "vppb_lowpass_sub2: NULL image");
  VPPB_REQUIRE(kind >= 0 && kind <= 2, VPPB_E_ARG, "vppb_lowpass_sub2: kind %d", kind);
  const int e = kind == 0 ? 1 : 8;
  VPPB_REQUIRE(in->elem_bytes == e && out->elem_bytes == e, VPPB_E_ARG, "vppb_lowpass_sub2: element size must be %d for kind %d", e, kind);
  VPPB_REQUIRE(in->border >= 2, VPPB_E_BORDER, "vppb_lowpass_sub2: input border %d < 2", in->border);
  // pyramid.hh:140,154: level size 1 + n/2; any smaller output is a prefix of it
  VPPB_REQUIRE(out->nrows <= 1 + in->nrows / 2 && out->ncols <= 1 + in->ncols / 2, VPPB_E_ARG,
               "vppb_lowpass_sub2: output %dx%d larger than 1+n/2 of input %dx%d", out->nrows, out->ncols, in->nrows, in->ncols);
  cudaStream_t st = as_stream(stream);
  const long long items = (long long)out->nrows * out->ncols * (kind == 0 ? 1 : 2);
  const int grid = grid_for(items, 256);
  if (kind == 0) k_lowpass_sub2<0><<<grid, 256, 0, st>>>(view(in), view(out), 2);
  else if (kind == 1) k_lowpass_sub2<1><<<grid, 256, 0, st>>>(view(in), view(out), 2);
  else k_lowpass_sub2<2><<<grid, 256, 0, st>>>(view(in), view(out), 2);
  VPPB_LAUNCH_CHECK("vppb_lowpass_sub2");
  return VPPB_OK;
}

}  // extern "C"
