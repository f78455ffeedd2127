// The remaining 3x3 stencils of the path: lbp_transform and local_maxima_filter.
// Reference: vpp/algorithms/lbp/lbp_transform.hh:7-38, vpp/algorithms/fast_detector/fast.hpp:555-575.
#include "common.cuh"

namespace vppb {

// ---- lbp_transform (unsigned char -> unsigned char) ---------------------------------------------------------------------------
// bit k of out(r, c) = neighbour k > centre, neighbours in the order (-1,-1) (-1,0) (-1,1) (0,-1) (0,1) (1,-1) (1,0) (1,1).
// HBM-bound (1 B in + 1 B out per pixel).  A thread owns 16 consecutive pixels of LBP_ROWS consecutive rows: it slides a window of
// three 16-byte row vectors down the image (each input row is loaded once per thread), forms the left / right neighbours of a row by
// funnel shifts across the four words (+ one byte on either side) and evaluates four pixels per packed byte compare.
constexpr int LBP_ROWS = 4;

struct LbpRow { uint32_t w[4]; uint32_t left, right; };  // bytes x0 .. x0+15, the byte before (in bits 24-31) and the byte after (bits 0-7)

__device__ __forceinline__ LbpRow lbp_load(const unsigned char* row, int x0) {
  LbpRow v;
  const uint4 q = *reinterpret_cast<const uint4*>(row + x0);
  v.w[0] = q.x; v.w[1] = q.y; v.w[2] = q.z; v.w[3] = q.w;
  v.left = (uint32_t)row[x0 - 1] << 24;
  v.right = (uint32_t)row[x0 + 16];
  return v;
}
// word k of the row shifted so that byte i holds pixel i - 1 (L) / pixel i + 1 (R)
__device__ __forceinline__ uint32_t lbp_l(const LbpRow& v, int k) { return __funnelshift_l(k ? v.w[k - 1] : v.left, v.w[k], 8); }
__device__ __forceinline__ uint32_t lbp_r(const LbpRow& v, int k) { return __funnelshift_r(v.w[k], k < 3 ? v.w[k + 1] : v.right, 8); }

__device__ __forceinline__ unsigned char lbp_scalar(const unsigned char* r0, const unsigned char* r1, const unsigned char* r2, int i) {
  const int c = r1[i];
  return (unsigned char)(((r0[i - 1] > c) << 0) + ((r0[i] > c) << 1) + ((r0[i + 1] > c) << 2) + ((r1[i - 1] > c) << 3) + ((r1[i + 1] > c) << 4) +
                         ((r2[i - 1] > c) << 5) + ((r2[i] > c) << 6) + ((r2[i + 1] > c) << 7));
}

template <bool VEC>
__global__ void __launch_bounds__(256) k_lbp_u8(Img in, Img out, int xthreads) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int xt = (int)(tid % xthreads);
  const int r_first = (int)(tid / xthreads) * LBP_ROWS;
  if (r_first >= in.nrows) return;
  const int x0 = xt * 16;
  const int r_end = min(r_first + LBP_ROWS, in.nrows);
  if (VEC && x0 + 16 <= in.ncols) {
    // all LBP_ROWS + 2 input rows of the thread are requested before the first is used: one memory latency per thread, not one per row
    LbpRow v[LBP_ROWS + 2];
#pragma unroll
    for (int j = 0; j < LBP_ROWS + 2; j++) {
      const int r = min(r_first - 1 + j, in.nrows);  // rows below the bottom border row are not read (nor used)
      v[j] = lbp_load(in.base + (long long)r * in.pitch, x0);
    }
#pragma unroll
    for (int j = 0; j < LBP_ROWS; j++) {
      const int r = r_first + j;
      if (r >= r_end) break;
      const LbpRow &a = v[j], &b = v[j + 1], &c = v[j + 2];
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t ctr = b.w[k];
        uint32_t acc = __vcmpgtu4(lbp_l(a, k), ctr) & 0x01010101u;
        acc |= __vcmpgtu4(a.w[k], ctr) & 0x02020202u;
        acc |= __vcmpgtu4(lbp_r(a, k), ctr) & 0x04040404u;
        acc |= __vcmpgtu4(lbp_l(b, k), ctr) & 0x08080808u;
        acc |= __vcmpgtu4(lbp_r(b, k), ctr) & 0x10101010u;
        acc |= __vcmpgtu4(lbp_l(c, k), ctr) & 0x20202020u;
        acc |= __vcmpgtu4(c.w[k], ctr) & 0x40404040u;
        acc |= __vcmpgtu4(lbp_r(c, k), ctr) & 0x80808080u;
        o[k] = acc;
      }
      *reinterpret_cast<uint4*>(out.base + (long long)r * out.pitch + x0) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    return;
  }
  for (int r = r_first; r < r_end; r++) {  // ragged right edge, unaligned images
    const unsigned char* r1 = in.base + (long long)r * in.pitch;
    unsigned char* o = out.base + (long long)r * out.pitch;
    for (int i = x0; i < min(x0 + 16, in.ncols); i++) o[i] = lbp_scalar(r1 - in.pitch, r1, r1 + in.pitch, i);
  }
}

// ---- local_maxima_filter (in place) ---------------------------------------------------------------------------------------------
// The reference zeroes every pixel that is not strictly greater than its 8 neighbours IN PLACE, so a pixel sees the already
// filtered values of the neighbours above and to the left (serial raster order; its pixel_wise runs rows in parallel and races).
// The result of the serial order is the unique solution X of
//     X(p) = orig(p)  if  orig(p) > max(X(NW), X(N), X(NE), X(W), orig(E), orig(SW), orig(S), orig(SE)),  else 0
// - a triangular system (X(p) only depends on X of earlier pixels), solved here by relaxation exactly like the sweeps of the
// semi-dense flow: every pass re-evaluates all pixels on the current X of their predecessors (runs of 16 pixels of a row are
// walked serially, so chains along a row shorten 16-fold), passes are separated by a
// grid-wide barrier, and a pass that changes nothing has reached the fixed point.  X only ever takes the values orig(p) and 0, the
// number of passes is the longest chain of decisions that flip (2 - 4 on score images, the length of a ramp at worst).
template <typename T> __device__ __forceinline__ T ld_cg(const T* p) { return __ldcg(p); }

constexpr int LMF_RUN = 16;
template <typename T>
__global__ void __launch_bounds__(256) k_local_maxima_filter(Img im, T* orig, int wcols, int* ctr) {
  const long long gthreads = (long long)gridDim.x * blockDim.x, gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int gen = 0;
  // orig <- the image with its border of 1 (tight rows of wcols = ncols + 2 elements)
  const long long wtotal = (long long)(im.nrows + 2) * wcols;
  for (long long i = gtid; i < wtotal; i += gthreads) {
    const int r = (int)(i / wcols) - 1, c = (int)(i % wcols) - 1;
    orig[i] = reinterpret_cast<const T*>(im.base + (long long)r * im.pitch)[c];
  }
  grid_barrier(ctr, gen);
  // a thread owns a run of LMF_RUN consecutive pixels of a row and walks it left to right with its own fresh result as the left
  // neighbour (Gauss-Seidel inside the run): a horizontal chain of dependent decisions costs one pass per run it crosses, not one per pixel
  const int runs_per_row = (im.ncols + LMF_RUN - 1) / LMF_RUN;
  const long long total = (long long)im.nrows * runs_per_row;
  for (int pass = 0;; pass++) {
    int changed = 0;
    for (long long i = gtid; i < total; i += gthreads) {
      const int r = (int)(i / runs_per_row), c0 = (int)(i % runs_per_row) * LMF_RUN;
      const int c1 = min(c0 + LMF_RUN, im.ncols);
      const T* o1 = orig + (long long)(r + 1) * wcols + (c0 + 1);   // orig(r, c0)
      const T* o2 = o1 + wcols;                                       // orig(r + 1, c0)
      T* x1 = reinterpret_cast<T*>(im.base + (long long)r * im.pitch) + c0;
      const T* x0 = reinterpret_cast<const T*>(im.base + (long long)(r - 1) * im.pitch) + c0;
      // successors: original values; predecessors: their current X (read from L2: other CTAs rewrite it between passes)
      T w = ld_cg(x1 - 1);
      T nl = ld_cg(x0 - 1), nm = ld_cg(x0);
      T sl = o2[-1], sm = o2[0];
      T a = o1[0];
      for (int c = c0; c < c1; c++, o1++, o2++, x0++, x1++) {
        const T nr = ld_cg(x0 + 1), sr = o2[1], e = o1[1];
        const bool is_max = a > e && a > sl && a > sm && a > sr && a > nl && a > nm && a > nr && a > w;
        const T res = is_max ? a : (T)0;
        if (ld_cg(x1) != res) { *x1 = res; changed = 1; }
        w = res; nl = nm; nm = nr; sl = sm; sm = sr; a = e;
      }
    }
    if (__any_sync(0xffffffffu, changed) && (threadIdx.x & 31) == 0) atomicAdd(ctr + 1 + pass % 3, 1);
    if (gtid == 0) ctr[1 + (pass + 1) % 3] = 0;  // next pass's counter: nobody reads or writes it during this pass
    grid_barrier(ctr, gen);
    if (ld_cg(ctr + 1 + pass % 3) == 0) break;  // grid-uniform: the counter is stable until every CTA has passed the next barrier
  }
}

}  // namespace vppb

using namespace vppb;

extern "C" {

int vppb_lbp_u8(const vppb_img* in, const vppb_img* out, void* stream) {
  VPPB_REQUIRE(in && out && in->base && out->base, VPPB_E_ARG, "vppb_lbp_u8: NULL argument");
  VPPB_REQUIRE(in->elem_bytes == 1 && out->elem_bytes == 1 && same_domain(in, out), VPPB_E_ARG, "vppb_lbp_u8: u8 images of one domain expected");
  VPPB_REQUIRE(in->border >= 1, VPPB_E_BORDER, "vppb_lbp_u8: the input needs a border of 1 pixel");
  if (in->nrows == 0 || in->ncols == 0) return VPPB_OK;
  const bool vec = ((uintptr_t)in->base % 16) == 0 && (in->pitch % 16) == 0 && ((uintptr_t)out->base % 16) == 0 && (out->pitch % 16) == 0;
  const int xthreads = (in->ncols + 15) / 16;
  const long long threads = (long long)xthreads * ((in->nrows + LBP_ROWS - 1) / LBP_ROWS);
  const int grid = (int)((threads + 255) / 256);
  if (vec) k_lbp_u8<true><<<grid, 256, 0, as_stream(stream)>>>(view(in), view(out), xthreads);
  else k_lbp_u8<false><<<grid, 256, 0, as_stream(stream)>>>(view(in), view(out), xthreads);
  VPPB_LAUNCH_CHECK("vppb_lbp_u8");
  return VPPB_OK;
}

int64_t vppb_local_maxima_filter_workspace_bytes(int32_t nrows, int32_t ncols, int32_t elem_bytes) {
  if (nrows < 0 || ncols < 0 || (elem_bytes != 1 && elem_bytes != 4)) return 0;
  return (((int64_t)(nrows + 2) * (ncols + 2) * elem_bytes + 255) / 256) * 256 + 256;
}

int vppb_local_maxima_filter(const vppb_img* img, void* workspace, int64_t workspace_bytes, void* stream) {
  VPPB_REQUIRE(img && img->base && workspace, VPPB_E_ARG, "vppb_local_maxima_filter: NULL argument");
  VPPB_REQUIRE(img->elem_bytes == 1 || img->elem_bytes == 4, VPPB_E_ARG, "vppb_local_maxima_filter: unsigned char or int pixels expected");
  VPPB_REQUIRE(img->border >= 1, VPPB_E_BORDER, "vppb_local_maxima_filter: the image needs a border of 1 pixel");
  const int64_t need = vppb_local_maxima_filter_workspace_bytes(img->nrows, img->ncols, img->elem_bytes);
  VPPB_REQUIRE(workspace_bytes >= need, VPPB_E_ARG, "vppb_local_maxima_filter: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)need);
  if (img->nrows == 0 || img->ncols == 0) return VPPB_OK;
  cudaStream_t st = as_stream(stream);
  int* ctr = reinterpret_cast<int*>(static_cast<unsigned char*>(workspace) + need - 256);
  VPPB_CUDA(cudaMemsetAsync(ctr, 0, 256, st));
  const long long total = (long long)img->nrows * ((img->ncols + LMF_RUN - 1) / LMF_RUN);
  long long blocks = (total + 255) / 256;
  if (img->elem_bytes == 1) {
    const int cap = cooperative_grid_limit(k_local_maxima_filter<unsigned char>, 256);
    if (blocks > cap) blocks = cap;
    VPPB_CUDA(launch_cooperative(k_local_maxima_filter<unsigned char>, (int)blocks, 256, st, view(img), static_cast<unsigned char*>(workspace), img->ncols + 2, ctr));
  } else {
    const int cap = cooperative_grid_limit(k_local_maxima_filter<int>, 256);
    if (blocks > cap) blocks = cap;
    VPPB_CUDA(launch_cooperative(k_local_maxima_filter<int>, (int)blocks, 256, st, view(img), static_cast<int*>(workspace), img->ncols + 2, ctr));
  }
  VPPB_LAUNCH_CHECK("vppb_local_maxima_filter");
  return VPPB_OK;
}

}  // extern "C"
