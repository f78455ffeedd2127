// FAST-9 corner detector (16-pixel ring, >= 9 contiguous brighter-or-darker pixels).
// Reference: vpp/algorithms/fast_detector/fast.hpp:253-508 (SIMD pruning tree == the 9-arc
// predicate on the ring *as implemented*, whose slots 4 and 12 are sampled on row r-3),
// :36-77 (score, true ring), :663-673 / :889-928 / :745-799 (plain, local-maxima, blockwise),
// mask semantics :310-317 (mask byte bit 4 gates the brighter arc, bit 0 the darker arc).
//
// GPU structure: pass 1 puts one warp on 32 consecutive pixels of a row; the two "necessary"
// ring pixels (top/bottom) are tested first and __any_sync lets the whole warp skip the other
// 14 loads when no lane can be a corner; the corner flags of the 32 lanes are collected with
// __ballot_sync into one word of a bitmask (1 bit / pixel).  Maxima modes refine that bitmask.
// Pass 2 scans per-row counts and emits keypoints in raster order (deterministic, unlike the
// reference's per-thread buffers flushed under `omp critical`).
// HBM-bound: 1 byte read per pixel (+1 with a mask) + 8 bytes per keypoint.
#include "common.cuh"

namespace vppb {

// ring slot -> (dr, dc).  Row 0: as implemented by fast9() (fast.hpp:327-461); row 1: true ring.
__constant__ signed char c_ring[2][16][2] = {
    {{-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}, {-3, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3}, {-3, -3}, {-1, -3}, {-2, -2}, {-3, -1}},
    {{-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}, {0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3}, {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}}};

__device__ __forceinline__ bool arc9(uint32_t m16) {
  uint32_t x = m16 | (m16 << 16);
  x &= x >> 1;
  x &= x >> 2;
  x &= x >> 4;  // 8 contiguous
  x &= x >> 1;  // 9 contiguous
  return x != 0;
}

// fast.hpp:36-77: sum of |v - a| over ring pixels beyond the threshold, max of the two polarities
__device__ __forceinline__ int fast9_score_at(const Img& im, int r, int c, int th) {
  const unsigned char* p = im.base + (long long)r * im.pitch + c;
  const int v = *p;
  int sum_inf = 0, sum_sup = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int a = p[(long long)c_ring[1][i][0] * im.pitch + c_ring[1][i][1]];
    const int diff = v - a;
    if (diff < -th) sum_inf -= diff;
    else if (diff > th) sum_sup += diff;
  }
  return max(sum_sup, sum_inf);
}

// pass 1: detection -> bitmask words (row-major, wpr words per row) and per-row counts
__global__ void __launch_bounds__(256) k_fast9_detect(Img im, Img mask, int has_mask, int th, int ring, uint32_t* bits, int wpr, int* rowcount) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long total = (long long)im.nrows * wpr;
  const int thb = th & 255;  // S::repeat(th) replicates the low byte
  for (long long w = warp0; w < total; w += nwarps) {
    const int r = (int)(w / wpr);
    const int c = (int)(w - (long long)r * wpr) * 32 + lane;
    const bool inside = c < im.ncols;
    const unsigned char* p = im.base + (long long)r * im.pitch + c;
    int m = 0xFF;
    if (has_mask) m = inside ? (int)__ldg(mask.base + (long long)r * mask.pitch + c) : 0;
    int v = 0, hi = 0, lo = 0;
    bool cand_b = false, cand_d = false;
    if (inside && m != 0) {
      v = __ldg(p);
      hi = min(v + thb, 255);  // u_adds
      lo = max(v - thb, 0);    // u_subs
      const int a0 = __ldg(p + (long long)c_ring[ring][0][0] * im.pitch + c_ring[ring][0][1]);
      const int a8 = __ldg(p + (long long)c_ring[ring][8][0] * im.pitch + c_ring[ring][8][1]);
      // every 9-arc of the 16-ring contains slot 0 or slot 8 (fast.hpp:326-337)
      cand_b = (m & 0x10) && (a0 > hi || a8 > hi);
      cand_d = (m & 0x01) && (a0 < lo || a8 < lo);
    }
    bool corner = false;
    if (__any_sync(0xffffffffu, cand_b || cand_d)) {
      if (cand_b || cand_d) {
        uint32_t mb = 0, md = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const int a = __ldg(p + (long long)c_ring[ring][i][0] * im.pitch + c_ring[ring][i][1]);
          mb |= (uint32_t)(a > hi) << i;
          md |= (uint32_t)(a < lo) << i;
        }
        corner = (cand_b && arc9(mb)) || (cand_d && arc9(md));
      }
    }
    const uint32_t word = __ballot_sync(0xffffffffu, corner);
    if (lane == 0) {
      bits[w] = word;
      if (word) atomicAdd(&rowcount[r], __popc(word));
    }
  }
}

__device__ __forceinline__ bool bit_at(const uint32_t* bits, int wpr, int nrows, int ncols, int r, int c) {
  if (r < 0 || r >= nrows || c < 0 || c >= ncols) return false;
  return (bits[(long long)r * wpr + (c >> 5)] >> (c & 31)) & 1u;
}

// local maxima (fast.hpp:889-928): keep a detected corner iff its u8 score (score/16) is strictly
// greater than the 8 neighbours' entries of the score image (0 where no corner was detected).
__global__ void __launch_bounds__(256) k_fast9_local_max(Img im, int th, const uint32_t* bits, uint32_t* bits_out, int wpr, int* rowcount) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long total = (long long)im.nrows * wpr;
  for (long long w = warp0; w < total; w += nwarps) {
    const uint32_t word = bits[w];
    bool keep = false;
    if (word) {  // warp-uniform
      const int r = (int)(w / wpr);
      const int c = (int)(w - (long long)r * wpr) * 32 + lane;
      if ((word >> lane) & 1u) {
        const int a = (fast9_score_at(im, r, c, th) / 16) & 255;
        keep = true;
#pragma unroll
        for (int dr = -1; dr <= 1; dr++)
#pragma unroll
          for (int dc = -1; dc <= 1; dc++) {
            if (dr == 0 && dc == 0) continue;
            int n = 0;
            if (bit_at(bits, wpr, im.nrows, im.ncols, r + dr, c + dc)) n = (fast9_score_at(im, r + dr, c + dc, th) / 16) & 255;
            keep = keep && (a > n);
          }
      }
    }
    const uint32_t out = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) {
      bits_out[w] = out;
      if (out) atomicAdd(&rowcount[(int)(w / wpr)], __popc(out));
    }
  }
}

// blockwise maxima (fast.hpp:745-799): per block_size x block_size cell anchored at (0,0), raster
// scan, strict '>' (first maximum wins), kept iff max > 0.  One thread per cell; bits_out zeroed by caller.
__global__ void __launch_bounds__(128) k_fast9_block_max(Img im, int th, int bs, const uint32_t* bits, int* cellkp, int wpr,
                                                        int* rowcount, int cells_r, int cells_c) {
  const long long total = (long long)cells_r * cells_c;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r0 = (int)(i / cells_c) * bs, c0 = (int)(i % cells_c) * bs;
    unsigned vmax = 0;
    int pr = -1, pc = -1;
    for (int r = r0; r < min(r0 + bs, im.nrows); r++) {
      int c = c0;
      const int cend = min(c0 + bs, im.ncols);
      while (c < cend) {
        const int wi = c >> 5;
        uint32_t word = bits[(long long)r * wpr + wi] >> (c & 31);
        const int span = min(32 - (c & 31), cend - c);
        if (span < 32) word &= (1u << span) - 1u;
        while (word) {
          const int b = __ffs(word) - 1;
          word &= word - 1;
          const unsigned v = (unsigned)((fast9_score_at(im, r, c + b, th) / 16) & 255);
          if (v > vmax) { vmax = v; pr = r; pc = c + b; }
        }
        c += span;
      }
    }
    cellkp[i] = vmax > 0 ? ((pr << 16) | pc) : -1;  // one keypoint at most per cell; emitted in CELL raster order
    if (vmax > 0) atomicAdd(&rowcount[(int)(i / cells_c)], 1);
  }
}

// exclusive scan of the per-row counts; single CTA. rowoff[nrows] = total.
__global__ void __launch_bounds__(1024) k_fast9_scan(const int* rowcount, int* rowoff, int nrows) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nrows; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nrows ? rowcount[i] : 0;
    int x = v;
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int s = warp_sums[threadIdx.x];
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, s, o);
        if (threadIdx.x >= o) s += y;
      }
      warp_sums[threadIdx.x] = s;
    }
    __syncthreads();
    const int warp_prefix = (threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0;
    const int incl = x + warp_prefix + carry;
    if (i < nrows) rowoff[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) rowoff[nrows] = carry;
}

// pass 2: one warp per row walks the bitmask and writes keypoints (and scores) in raster order
__global__ void __launch_bounds__(256) k_fast9_emit(Img im, int th, const uint32_t* bits, int wpr, const int* rowoff, vppb_int2* kps,
                                                   int* scores, int score_div, int capacity) {
  const int lane = threadIdx.x & 31;
  const int warp0 = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int nwarps = (int)(((long long)gridDim.x * blockDim.x) >> 5);
  for (int r = warp0; r < im.nrows; r += nwarps) {
    int off = rowoff[r];
    if (rowoff[r + 1] == off) continue;
    for (int w0 = 0; w0 < wpr; w0 += 32) {
      const int wi = w0 + lane;
      uint32_t word = wi < wpr ? bits[(long long)r * wpr + wi] : 0u;
      const int cnt = __popc(word);
      int incl = cnt;
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      int pos = off + incl - cnt;
      while (word) {
        const int b = __ffs(word) - 1;
        word &= word - 1;
        if (pos < capacity) {
          const int c = wi * 32 + b;
          kps[pos].r = r;
          kps[pos].c = c;
          if (scores) {
            const int s = fast9_score_at(im, r, c, th);
            scores[pos] = score_div ? ((s / 16) & 255) : s;
          }
        }
        pos++;
      }
      off += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
}

// blockwise emit: one warp per row of cells, keypoints leave in cell raster order (the serial order of fast.hpp:763-790)
__global__ void __launch_bounds__(256) k_fast9_emit_cells(Img im, int th, const int* cellkp, int cells_r, int cells_c, const int* rowoff, vppb_int2* kps,
                                                         int* scores, int capacity) {
  const int lane = threadIdx.x & 31;
  const int warp0 = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int nwarps = (int)(((long long)gridDim.x * blockDim.x) >> 5);
  for (int cr = warp0; cr < cells_r; cr += nwarps) {
    int off = rowoff[cr];
    if (rowoff[cr + 1] == off) continue;
    for (int c0 = 0; c0 < cells_c; c0 += 32) {
      const int cc = c0 + lane;
      const int v = cc < cells_c ? cellkp[(long long)cr * cells_c + cc] : -1;
      const unsigned m = __ballot_sync(0xffffffffu, v >= 0);
      if (v >= 0) {
        const int pos = off + __popc(m & ((1u << lane) - 1u));
        if (pos < capacity) {
          const int r = v >> 16, c = v & 0xFFFF;
          kps[pos].r = r; kps[pos].c = c;
          if (scores) scores[pos] = (fast9_score_at(im, r, c, th) / 16) & 255;
        }
      }
      off += __popc(m);
    }
  }
}

__global__ void k_fast9_scores(Img im, int th, const vppb_int2* kps, int n, int* scores) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    scores[i] = fast9_score_at(im, kps[i].r, kps[i].c, th);
}

struct FastWs {
  uint32_t* bits_a;
  uint32_t* bits_b;
  int* rowcount;
  int* rowoff;
  int* cellkp;
  long long bytes;
};

static FastWs fast_ws_layout(void* base, int nrows, int ncols, int block_size) {
  FastWs w;
  const long long wpr = (ncols + 31) / 32;
  const long long bits_bytes = ((long long)nrows * wpr * 4 + 255) / 256 * 256;
  const long long rows_bytes = (((long long)nrows + 1) * 4 + 255) / 256 * 256;
  unsigned char* p = static_cast<unsigned char*>(base);
  w.bits_a = reinterpret_cast<uint32_t*>(p);
  w.bits_b = reinterpret_cast<uint32_t*>(p + bits_bytes);
  w.rowcount = reinterpret_cast<int*>(p + 2 * bits_bytes);
  w.rowoff = reinterpret_cast<int*>(p + 2 * bits_bytes + rows_bytes);
  w.cellkp = reinterpret_cast<int*>(p + 2 * bits_bytes + 2 * rows_bytes);
  const int bs = block_size > 0 ? block_size : 10;
  const long long cells = (long long)((nrows + bs - 1) / bs) * ((ncols + bs - 1) / bs);
  w.bytes = 2 * bits_bytes + 2 * rows_bytes + ((cells * 4 + 255) / 256) * 256;
  return w;
}

}  // namespace vppb

using namespace vppb;

extern "C" {

int64_t vppb_fast9_workspace_bytes(int32_t nrows, int32_t ncols, int32_t block_size) {
  if (nrows <= 0 || ncols <= 0) return 0;
  return fast_ws_layout(nullptr, nrows, ncols, block_size).bytes;
}

int vppb_fast9_u8(const vppb_img* img, int32_t th, const vppb_img* mask, int32_t mode, int32_t block_size, int32_t ring,
                  void* workspace, int64_t workspace_bytes, vppb_int2* kps_out, int32_t* scores_out, int32_t capacity,
                  int32_t* count_out, void* stream) {
  VPPB_REQUIRE(img && img->base && workspace && count_out, VPPB_E_ARG, "vppb_fast9_u8: NULL argument");
  VPPB_REQUIRE(img->elem_bytes == 1, VPPB_E_ARG, "vppb_fast9_u8: image must be u8");
  // fast.hpp:937-938
  VPPB_REQUIRE(img->border >= 3, VPPB_E_BORDER, "Image need a border of 3px at least for the FAST detector");
  VPPB_REQUIRE(mode >= 0 && mode <= 2 && (ring == 0 || ring == 1), VPPB_E_ARG, "vppb_fast9_u8: bad mode/ring");
  VPPB_REQUIRE(mode != VPPB_FAST_BLOCKWISE || block_size > 0, VPPB_E_ARG, "vppb_fast9_u8: block_size must be > 0");
  VPPB_REQUIRE(capacity == 0 || kps_out, VPPB_E_ARG, "vppb_fast9_u8: NULL keypoint buffer");
  const bool has_mask = mask && mask->base;
  if (has_mask)
    VPPB_REQUIRE(mask->elem_bytes == 1 && mask->nrows >= img->nrows && mask->ncols >= img->ncols, VPPB_E_ARG,
                 "vppb_fast9_u8: mask must be u8 and cover the image");
  VPPB_REQUIRE(img->nrows < 65536 && img->ncols < 65536, VPPB_E_ARG, "vppb_fast9_u8: image larger than 65535 in one dimension");
  FastWs ws = fast_ws_layout(workspace, img->nrows, img->ncols, mode == VPPB_FAST_BLOCKWISE ? block_size : 10);
  VPPB_REQUIRE(workspace_bytes >= ws.bytes, VPPB_E_ARG, "vppb_fast9_u8: workspace %lld < %lld bytes", (long long)workspace_bytes, ws.bytes);
  cudaStream_t st = as_stream(stream);
  const int wpr = (img->ncols + 31) / 32;
  const long long words = (long long)img->nrows * wpr;
  Img im = view(img);
  Img mk = has_mask ? view(mask) : im;
  const int sms = sm_count();

  VPPB_CUDA(cudaMemsetAsync(ws.rowcount, 0, ((size_t)img->nrows + 1) * sizeof(int), st));
  {
    long long blocks = (words + 7) / 8;  // 8 warps per CTA
    int grid = (int)(blocks < (long long)sms * 8 ? blocks : (long long)sms * 8);
    k_fast9_detect<<<grid, 256, 0, st>>>(im, mk, has_mask ? 1 : 0, th, ring, ws.bits_a, wpr, ws.rowcount);
  }
  const uint32_t* final_bits = ws.bits_a;
  if (mode != VPPB_FAST_ALL) {
    VPPB_CUDA(cudaMemsetAsync(ws.rowcount, 0, ((size_t)img->nrows + 1) * sizeof(int), st));
    if (mode == VPPB_FAST_LOCAL_MAXIMA) {
      long long blocks = (words + 7) / 8;
      int grid = (int)(blocks < (long long)sms * 8 ? blocks : (long long)sms * 8);
      k_fast9_local_max<<<grid, 256, 0, st>>>(im, th, ws.bits_a, ws.bits_b, wpr, ws.rowcount);
    } else {
      const int cells_r = (img->nrows + block_size - 1) / block_size, cells_c = (img->ncols + block_size - 1) / block_size;
      long long cells = (long long)cells_r * cells_c;
      long long blocks = (cells + 127) / 128;
      int grid = (int)(blocks < (long long)sms * 16 ? blocks : (long long)sms * 16);
      k_fast9_block_max<<<grid, 128, 0, st>>>(im, th, block_size, ws.bits_a, ws.cellkp, wpr, ws.rowcount, cells_r, cells_c);
      k_fast9_scan<<<1, 1024, 0, st>>>(ws.rowcount, ws.rowoff, cells_r);
      long long eb = ((long long)cells_r + 7) / 8;
      k_fast9_emit_cells<<<(int)(eb < (long long)sms * 8 ? eb : (long long)sms * 8), 256, 0, st>>>(im, th, ws.cellkp, cells_r, cells_c, ws.rowoff, kps_out, scores_out,
                                                                                                 capacity);
      VPPB_LAUNCH_CHECK("vppb_fast9_u8");
      int total = 0;
      VPPB_CUDA(cudaMemcpyAsync(&total, ws.rowoff + cells_r, sizeof(int), cudaMemcpyDeviceToHost, st));
      VPPB_CUDA(cudaStreamSynchronize(st));
      *count_out = total;
      VPPB_REQUIRE(total <= capacity, VPPB_E_CAPACITY, "vppb_fast9_u8: %d keypoints exceed the capacity %d", total, capacity);
      return VPPB_OK;
    }
    final_bits = ws.bits_b;
  }
  k_fast9_scan<<<1, 1024, 0, st>>>(ws.rowcount, ws.rowoff, img->nrows);
  {
    long long blocks = ((long long)img->nrows + 7) / 8;
    int grid = (int)(blocks < (long long)sms * 8 ? blocks : (long long)sms * 8);
    k_fast9_emit<<<grid, 256, 0, st>>>(im, th, final_bits, wpr, ws.rowoff, kps_out, scores_out, mode != VPPB_FAST_ALL ? 1 : 0, capacity);
  }
  VPPB_LAUNCH_CHECK("vppb_fast9_u8");
  int total = 0;
  VPPB_CUDA(cudaMemcpyAsync(&total, ws.rowoff + img->nrows, sizeof(int), cudaMemcpyDeviceToHost, st));
  VPPB_CUDA(cudaStreamSynchronize(st));
  *count_out = total;
  VPPB_REQUIRE(total <= capacity, VPPB_E_CAPACITY, "vppb_fast9_u8: %d keypoints exceed the capacity %d", total, capacity);
  return VPPB_OK;
}

int vppb_fast9_scores(const vppb_img* img, int32_t th, const vppb_int2* kps, int32_t n, int32_t* scores_out, void* stream) {
  VPPB_REQUIRE(img && img->base && (n == 0 || (kps && scores_out)), VPPB_E_ARG, "vppb_fast9_scores: NULL argument");
  VPPB_REQUIRE(img->elem_bytes == 1, VPPB_E_ARG, "vppb_fast9_scores: image must be u8");
  VPPB_REQUIRE(img->border >= 3, VPPB_E_BORDER, "vppb_fast9_scores: border %d < 3", img->border);
  if (n == 0) return VPPB_OK;
  int grid = (n + 255) / 256;
  k_fast9_scores<<<grid, 256, 0, as_stream(stream)>>>(view(img), th, kps, n, scores_out);
  VPPB_LAUNCH_CHECK("vppb_fast9_scores");
  return VPPB_OK;
}

}  // extern "C"
