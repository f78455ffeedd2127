// 5x5 box stencil:  out = (sum of the 25 neighbours) / 25, per channel, integer division.
// Reference: the relative_access / box_nbh2d user kernels of benchmarks/box_5x5_filter2.cc:71-81
// (image2d<int>) and examples/box_filter.cc:23-32 (image2d<vuchar3>, vint3 accumulator).
//
// Byte images (u8, vuchar3) run as byte streams: a vuchar3 row is 3*ncols bytes and the
// horizontal taps sit CS = 3 bytes apart.  One CTA streams a 1024-byte-wide column strip
// top-to-bottom: a producer thread issues TMA tile loads (cp.async.bulk.tensor.2d, 10 rows x
// 1056 bytes per stage, 4-stage mbarrier ring), 64 consumer threads each own 16 output bytes,
// do the horizontal 5-tap sum in packed 16-bit lanes, and keep the vertical 5-row window as a
// register ring, so every input byte is fetched from L2/HBM once per strip and the vertical halo
// costs 4 rows per segment.  HBM-bound: 2 bytes of traffic per output byte.
#include "common.cuh"
#include "tma.cuh"

namespace vppb {

// ------------------------------------------------------------------ tensor map encoding
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_tensor_map_2d(CUtensorMap* map, void* origin, CUtensorMapDataType elem, int elem_bytes, uint64_t width,
                         uint64_t height, uint64_t pitch, uint32_t box_w, uint32_t box_h) {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
      set_error("cuTensorMapEncodeTiled entry point unavailable (%d)", (int)e);
      return VPPB_E_CUDA;
    }
    fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  (void)elem_bytes;
  cuuint64_t gdim[2] = {width, height};
  cuuint64_t gstr[1] = {pitch};
  cuuint32_t box[2] = {box_w, box_h};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, elem, 2, origin, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): origin %p w %llu h %llu pitch %llu box %ux%u", (int)r, origin,
              (unsigned long long)width, (unsigned long long)height, (unsigned long long)pitch, box_w, box_h);
    return VPPB_E_CUDA;
  }
  return VPPB_OK;
}

// ------------------------------------------------------------------ direct kernels (any layout)
// One thread per output element; used for images the TMA path cannot describe (unaligned
// external buffers) and as the int32 path.  T = element, CS = tap stride in elements.
template <typename T, typename ACC, int CS>
__global__ void k_box5_direct(Img in, Img out, int row_elems) {
  long long total = (long long)out.nrows * row_elems;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int r = (int)(i / row_elems);
    int x = (int)(i - (long long)r * row_elems);
    ACC s = 0;
#pragma unroll
    for (int dy = -2; dy <= 2; dy++) {
      const T* p = row_ptr<T>(in, r + dy) + x;
#pragma unroll
      for (int dx = -2; dx <= 2; dx++) s += (ACC)p[dx * CS];
    }
    row_ptr<T>(out, r)[x] = (T)(s / 25);
  }
}

// ------------------------------------------------------------------ TMA strip kernel (bytes)
constexpr int BX_TW = 1024;           // output bytes per strip row
constexpr int BX_BOXW = 1056;         // TW + 16 left + 16 right (halo rounded to the 16-byte TMA granule)
constexpr int BX_CH = 10;             // rows per stage: two turns of the 5-row ring
constexpr int BX_STAGES = 4;
constexpr int BX_STAGE_BYTES = 10624; // BX_CH * BX_BOXW rounded up to 128
constexpr int BX_CONSUMERS = 64;
constexpr int BX_THREADS = BX_CONSUMERS + 32;
constexpr unsigned BX_DIV25 = 671089u;  // floor(s/25) == (s * 671089) >> 24 for 0 <= s <= 6375 (checked exhaustively)

// Horizontal 5-tap sums of the 16 bytes a thread owns, as 8 packed 16-bit pairs.
// Lane layout: E pair of word i = (byte0, byte2), O pair = (byte1, byte3).
template <int CS>
__device__ __forceinline__ void box_hsum_row(const unsigned char* srow, uint32_t HE[4], uint32_t HO[4]) {
  const uint2 a = *reinterpret_cast<const uint2*>(srow + 8);
  const uint4 b = *reinterpret_cast<const uint4*>(srow + 16);
  const uint2 c = *reinterpret_cast<const uint2*>(srow + 32);
  const uint32_t w[8] = {a.x, a.y, b.x, b.y, b.z, b.w, c.x, c.y};  // words k-2 .. k+5, own = w[2..5]
  uint32_t E[8], O[8], SE[7], SO[7];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    E[i] = w[i] & 0x00FF00FFu;
    O[i] = __byte_perm(w[i], 0u, 0x4341);
  }
#pragma unroll
  for (int i = 0; i < 7; i++) {
    SE[i] = __byte_perm(E[i], E[i + 1], 0x5432);  // (byte2 of word i, byte0 of word i+1)
    SO[i] = __byte_perm(O[i], O[i + 1], 0x5432);  // (byte3 of word i, byte1 of word i+1)
  }
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int i = j + 2;
    if (CS == 3) {
      // lanes (x, x+2): taps x-6, x-3, x, x+3, x+6
      HE[j] = SE[i - 2] + O[i - 1] + E[i] + SO[i] + SE[i + 1];
      // lanes (x+1, x+3)
      HO[j] = SO[i - 2] + SE[i - 1] + O[i] + E[i + 1] + SO[i + 1];
    } else {
      // CS == 1: taps x-2 .. x+2
      HE[j] = SE[i - 1] + SO[i - 1] + E[i] + O[i] + SE[i];
      HO[j] = SO[i - 1] + E[i] + O[i] + SE[i] + SO[i];
    }
  }
}

__device__ __forceinline__ uint32_t box_div_pack(uint32_t se, uint32_t so) {
  const uint32_t pel = (se & 0xFFFFu) * BX_DIV25, peh = (se >> 16) * BX_DIV25;
  const uint32_t pol = (so & 0xFFFFu) * BX_DIV25, poh = (so >> 16) * BX_DIV25;
  const uint32_t lo = __byte_perm(pel, pol, 0x0073);  // (q(x), q(x+1))
  const uint32_t hi = __byte_perm(peh, poh, 0x0073);  // (q(x+2), q(x+3))
  return __byte_perm(lo, hi, 0x5410);
}

template <int CS>
__global__ void __launch_bounds__(BX_THREADS) k_box5_bytes_tma(const __grid_constant__ CUtensorMap tmap, Img out, int rowbytes,
                                                              int strips, int segs, int seg_chunks, int vec_store) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + BX_STAGES * BX_STAGE_BYTES);
  uint64_t* empty = full + BX_STAGES;
  const int tid = threadIdx.x;
  const int items = strips * segs;
  const int seg_rows = seg_chunks * BX_CH - 4;

  if (tid == 0) {
    for (int s = 0; s < BX_STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], BX_CONSUMERS / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (tid >= BX_CONSUMERS) {
    // ---------------- producer: one elected thread feeds the ring
    if (tid == BX_CONSUMERS) {
      tma_prefetch_desc(&tmap);
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int strip = item % strips, seg = item / strips;
        const int xe = strip * (BX_TW / 8);  // tensor elements are 8 bytes; origin sits 16 bytes left of x = 0
        const int y0 = seg * seg_rows;       // tensor row 0 is image row -2
        for (int ch = 0; ch < seg_chunks; ch++) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], BX_CH * BX_BOXW);
          tma_load_2d(smem + stage * BX_STAGE_BYTES, &tmap, xe, y0 + ch * BX_CH, &full[stage]);
          if (++stage == BX_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  // ---------------- consumers
  int stage = 0;
  uint32_t phase = 0;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int strip = item % strips, seg = item / strips;
    const int x = strip * BX_TW + tid * 16;  // first output byte of this thread
    const int y0 = seg * seg_rows;
    const int y_end = min(y0 + seg_rows, out.nrows);
    uint32_t ringE[5][4], ringO[5][4], VE[4], VO[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      VE[q] = 0; VO[q] = 0;
#pragma unroll
      for (int s = 0; s < 5; s++) { ringE[s][q] = 0; ringO[s][q] = 0; }
    }
    for (int ch = 0; ch < seg_chunks; ch++) {
      mbar_wait(&full[stage], phase);
      const unsigned char* sbase = smem + stage * BX_STAGE_BYTES + tid * 16;
#pragma unroll
      for (int j = 0; j < BX_CH; j++) {
        const int slot = j % 5;
        uint32_t HE[4], HO[4];
        box_hsum_row<CS>(sbase + j * BX_BOXW, HE, HO);
#pragma unroll
        for (int q = 0; q < 4; q++) {
          VE[q] += HE[q] - ringE[slot][q];
          VO[q] += HO[q] - ringO[slot][q];
          ringE[slot][q] = HE[q];
          ringO[slot][q] = HO[q];
        }
        const int in_idx = ch * BX_CH + j;  // input row y0 - 2 + in_idx just entered the window
        const int oy = y0 + in_idx - 4;
        if (in_idx >= 4 && oy < y_end && x < rowbytes) {
          uint4 o;
          o.x = box_div_pack(VE[0], VO[0]);
          o.y = box_div_pack(VE[1], VO[1]);
          o.z = box_div_pack(VE[2], VO[2]);
          o.w = box_div_pack(VE[3], VO[3]);
          unsigned char* dst = out.base + (long long)oy * out.pitch + x;
          if (vec_store && x + 16 <= rowbytes) {
            *reinterpret_cast<uint4*>(dst) = o;
          } else {
            const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int k = 0; k < 16; k++)
              if (x + k < rowbytes) dst[k] = (unsigned char)(ow[k >> 2] >> ((k & 3) * 8));
          }
        }
      }
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&empty[stage]);
      if (++stage == BX_STAGES) { stage = 0; phase ^= 1; }
    }
  }
}

static int layout_pitch(const vppb_img* i) {
  if (i->align <= 0) return -1;
  long long bs = (long long)i->border * i->elem_bytes;
  if (bs % i->align) bs += i->align - (bs % i->align);
  long long p = (long long)i->ncols * i->elem_bytes + 2 * bs;
  if (p % i->align) p += i->align - (p % i->align);
  return (int)p;
}

// TMA needs: library layout (so the bytes left/right of the domain belong to the row),
// 16-byte aligned pixel (0,0) and pitch, border >= 2 with at least 16 bytes of row before column 0.
static bool tma_eligible(const vppb_img* in) {
  if (in->align < 16 || in->border < 2) return false;
  if (((uintptr_t)in->base % 16) || (in->pitch % 16)) return false;
  if (layout_pitch(in) != in->pitch) return false;
  return true;
}

template <int CS>
static int box5_bytes(const vppb_img* in, const vppb_img* out, void* stream, const char* name) {
  VPPB_REQUIRE(in && out && in->base && out->base, VPPB_E_ARG, "%s: NULL image", name);
  VPPB_REQUIRE(in->elem_bytes == CS && out->elem_bytes == CS, VPPB_E_ARG, "%s: element size must be %d", name, CS);
  VPPB_REQUIRE(same_domain(in, out), VPPB_E_ARG, "%s: domains differ", name);
  VPPB_REQUIRE(in->border >= 2, VPPB_E_BORDER, "%s: input border %d < 2", name, in->border);
  cudaStream_t st = as_stream(stream);
  const int rowbytes = in->ncols * CS;
  if (tma_eligible(in)) {
    CUtensorMap tmap;
    unsigned char* origin = static_cast<unsigned char*>(in->base) - 2LL * in->pitch - 16;
    const uint64_t width_el = ((uint64_t)rowbytes + 32 + 7) / 8;
    int rc = encode_tensor_map_2d(&tmap, origin, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, width_el, (uint64_t)in->nrows + 4,
                                  (uint64_t)in->pitch, BX_BOXW / 8, BX_CH);
    if (rc) return rc;
    const int strips = (rowbytes + BX_TW - 1) / BX_TW;
    const int sms = sm_count();
    int seg_chunks = 2;
    const int cand[5] = {8, 6, 4, 3, 2};
    for (int k = 0; k < 5; k++) {
      int sr = cand[k] * BX_CH - 4;
      long long it = (long long)strips * ((in->nrows + sr - 1) / sr);
      if (it >= (3LL * sms) / 2) { seg_chunks = cand[k]; break; }
    }
    const int seg_rows = seg_chunks * BX_CH - 4;
    const int segs = (in->nrows + seg_rows - 1) / seg_rows;
    const int items = strips * segs;
    const int grid = items < sms * 5 ? items : sms * 5;
    const int vec_store = (((uintptr_t)out->base % 16) == 0 && (out->pitch % 16) == 0) ? 1 : 0;
    const size_t smem = BX_STAGES * BX_STAGE_BYTES + 2 * BX_STAGES * sizeof(uint64_t);
    k_box5_bytes_tma<CS><<<grid, BX_THREADS, smem, st>>>(tmap, view(out), rowbytes, strips, segs, seg_chunks, vec_store);
  } else {
    long long total = (long long)in->nrows * rowbytes;
    long long blocks = (total + 255) / 256;
    long long cap = (long long)sm_count() * 16;
    k_box5_direct<unsigned char, int, CS><<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(view(in), view(out), rowbytes);
  }
  VPPB_LAUNCH_CHECK(name);
  return VPPB_OK;
}

}  // namespace vppb

using namespace vppb;

extern "C" {

int vppb_box5x5_u8c3(const vppb_img* in, const vppb_img* out, void* stream) {
  return box5_bytes<3>(in, out, stream, "vppb_box5x5_u8c3");
}

int vppb_box5x5_u8(const vppb_img* in, const vppb_img* out, void* stream) {
  return box5_bytes<1>(in, out, stream, "vppb_box5x5_u8");
}

int vppb_box5x5_i32(const vppb_img* in, const vppb_img* out, void* stream) {
  VPPB_REQUIRE(in && out && in->base && out->base, VPPB_E_ARG, "vppb_box5x5_i32: NULL image");
  VPPB_REQUIRE(in->elem_bytes == 4 && out->elem_bytes == 4, VPPB_E_ARG, "vppb_box5x5_i32: element size must be 4");
  VPPB_REQUIRE(same_domain(in, out), VPPB_E_ARG, "vppb_box5x5_i32: domains differ");
  VPPB_REQUIRE(in->border >= 2, VPPB_E_BORDER, "vppb_box5x5_i32: input border %d < 2", in->border);
  long long total = (long long)in->nrows * in->ncols;
  long long blocks = (total + 255) / 256;
  long long cap = (long long)sm_count() * 16;
  // int accumulation wraps like the reference's `int sum` on two's-complement hardware
  k_box5_direct<int, int, 1><<<(int)(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>(view(in), view(out), in->ncols);
  VPPB_LAUNCH_CHECK("vppb_box5x5_i32");
  return VPPB_OK;
}

}  // extern "C"
