// rgb_to_graylevel for 8-bit RGB / RGBA frames, and the fused frame ingest that every caller of the path performs
// before it: clone(frame, _border = b); fill_border_mirror; rgb_to_graylevel (examples/video_extruder.cc:46-48,
// examples/fast_detector.cc:19-20) - three full-frame passes in the reference, one launch here.
// Reference: vpp/core/colorspace_conversions.hh:10-47: o = (i[0] + i[1] + i[2]) / 3 in int arithmetic (truncating),
// applied to every pixel of in.domain_with_border(); a 4th channel is ignored.
// HBM-bound: 3 (or 4) bytes read + 1 byte written per pixel.
#include "common.cuh"

namespace vppb {

// (a + b + c) / 3, exactly: 43691 / 2^17 = 1/3 + 1/393216, and s / 393216 < 1/3 for every s <= 765 (checked
// exhaustively in tests/test_abi.py); the product stays below 2^26
__host__ __device__ __forceinline__ unsigned gray3(unsigned a, unsigned b, unsigned c) { return ((a + b + c) * 43691u) >> 17; }

__device__ __forceinline__ unsigned byte_at(const uint32_t* w, int i) { return (w[i >> 2] >> (8 * (i & 3))) & 0xFFu; }

// Work items: [0, n_vec) = 16 consecutive output pixels of one row (48 / 64 input bytes as 16-byte loads, one 16-byte
// store); [n_vec, n_vec + n_tail) = single pixels right of the last full group; then, if mb > 0, one item per pixel of
// out's border frame of width mb, computed from the mirrored domain position of `in` (so it waits for nobody).
// frame = how many border rows / columns around the domain the group / tail items cover (0, or the common border in
// the reference's domain_with_border form; the vector path is only used for frame == 0).
template <int CH>
__global__ void __launch_bounds__(256) k_rgb_to_gray(Img in, Img out, int groups_per_row, int frame, int mb, int vec_ok) {
  const int nr = out.nrows + 2 * frame, nc = out.ncols + 2 * frame;
  const long long n_vec = vec_ok ? (long long)nr * groups_per_row : 0;
  const int tail0 = vec_ok ? groups_per_row * 16 : 0;  // first column (frame coordinates) handled pixel by pixel
  const long long n_tail = (long long)nr * (nc - tail0);
  const long long wfull = out.ncols + 2LL * mb, n_top = (long long)mb * wfull, n_side = (long long)out.nrows * mb;
  const long long total = n_vec + n_tail + 2 * n_top + 2 * n_side;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (i < n_vec) {
      const int r = (int)(i / groups_per_row), c0 = (int)(i - (long long)r * groups_per_row) * 16;
      const uint4* src = reinterpret_cast<const uint4*>(row_ptr<unsigned char>(in, r) + (long long)c0 * CH);
      uint32_t w[4 * CH];
#pragma unroll
      for (int k = 0; k < CH; k++) {
        const uint4 v = __ldg(src + k);
        w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
      }
      uint32_t o[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        uint32_t acc = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int p = (4 * q + j) * CH;
          acc |= gray3(byte_at(w, p), byte_at(w, p + 1), byte_at(w, p + 2)) << (8 * j);
        }
        o[q] = acc;
      }
      *reinterpret_cast<uint4*>(row_ptr<unsigned char>(out, r) + c0) = make_uint4(o[0], o[1], o[2], o[3]);
      continue;
    }
    int r, c, sr, sc;
    if (i < n_vec + n_tail) {
      const long long j = i - n_vec;
      const int w = nc - tail0;
      r = (int)(j / w) - frame;
      c = tail0 + (int)(j - (long long)(r + frame) * w) - frame;
      sr = r; sc = c;
    } else {
      long long j = i - n_vec - n_tail;
      if (j < n_top) { r = (int)(j / wfull) - mb; c = (int)(j % wfull) - mb; }
      else if (j < 2 * n_top) { j -= n_top; r = out.nrows + (int)(j / wfull); c = (int)(j % wfull) - mb; }
      else if (j < 2 * n_top + n_side) { j -= 2 * n_top; r = (int)(j / mb); c = (int)(j % mb) - mb; }
      else { j -= 2 * n_top + n_side; r = (int)(j / mb); c = out.ncols + (int)(j % mb); }
      sr = r < 0 ? -r - 1 : (r >= out.nrows ? 2 * out.nrows - r - 1 : r);  // fill.hh:59-82
      sc = c < 0 ? -c - 1 : (c >= out.ncols ? 2 * out.ncols - c - 1 : c);
    }
    const unsigned char* s = row_ptr<unsigned char>(in, sr) + (long long)sc * CH;
    row_ptr<unsigned char>(out, r)[c] = (unsigned char)gray3(s[0], s[1], s[2]);
  }
}

static int rgb_to_gray(const vppb_img* in, const vppb_img* out, int mirror, void* stream, const char* name) {
  VPPB_REQUIRE(in && out && in->base && out->base, VPPB_E_ARG, "%s: NULL image", name);
  VPPB_REQUIRE((in->elem_bytes == 3 || in->elem_bytes == 4) && out->elem_bytes == 1, VPPB_E_ARG, "%s: needs 3- or 4-byte input pixels and a u8 output", name);
  VPPB_REQUIRE(same_domain(in, out), VPPB_E_ARG, "%s: domains differ", name);
  int frame = 0, mb = 0;
  if (mirror) {
    mb = out->border;
    VPPB_REQUIRE(mb <= out->nrows && mb <= out->ncols, VPPB_E_BORDER, "%s: border %d larger than the image", name, mb);
  } else {
    frame = out->border;  // domain_with_border() of the output (colorspace_conversions.hh:26-27: out has the input's border)
    VPPB_REQUIRE(in->border >= frame, VPPB_E_BORDER, "%s: input border %d < output border %d", name, in->border, frame);
  }
  const int ch = in->elem_bytes;
  const int vec_ok = frame == 0 && ((uintptr_t)in->base % 16) == 0 && (in->pitch % 16) == 0 && ((uintptr_t)out->base % 16) == 0 && (out->pitch % 16) == 0;
  const int groups = vec_ok ? out->ncols / 16 : 0;
  const long long nr = out->nrows + 2LL * frame, nc = out->ncols + 2LL * frame;
  const long long items = nr * groups + nr * (nc - 16LL * groups) + 2LL * mb * (out->ncols + 2LL * mb) + 2LL * out->nrows * mb;
  long long blocks = (items + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  const int grid = (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
  if (ch == 3) k_rgb_to_gray<3><<<grid, 256, 0, as_stream(stream)>>>(view(in), view(out), groups, frame, mb, vec_ok);
  else k_rgb_to_gray<4><<<grid, 256, 0, as_stream(stream)>>>(view(in), view(out), groups, frame, mb, vec_ok);
  VPPB_LAUNCH_CHECK(name);
  return VPPB_OK;
}

}  // namespace vppb

using namespace vppb;

extern "C" {

int vppb_rgb_to_graylevel_u8(const vppb_img* in, const vppb_img* out, void* stream) { return rgb_to_gray(in, out, 0, stream, "vppb_rgb_to_graylevel_u8"); }
int vppb_rgb_to_graylevel_u8_mirror(const vppb_img* in, const vppb_img* out, void* stream) {
  return rgb_to_gray(in, out, 1, stream, "vppb_rgb_to_graylevel_u8_mirror");
}

}  // extern "C"
