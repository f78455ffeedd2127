// pixel_wise named kernels: a = b + c, fill, copy, border fills, sum.
// Reference semantics: vpp/core/pixel_wise.hpp:69-165 (the row-parallel map), fill.hh:12-121,
// copy.hh:10-27, sum.hh:12-19.  Every kernel is a streaming map: 16-byte vector accesses on the
// 128-byte aligned rows, grid sized in multiples of the SM count, no shared memory (each byte is
// touched once, so staging would only add latency).
#include "common.cuh"

namespace vppb {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

// grid for a streaming kernel over `items` work items where each thread handles kUnroll of them
static int stream_grid(long long items, int ctas_per_sm = 8) {
  long long per_cta = (long long)kThreads * kUnroll;
  long long need = (items + per_cta - 1) / per_cta;
  long long cap = (long long)sm_count() * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ---------------------------------------------------------------- a = b + c (int32)
// rows x nvec 16-byte vectors; when all three images are gap-free (pitch == row bytes) the
// host collapses the image to one long row.
__global__ void __launch_bounds__(kThreads) k_add_i32_vec(Img a, Img b, Img c, int nvec, long long total) {
  const long long stride = (long long)gridDim.x * kThreads * kUnroll;
  for (long long i0 = (long long)blockIdx.x * kThreads * kUnroll + threadIdx.x; i0 < total; i0 += stride) {
    int4 vb[kUnroll], vc[kUnroll];
    long long oa[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; u++) {
      long long i = i0 + (long long)u * kThreads;
      if (i < total) {
        long long r = i / nvec;
        int v = (int)(i - r * nvec);
        oa[u] = r * a.pitch + (long long)v * 16;
        vb[u] = ld_stream(reinterpret_cast<const int4*>(b.base + r * b.pitch + (long long)v * 16));
        vc[u] = ld_stream(reinterpret_cast<const int4*>(c.base + r * c.pitch + (long long)v * 16));
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; u++) {
      long long i = i0 + (long long)u * kThreads;
      if (i < total) {
        int4 s;
        s.x = vb[u].x + vc[u].x; s.y = vb[u].y + vc[u].y; s.z = vb[u].z + vc[u].z; s.w = vb[u].w + vc[u].w;
        st_stream(reinterpret_cast<int4*>(a.base + oa[u]), s);
      }
    }
  }
}

// scalar path: columns [c0, ncols) of every row (vector tail, or everything when unaligned)
__global__ void k_add_i32_scalar(Img a, Img b, Img c, int c0) {
  int w = a.ncols - c0;
  long long total = (long long)a.nrows * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int r = (int)(i / w);
    int col = c0 + (int)(i - (long long)r * w);
    row_ptr<int>(a, r)[col] = row_ptr<int>(b, r)[col] + row_ptr<int>(c, r)[col];
  }
}

static bool aligned16(const vppb_img* i) { return ((uintptr_t)i->base % 16) == 0 && (i->pitch % 16) == 0; }

// ---------------------------------------------------------------- fill
struct FillVal { unsigned char b[64]; };

// power-of-two element sizes <= 16: one 16-byte pattern for every aligned vector of a row range.
// Range = bytes [x0, x0+wbytes) relative to row pointer `base + r*pitch`, rows [r0, r0+rows).
__global__ void __launch_bounds__(kThreads) k_fill_vec(unsigned char* base, long long pitch, int rows, int nvec, int4 pat) {
  const long long total = (long long)rows * nvec;
  const long long stride = (long long)gridDim.x * kThreads * kUnroll;
  for (long long i0 = (long long)blockIdx.x * kThreads * kUnroll + threadIdx.x; i0 < total; i0 += stride) {
#pragma unroll
    for (int u = 0; u < kUnroll; u++) {  // kUnroll independent 16-byte stores in flight per thread
      const long long i = i0 + (long long)u * kThreads;
      if (i < total) {
        const long long r = i / nvec;
        const int v = (int)(i - r * nvec);
        st_stream(reinterpret_cast<int4*>(base + r * pitch + (long long)v * 16), pat);
      }
    }
  }
}

// generic element-wise fill of a rows x cols element rectangle starting at `base`
__global__ void k_fill_elem(unsigned char* base, long long pitch, int rows, int cols, int elem, FillVal val) {
  long long total = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / cols;
    int c = (int)(i - r * cols);
    unsigned char* p = base + r * pitch + (long long)c * elem;
    for (int k = 0; k < elem; k++) p[k] = val.b[k];
  }
}

static int fill_rect(unsigned char* base, long long pitch, int rows, int cols, int elem, const void* value, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return VPPB_OK;
  FillVal val;
  memset(&val, 0, sizeof(val));
  memcpy(val.b, value, elem);
  bool pow2 = (elem == 1 || elem == 2 || elem == 4 || elem == 8 || elem == 16);
  long long wbytes = (long long)cols * elem;
  if (pow2 && (pitch % 16) == 0 && wbytes >= 64) {
    // head (to the next 16-byte boundary), vector body, tail
    uintptr_t addr = (uintptr_t)base;
    int head_b = (int)((16 - (addr % 16)) % 16);
    // head must end on an element boundary: elem divides 16 and base is element-aligned for
    // every layout the library produces; otherwise take the generic path.
    if (addr % elem == 0) {
      int head_e = head_b / elem;
      long long body_b = ((wbytes - head_b) / 16) * 16;
      int nvec = (int)(body_b / 16);
      int tail_e = (int)((wbytes - head_b - body_b) / elem);
      unsigned char patb[16];
      for (int k = 0; k < 16; k++) patb[k] = val.b[k % elem];
      int4 pat;
      memcpy(&pat, patb, 16);
      if (head_e > 0) {
        k_fill_elem<<<stream_grid((long long)rows * head_e), kThreads, 0, st>>>(base, pitch, rows, head_e, elem, val);
      }
      if (nvec > 0) {
        long long total = (long long)rows * nvec;
        if (pitch == body_b && head_b == 0 && total < 0x7fffffffLL) { nvec = (int)total; rows = 1; }  // gap-free: one long row
        k_fill_vec<<<stream_grid(total), kThreads, 0, st>>>(base + head_b, pitch, rows, nvec, pat);
      }
      if (tail_e > 0) {
        k_fill_elem<<<stream_grid((long long)rows * tail_e), kThreads, 0, st>>>(base + head_b + body_b, pitch, rows, tail_e, elem, val);
      }
      VPPB_LAUNCH_CHECK("fill");
      return VPPB_OK;
    }
  }
  k_fill_elem<<<stream_grid((long long)rows * cols), kThreads, 0, st>>>(base, pitch, rows, cols, elem, val);
  VPPB_LAUNCH_CHECK("fill");
  return VPPB_OK;
}

// ---------------------------------------------------------------- copy
__global__ void __launch_bounds__(kThreads) k_copy_vec(const unsigned char* src, long long spitch, unsigned char* dst, long long dpitch,
                                                        int nvec, long long total) {
  const long long stride = (long long)gridDim.x * kThreads * kUnroll;
  for (long long i0 = (long long)blockIdx.x * kThreads * kUnroll + threadIdx.x; i0 < total; i0 += stride) {
    int4 v[kUnroll];
    long long od[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; u++) {
      long long i = i0 + (long long)u * kThreads;
      if (i < total) {
        long long r = i / nvec;
        int k = (int)(i - r * nvec);
        od[u] = r * dpitch + (long long)k * 16;
        v[u] = ld_stream(reinterpret_cast<const int4*>(src + r * spitch + (long long)k * 16));
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; u++)
      if (i0 + (long long)u * kThreads < total) st_stream(reinterpret_cast<int4*>(dst + od[u]), v[u]);
  }
}

__global__ void k_copy_bytes(const unsigned char* src, long long spitch, unsigned char* dst, long long dpitch, int rows, int wbytes) {
  long long total = (long long)rows * wbytes;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / wbytes;
    int k = (int)(i - r * wbytes);
    dst[r * dpitch + k] = src[r * spitch + k];
  }
}

static int copy_rect(const unsigned char* src, long long spitch, unsigned char* dst, long long dpitch, int rows, long long wbytes,
                     cudaStream_t st) {
  if (rows <= 0 || wbytes <= 0) return VPPB_OK;
  bool al = ((uintptr_t)src % 16) == ((uintptr_t)dst % 16) && (spitch % 16) == 0 && (dpitch % 16) == 0;
  long long head = 0, body = 0;
  if (al && wbytes >= 64) {
    head = (16 - ((uintptr_t)src % 16)) % 16;
    body = ((wbytes - head) / 16) * 16;
  }
  if (head > 0) k_copy_bytes<<<stream_grid((long long)rows * head), kThreads, 0, st>>>(src, spitch, dst, dpitch, rows, (int)head);
  if (body > 0) {
    int nvec = (int)(body / 16);
    long long total = (long long)rows * nvec;
    const unsigned char* s = src + head;
    unsigned char* d = dst + head;
    if (spitch == wbytes && dpitch == wbytes && head == 0 && body == wbytes && total < 0x7fffffffLL)
      nvec = (int)total;  // gap-free: one long row
    k_copy_vec<<<stream_grid(total), kThreads, 0, st>>>(s, spitch, d, dpitch, nvec, total);
  }
  long long tail = wbytes - head - body;
  if (tail > 0)
    k_copy_bytes<<<stream_grid((long long)rows * tail), kThreads, 0, st>>>(src + head + body, spitch, dst + head + body, dpitch, rows, (int)tail);
  VPPB_LAUNCH_CHECK("copy");
  return VPPB_OK;
}

// ---------------------------------------------------------------- border fills (fill.hh:32-121)
// One thread per border pixel of the frame; mode 0 = value, 1 = mirror, 2 = closest.
__global__ void k_fill_border(Img im, int elem, int mode, FillVal val) {
  const int b = im.border, nr = im.nrows, nc = im.ncols;
  const long long wfull = nc + 2LL * b;
  const long long n_top = (long long)b * wfull;      // rows [-b,-1]
  const long long n_side = (long long)nr * b;        // rows [0,nr) x cols [-b,-1] (and the right strip)
  const long long total = 2 * n_top + 2 * n_side;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int r, c;
    if (i < n_top) { r = (int)(i / wfull) - b; c = (int)(i % wfull) - b; }
    else if (i < 2 * n_top) { long long j = i - n_top; r = nr + (int)(j / wfull); c = (int)(j % wfull) - b; }
    else if (i < 2 * n_top + n_side) { long long j = i - 2 * n_top; r = (int)(j / b); c = (int)(j % b) - b; }
    else { long long j = i - 2 * n_top - n_side; r = (int)(j / b); c = nc + (int)(j % b); }
    unsigned char* dst = im.base + (long long)r * im.pitch + (long long)c * elem;
    if (mode == 0) {
      for (int k = 0; k < elem; k++) dst[k] = val.b[k];
    } else {
      int sr, sc;
      if (mode == 1) {  // fill.hh:59-82: img(-1-k) = img(k), img(n+k) = img(n-1-k)
        sr = r < 0 ? -r - 1 : (r >= nr ? 2 * nr - r - 1 : r);
        sc = c < 0 ? -c - 1 : (c >= nc ? 2 * nc - c - 1 : c);
      } else {          // fill.hh:93-120: clamp
        sr = r < 0 ? 0 : (r >= nr ? nr - 1 : r);
        sc = c < 0 ? 0 : (c >= nc ? nc - 1 : c);
      }
      const unsigned char* src = im.base + (long long)sr * im.pitch + (long long)sc * elem;
      for (int k = 0; k < elem; k++) dst[k] = src[k];
    }
  }
}

static int fill_border(const vppb_img* img, int mode, const void* value, void* stream) {
  VPPB_REQUIRE(img && img->base, VPPB_E_ARG, "fill_border: NULL image");
  VPPB_REQUIRE(img->elem_bytes > 0 && img->elem_bytes <= 64, VPPB_E_ARG, "fill_border: element size %d unsupported", img->elem_bytes);
  if (img->border == 0) return VPPB_OK;
  VPPB_REQUIRE(mode == 0 || (img->border <= img->nrows && img->border <= img->ncols), VPPB_E_BORDER,
               "fill_border: border %d larger than the image", img->border);
  FillVal val;
  memset(&val, 0, sizeof(val));
  if (value) memcpy(val.b, value, img->elem_bytes);
  long long total = 2LL * img->border * (img->ncols + 2LL * img->border) + 2LL * img->nrows * img->border;
  k_fill_border<<<stream_grid(total), kThreads, 0, as_stream(stream)>>>(view(img), img->elem_bytes, mode, val);
  VPPB_LAUNCH_CHECK("fill_border");
  return VPPB_OK;
}

// ---------------------------------------------------------------- sum (sum.hh:12-19)
template <typename T>
__global__ void k_sum(Img im, int* out) {
  int acc = 0;  // plus_promotion<char/uchar/int> == int; unsigned wrap == int wrap bit-wise
  long long total = (long long)im.nrows * im.ncols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int r = (int)(i / im.ncols);
    int c = (int)(i - (long long)r * im.ncols);
    acc += (int)row_ptr<T>(im, r)[c];
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc != 0) atomicAdd(out, acc);
}

static int* sum_scratch() {
  static thread_local int* p = nullptr;
  static thread_local int dev_of = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!p || dev_of != dev) {
    if (cudaMalloc(&p, 256) != cudaSuccess) return nullptr;
    dev_of = dev;
  }
  return p;
}

// ---------------------------------------------------------------- batched halo rows <-> staging
// One launch moves the `halo` edge (or border) rows of up to 64 tiles; bytes are moved individually
// (the row block starts border*elem bytes left of column 0, so no common vector alignment exists).
constexpr int kHaloBatch = 64;
struct HaloBatch {
  unsigned char* img_rows[kHaloBatch];  // address of the first byte of the first row block of image i
  int n, pitch, rows, wbytes;
};
__global__ void k_halo_batch(HaloBatch b, unsigned char* staging, int to_staging) {
  const long long per_img = (long long)b.rows * b.wbytes;
  const long long total = per_img * b.n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int img = (int)(i / per_img);
    const long long j = i - (long long)img * per_img;
    const int r = (int)(j / b.wbytes), x = (int)(j - (long long)r * b.wbytes);
    unsigned char* p = b.img_rows[img] + (long long)r * b.pitch + x;
    if (to_staging) staging[i] = *p; else *p = staging[i];
  }
}

}  // namespace vppb

using namespace vppb;

extern "C" {

int vppb_pw_add_i32(const vppb_img* a, const vppb_img* b, const vppb_img* c, void* stream) {
  VPPB_REQUIRE(a && b && c && a->base && b->base && c->base, VPPB_E_ARG, "vppb_pw_add_i32: NULL image");
  VPPB_REQUIRE(a->elem_bytes == 4 && b->elem_bytes == 4 && c->elem_bytes == 4, VPPB_E_ARG, "vppb_pw_add_i32: elem_bytes must be 4");
  // pixel_wise iterates the domain of its FIRST argument (pixel_wise.hpp:149-150); the others must cover it.
  VPPB_REQUIRE(b->nrows >= a->nrows && b->ncols >= a->ncols && c->nrows >= a->nrows && c->ncols >= a->ncols, VPPB_E_ARG,
               "vppb_pw_add_i32: operand domains smaller than the first argument's");
  cudaStream_t st = as_stream(stream);
  Img va = view(a), vb = view(b), vc = view(c);
  if (aligned16(a) && aligned16(b) && aligned16(c)) {
    int nvec = a->ncols / 4;
    int rows = a->nrows;
    long long rowb = (long long)a->ncols * 4;
    if ((a->ncols % 4) == 0 && a->pitch == rowb && b->pitch == rowb && c->pitch == rowb && (long long)nvec * rows < 0x7fffffffLL) {
      nvec *= rows;  // gap-free images: one long row
      rows = 1;
    }
    long long total = (long long)rows * nvec;
    if (total > 0) k_add_i32_vec<<<stream_grid(total), kThreads, 0, st>>>(va, vb, vc, nvec, total);
    if (a->ncols % 4) k_add_i32_scalar<<<stream_grid((long long)a->nrows * (a->ncols % 4)), kThreads, 0, st>>>(va, vb, vc, a->ncols - a->ncols % 4);
  } else {
    k_add_i32_scalar<<<stream_grid((long long)a->nrows * a->ncols), kThreads, 0, st>>>(va, vb, vc, 0);
  }
  VPPB_LAUNCH_CHECK("vppb_pw_add_i32");
  return VPPB_OK;
}

int vppb_fill(const vppb_img* img, const void* value, int with_border, void* stream) {
  VPPB_REQUIRE(img && img->base && value, VPPB_E_ARG, "vppb_fill: NULL argument");
  VPPB_REQUIRE(img->elem_bytes > 0 && img->elem_bytes <= 64, VPPB_E_ARG, "vppb_fill: element size %d unsupported", img->elem_bytes);
  int b = with_border ? img->border : 0;
  unsigned char* start = static_cast<unsigned char*>(img->base) - (long long)b * img->pitch - (long long)b * img->elem_bytes;
  return fill_rect(start, img->pitch, img->nrows + 2 * b, img->ncols + 2 * b, img->elem_bytes, value, as_stream(stream));
}

int vppb_copy2d(const vppb_img* src, const vppb_img* dst, int with_border, void* stream) {
  VPPB_REQUIRE(src && dst && src->base && dst->base, VPPB_E_ARG, "vppb_copy2d: NULL image");
  VPPB_REQUIRE(src->elem_bytes == dst->elem_bytes, VPPB_E_ARG, "vppb_copy2d: element sizes differ");
  int b = 0;
  if (with_border) {  // copy.hh:22-27
    VPPB_REQUIRE(same_domain(src, dst), VPPB_E_ARG, "vppb_copy2d: copy_with_border needs equal domains");
    VPPB_REQUIRE(src->border <= dst->border, VPPB_E_BORDER, "vppb_copy2d: src border %d > dst border %d", src->border, dst->border);
    b = src->border;
  } else {            // copy.hh:10-13: domain of the first argument
    VPPB_REQUIRE(dst->nrows >= src->nrows && dst->ncols >= src->ncols, VPPB_E_ARG, "vppb_copy2d: dst smaller than src");
  }
  int e = src->elem_bytes;
  const unsigned char* s = static_cast<const unsigned char*>(src->base) - (long long)b * src->pitch - (long long)b * e;
  unsigned char* d = static_cast<unsigned char*>(dst->base) - (long long)b * dst->pitch - (long long)b * e;
  return copy_rect(s, src->pitch, d, dst->pitch, src->nrows + 2 * b, (long long)(src->ncols + 2 * b) * e, as_stream(stream));
}

// copy(src, dst) + fill_border_mirror(dst) in one launch: the border pixels of dst are read straight from the
// mirrored position in src (the same value dst's domain receives), so nothing depends on the copy having landed.
// Work items: rows*nvec 16-byte vectors, then rows*tail single bytes, then one item per border pixel of dst.
__global__ void __launch_bounds__(kThreads) k_copy_mirror(Img src, Img dst, int nvec, int tail, int elem) {
  const long long total = copy_mirror_items(dst, nvec, tail);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) copy_mirror_item(src, dst, nvec, tail, elem, i);
}

int vppb_copy2d_mirror(const vppb_img* src, const vppb_img* dst, void* stream) {
  VPPB_REQUIRE(src && dst && src->base && dst->base, VPPB_E_ARG, "vppb_copy2d_mirror: NULL image");
  VPPB_REQUIRE(src->elem_bytes == dst->elem_bytes && src->elem_bytes <= 64, VPPB_E_ARG, "vppb_copy2d_mirror: element sizes differ");
  VPPB_REQUIRE(same_domain(src, dst), VPPB_E_ARG, "vppb_copy2d_mirror: needs equal domains");
  VPPB_REQUIRE(dst->border <= dst->nrows && dst->border <= dst->ncols, VPPB_E_BORDER, "vppb_copy2d_mirror: border %d larger than the image",
               dst->border);
  const long long wbytes = (long long)src->ncols * src->elem_bytes;
  const bool al = ((uintptr_t)src->base % 16) == 0 && ((uintptr_t)dst->base % 16) == 0 && (src->pitch % 16) == 0 && (dst->pitch % 16) == 0;
  if (!al || src->base == dst->base) {  // views / odd alignments: the two separate steps
    if (src->base != dst->base) {
      int rc = vppb_copy2d(src, dst, 0, stream);
      if (rc != VPPB_OK) return rc;
    }
    return fill_border(dst, 1, nullptr, stream);
  }
  const int nvec = (int)(wbytes / 16), tail = (int)(wbytes - (long long)nvec * 16);
  const long long b = dst->border;
  const long long total = (long long)src->nrows * (nvec + tail) + 2 * b * (dst->ncols + 2 * b) + 2 * b * dst->nrows;
  if (total <= 0) return VPPB_OK;
  k_copy_mirror<<<stream_grid(total), kThreads, 0, as_stream(stream)>>>(view(src), view(dst), nvec, tail, src->elem_bytes);
  VPPB_LAUNCH_CHECK("vppb_copy2d_mirror");
  return VPPB_OK;
}

int vppb_fill_border_value(const vppb_img* img, const void* value, void* stream) {
  VPPB_REQUIRE(value, VPPB_E_ARG, "vppb_fill_border_value: NULL value");
  return fill_border(img, 0, value, stream);
}
int vppb_fill_border_mirror(const vppb_img* img, void* stream) { return fill_border(img, 1, nullptr, stream); }
int vppb_fill_border_closest(const vppb_img* img, void* stream) { return fill_border(img, 2, nullptr, stream); }

int vppb_sum_i32(const vppb_img* img, int is_signed, int64_t* out_host, void* stream) {
  VPPB_REQUIRE(img && img->base && out_host, VPPB_E_ARG, "vppb_sum_i32: NULL argument");
  VPPB_REQUIRE(img->elem_bytes == 1 || img->elem_bytes == 4, VPPB_E_ARG, "vppb_sum_i32: only 1- or 4-byte scalars");
  int* scratch = sum_scratch();
  VPPB_REQUIRE(scratch, VPPB_E_CUDA, "vppb_sum_i32: scratch allocation failed");
  cudaStream_t st = as_stream(stream);
  VPPB_CUDA(cudaMemsetAsync(scratch, 0, sizeof(int), st));
  int grid = stream_grid((long long)img->nrows * img->ncols);
  if (img->elem_bytes == 4) k_sum<int><<<grid, kThreads, 0, st>>>(view(img), scratch);
  else if (is_signed) k_sum<signed char><<<grid, kThreads, 0, st>>>(view(img), scratch);
  else k_sum<unsigned char><<<grid, kThreads, 0, st>>>(view(img), scratch);
  VPPB_LAUNCH_CHECK("vppb_sum_i32");
  int h = 0;
  VPPB_CUDA(cudaMemcpyAsync(&h, scratch, sizeof(int), cudaMemcpyDeviceToHost, st));
  VPPB_CUDA(cudaStreamSynchronize(st));
  *out_host = (int64_t)h;
  return VPPB_OK;
}

// ---- multi-GPU row tiles: the border rows of a tile ARE its halo; these two move the edge rows
// (full frame width: ncols + 2*border elements) between the image and a contiguous staging buffer
// that the one grouped NCCL neighbour exchange per frame sends / receives.
int64_t vppb_halo_bytes(const vppb_img* img, int32_t halo) {
  if (!img || halo <= 0) return 0;
  return (int64_t)halo * (img->ncols + 2LL * img->border) * img->elem_bytes;
}

int vppb_halo_pack(const vppb_img* img, int32_t halo, int which, void* staging, void* stream) {
  VPPB_REQUIRE(img && img->base && staging, VPPB_E_ARG, "vppb_halo_pack: NULL argument");
  VPPB_REQUIRE(halo > 0 && halo <= img->nrows, VPPB_E_ARG, "vppb_halo_pack: halo %d out of range", halo);
  const long long wbytes = (img->ncols + 2LL * img->border) * img->elem_bytes;
  const int r0 = which == 0 ? 0 : img->nrows - halo;
  const unsigned char* src = static_cast<const unsigned char*>(img->base) + (long long)r0 * img->pitch - (long long)img->border * img->elem_bytes;
  return copy_rect(src, img->pitch, static_cast<unsigned char*>(staging), wbytes, halo, wbytes, as_stream(stream));
}

int vppb_halo_unpack(const vppb_img* img, int32_t halo, int which, const void* staging, void* stream) {
  VPPB_REQUIRE(img && img->base && staging, VPPB_E_ARG, "vppb_halo_unpack: NULL argument");
  VPPB_REQUIRE(halo > 0 && halo <= img->border, VPPB_E_BORDER, "vppb_halo_unpack: halo %d exceeds the border %d", halo, img->border);
  const long long wbytes = (img->ncols + 2LL * img->border) * img->elem_bytes;
  const int r0 = which == 0 ? -halo : img->nrows;
  unsigned char* dst = static_cast<unsigned char*>(img->base) + (long long)r0 * img->pitch - (long long)img->border * img->elem_bytes;
  return copy_rect(static_cast<const unsigned char*>(staging), wbytes, dst, img->pitch, halo, wbytes, as_stream(stream));
}

// Batched forms: the same rows of n tiles of identical geometry, packed back to back (tile i at i * vppb_halo_bytes).
static int halo_batch(const vppb_img* imgs, int n, int halo, int which, void* staging, void* stream, bool pack) {
  VPPB_REQUIRE(imgs && staging && n > 0, VPPB_E_ARG, "vppb_halo_%s_batch: NULL argument", pack ? "pack" : "unpack");
  const vppb_img& g = imgs[0];
  VPPB_REQUIRE(halo > 0 && (pack ? halo <= g.nrows : halo <= g.border), pack ? VPPB_E_ARG : VPPB_E_BORDER, "vppb_halo_batch: halo %d out of range", halo);
  const int wbytes = (int)((g.ncols + 2LL * g.border) * g.elem_bytes);
  const long long per = (long long)halo * wbytes;
  for (int i0 = 0; i0 < n; i0 += kHaloBatch) {
    HaloBatch b;
    b.n = n - i0 < kHaloBatch ? n - i0 : kHaloBatch;
    b.pitch = g.pitch; b.rows = halo; b.wbytes = wbytes;
    for (int k = 0; k < b.n; k++) {
      const vppb_img& im = imgs[i0 + k];
      VPPB_REQUIRE(im.base && im.nrows == g.nrows && im.ncols == g.ncols && im.pitch == g.pitch && im.border == g.border && im.elem_bytes == g.elem_bytes,
                   VPPB_E_ARG, "vppb_halo_batch: tile %d has a different geometry", i0 + k);
      const int r0 = pack ? (which == 0 ? 0 : im.nrows - halo) : (which == 0 ? -halo : im.nrows);
      b.img_rows[k] = static_cast<unsigned char*>(im.base) + (long long)r0 * im.pitch - (long long)im.border * im.elem_bytes;
    }
    const long long total = per * b.n;
    k_halo_batch<<<stream_grid(total / 4 + 1), kThreads, 0, as_stream(stream)>>>(b, static_cast<unsigned char*>(staging) + (long long)i0 * per, pack ? 1 : 0);
  }
  VPPB_LAUNCH_CHECK("vppb_halo_batch");
  return VPPB_OK;
}
int vppb_halo_pack_batch(const vppb_img* imgs, int32_t n, int32_t halo, int which, void* staging, void* stream) {
  return halo_batch(imgs, n, halo, which, staging, stream, true);
}
int vppb_halo_unpack_batch(const vppb_img* imgs, int32_t n, int32_t halo, int which, const void* staging, void* stream) {
  return halo_batch(imgs, n, halo, which, const_cast<void*>(staging), stream, false);
}

}  // extern "C"
