// Library plumbing: error text, device selection, image2d<V> storage in pitched HBM.
// Layout follows imageNd<V,N>::allocate (reference vpp/core/imageNd.hpp:151-196) with the row
// alignment raised to 128 B (one L2 line / TMA-friendly) instead of the CPU's 16/32 B.
#include <atomic>
#include "common.cuh"

#include <stdarg.h>

namespace vppb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return VPPB_E_CUDA;
}

int sm_count() {
  static std::atomic<int> cached{0};
  int n = cached.load(std::memory_order_relaxed);
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached.store(n, std::memory_order_relaxed);
    else
      return 148;
  }
  return n;
}

}  // namespace vppb

using namespace vppb;

extern "C" {

int vppb_version(void) { return VPPB_VERSION; }

const char* vppb_last_error(void) { return g_err; }

int vppb_device_count(int* n) {
  VPPB_REQUIRE(n, VPPB_E_ARG, "vppb_device_count: NULL");
  VPPB_CUDA(cudaGetDeviceCount(n));
  return VPPB_OK;
}

int vppb_init(int device) {
  VPPB_CUDA(cudaSetDevice(device));
  VPPB_CUDA(cudaFree(0));
  return VPPB_OK;
}

int vppb_sync(void* stream) {
  VPPB_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return VPPB_OK;
}

// imageNd.hpp:161-170: border bytes rounded up to `align`, pitch rounded up to `align`.
static int layout(int nrows, int ncols, int elem, int border, int align, int* pitch, int64_t* total,
                  int64_t* origin) {
  if (nrows <= 0 || ncols <= 0 || elem <= 0 || border < 0 || align <= 0) return VPPB_E_ARG;
  int64_t border_size = (int64_t)border * elem;
  int64_t border_padding = 0;
  if (border_size % align) {
    border_padding = align - (border_size % align);
    border_size += border_padding;
  }
  int64_t p = (int64_t)ncols * elem + border_size * 2;
  if (p % align) p += align - (p % align);
  if (p > 0x7fffffff) return VPPB_E_ARG;
  int64_t rows = (int64_t)nrows + 2 * (int64_t)border;
  if (pitch) *pitch = (int)p;
  if (total) *total = rows * p;
  // imageNd.hpp:193-194: begin_ = data_ + border_padding + offset(border, border)
  if (origin) *origin = border_padding + (int64_t)border * p + (int64_t)border * elem;
  return VPPB_OK;
}

int vppb_layout(int32_t nrows, int32_t ncols, int32_t elem_bytes, int32_t border, int32_t align,
                int32_t* pitch, int64_t* total_bytes, int64_t* origin_offset) {
  int rc = layout(nrows, ncols, elem_bytes, border, align, pitch, total_bytes, origin_offset);
  if (rc) set_error("vppb_layout: invalid geometry %dx%d elem %d border %d align %d", nrows, ncols, elem_bytes,
                    border, align);
  return rc;
}

int vppb_wrap(vppb_img* out, void* device_buffer, int32_t nrows, int32_t ncols, int32_t elem_bytes,
              int32_t border, int32_t align) {
  VPPB_REQUIRE(out && device_buffer, VPPB_E_ARG, "vppb_wrap: NULL argument");
  int pitch; int64_t total, origin;
  int rc = layout(nrows, ncols, elem_bytes, border, align, &pitch, &total, &origin);
  VPPB_REQUIRE(rc == 0, rc, "vppb_wrap: invalid geometry");
  VPPB_REQUIRE(((uintptr_t)device_buffer % align) == 0, VPPB_E_ARG, "vppb_wrap: buffer not %d-byte aligned", align);
  out->base = static_cast<unsigned char*>(device_buffer) + origin;
  out->alloc = nullptr;
  out->nrows = nrows; out->ncols = ncols; out->pitch = pitch; out->border = border;
  out->elem_bytes = elem_bytes; out->align = align;
  return VPPB_OK;
}

int vppb_alloc(vppb_img* out, int32_t nrows, int32_t ncols, int32_t elem_bytes, int32_t border, int32_t align) {
  VPPB_REQUIRE(out, VPPB_E_ARG, "vppb_alloc: NULL");
  int pitch; int64_t total, origin;
  int rc = layout(nrows, ncols, elem_bytes, border, align, &pitch, &total, &origin);
  VPPB_REQUIRE(rc == 0, rc, "vppb_alloc: invalid geometry %dx%d elem %d border %d align %d", nrows, ncols,
               elem_bytes, border, align);
  void* p = nullptr;
  // cudaMalloc returns >= 256-byte aligned memory; larger alignments get the reference's slack scheme.
  int64_t slack = align > 256 ? align : 0;
  VPPB_CUDA(cudaMalloc(&p, (size_t)(total + slack)));
  unsigned char* data = static_cast<unsigned char*>(p);
  if ((uintptr_t)data % align) data += align - ((uintptr_t)data % align);
  out->base = data + origin;
  out->alloc = p;
  out->nrows = nrows; out->ncols = ncols; out->pitch = pitch; out->border = border;
  out->elem_bytes = elem_bytes; out->align = align;
  return VPPB_OK;
}

int vppb_free(vppb_img* img) {
  VPPB_REQUIRE(img, VPPB_E_ARG, "vppb_free: NULL");
  if (img->alloc) VPPB_CUDA(cudaFree(img->alloc));
  img->alloc = nullptr;
  img->base = nullptr;
  return VPPB_OK;
}

int vppb_subimage(const vppb_img* img, int32_t r0, int32_t c0, int32_t r1, int32_t c1, vppb_img* out) {
  VPPB_REQUIRE(img && out && img->base, VPPB_E_ARG, "vppb_subimage: NULL");
  VPPB_REQUIRE(r0 <= r1 && c0 <= c1 && r0 >= -img->border && c0 >= -img->border &&
                   r1 < img->nrows + img->border && c1 < img->ncols + img->border,
               VPPB_E_ARG, "vppb_subimage: box (%d,%d)-(%d,%d) outside the buffer", r0, c0, r1, c1);
  *out = *img;
  out->alloc = nullptr;  // views never own
  out->base = static_cast<unsigned char*>(img->base) + (int64_t)r0 * img->pitch + (int64_t)c0 * img->elem_bytes;
  out->nrows = r1 - r0 + 1;
  out->ncols = c1 - c0 + 1;
  // the reference keeps border_ unchanged for subimages (imageNd.hpp:327-339 copies the descriptor);
  // the addressable frame of a view is whatever the parent buffer provides.
  return VPPB_OK;
}

static int xfer(const vppb_img* im, void* host, int64_t host_pitch, int with_border, void* stream, bool up) {
  VPPB_REQUIRE(im && im->base && host, VPPB_E_ARG, "vppb_%s: NULL argument", up ? "upload" : "download");
  int b = with_border ? im->border : 0;
  int64_t width = (int64_t)(im->ncols + 2 * b) * im->elem_bytes;
  int64_t height = im->nrows + 2 * b;
  VPPB_REQUIRE(host_pitch >= width, VPPB_E_ARG, "host pitch %lld < row bytes %lld", (long long)host_pitch,
               (long long)width);
  unsigned char* d = static_cast<unsigned char*>(im->base) - (int64_t)b * im->pitch - (int64_t)b * im->elem_bytes;
  unsigned char* h = static_cast<unsigned char*>(host) - (int64_t)b * host_pitch - (int64_t)b * im->elem_bytes;
  if (host_pitch == width && im->pitch == width) {  // gap-free on both sides: one linear copy instead of a pitched one
    if (up) VPPB_CUDA(cudaMemcpyAsync(d, h, (size_t)(width * height), cudaMemcpyHostToDevice, as_stream(stream)));
    else VPPB_CUDA(cudaMemcpyAsync(h, d, (size_t)(width * height), cudaMemcpyDeviceToHost, as_stream(stream)));
    return VPPB_OK;
  }
  if (up)
    VPPB_CUDA(cudaMemcpy2DAsync(d, im->pitch, h, host_pitch, width, height, cudaMemcpyHostToDevice, as_stream(stream)));
  else
    VPPB_CUDA(cudaMemcpy2DAsync(h, host_pitch, d, im->pitch, width, height, cudaMemcpyDeviceToHost, as_stream(stream)));
  return VPPB_OK;
}

int vppb_upload(const vppb_img* dst, const void* host, int64_t host_pitch, int with_border, void* stream) {
  return xfer(dst, const_cast<void*>(host), host_pitch, with_border, stream, true);
}

int vppb_download(const vppb_img* src, void* host, int64_t host_pitch, int with_border, void* stream) {
  return xfer(src, host, host_pitch, with_border, stream, false);
}

}  // extern "C"
