// TEST INFRASTRUCTURE ONLY: stand-in for vpp_b200/csrc/tma.cuh.  The tensor map is a plain record of what
// cuTensorMapEncodeTiled was given; a TMA load copies the box at once (out-of-bounds elements read as zero) and
// completes its bytes on the mbarrier; mbarriers live in a side table of the emulated block.  The rules the hardware
// enforces are asserted here, so that breaking one fails on the CPU instead of raising "illegal instruction" on the
// GPU: 16-byte aligned global address and pitch, box <= 256 elements per dimension, 128-byte aligned shared
// destination, and the rule measured on B200 (dbg/tma_test.cu): innermost coordinate x element size is a multiple of 16.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

namespace emu {
struct TensorMap2d { unsigned char* origin; uint64_t elem_bytes, width, height, pitch, box_w, box_h, magic; };
void mbar_init(uint64_t* bar, uint32_t count);
void mbar_arrive(uint64_t* bar, uint32_t expect_tx);
void mbar_complete_tx(uint64_t* bar, uint32_t bytes);
bool mbar_phase_done(uint64_t* bar, uint32_t parity);
void yield();
void fail(const char* what);
}  // namespace emu

namespace vppb {

int encode_tensor_map_2d(CUtensorMap* map, void* origin, CUtensorMapDataType elem, int elem_bytes, uint64_t width,
                         uint64_t height, uint64_t pitch, uint32_t box_w, uint32_t box_h);

inline void mbar_init(uint64_t* bar, uint32_t count) { emu::mbar_init(bar, count); }
inline void fence_barrier_init() {}
inline void fence_proxy_async() {}
inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) { emu::mbar_arrive(bar, bytes); }
inline void mbar_arrive(uint64_t* bar) { emu::mbar_arrive(bar, 0); }
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) { return emu::mbar_phase_done(bar, parity); }
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!emu::mbar_phase_done(bar, parity)) emu::yield();
}
inline void tma_load_2d(void* smem_dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  const emu::TensorMap2d* t = reinterpret_cast<const emu::TensorMap2d*>(map);
  if (t->magic != 0x7e4503a9ull) emu::fail("tma_load_2d: tensor map was not encoded");
  if (((uintptr_t)smem_dst % 128) != 0) emu::fail("tma_load_2d: shared destination not 128-byte aligned");
  if ((((long long)x * (long long)t->elem_bytes) % 16) != 0) emu::fail("tma_load_2d: innermost coordinate x element size not a multiple of 16 bytes");
  unsigned char* dst = static_cast<unsigned char*>(smem_dst);
  for (uint64_t r = 0; r < t->box_h; r++)
    for (uint64_t c = 0; c < t->box_w; c++) {
      const long long gy = (long long)y + (long long)r, gx = (long long)x + (long long)c;
      unsigned char* d = dst + (r * t->box_w + c) * t->elem_bytes;
      if (gy < 0 || gx < 0 || gy >= (long long)t->height || gx >= (long long)t->width) memset(d, 0, t->elem_bytes);
      else memcpy(d, t->origin + gy * (long long)t->pitch + gx * (long long)t->elem_bytes, t->elem_bytes);
    }
  emu::mbar_complete_tx(bar, (uint32_t)(t->box_w * t->box_h * t->elem_bytes));
}
inline void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  if (((uintptr_t)smem_dst % 16) != 0 || ((uintptr_t)gsrc % 16) != 0 || (bytes % 16) != 0) emu::fail("tma_load_1d: source, destination and size must be multiples of 16 bytes");
  memcpy(smem_dst, gsrc, bytes);
  emu::mbar_complete_tx(bar, bytes);
}
inline void tma_prefetch_desc(const CUtensorMap*) {}

}  // namespace vppb
