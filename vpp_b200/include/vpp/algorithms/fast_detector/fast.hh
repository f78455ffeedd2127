// fast9(A, th, [_local_maxima | _blockwise, _block_size =, _mask =, _scores = &vec])
// (reference: vpp/algorithms/fast_detector/fast.hh:25-28, fast.hpp:931-955).  Keypoints come back in
// raster order; `_ring = 1` (extension) selects the true Bresenham ring instead of the ring fast9()
// actually samples (fast.hpp:367-368).
#pragma once
#include <algorithm>
#include <memory>
#include <vector>
#include <vpp/core/image2d.hh>

namespace vpp {

namespace internals {
struct device_array {
  vppb_img img;
  explicit device_array(size_t bytes) { vppb_check(vppb_alloc(&img, 1, (int)(bytes ? bytes : 1), 1, 0, 256)); }
  ~device_array() { vppb_free(&img); }
  void* ptr() const { return img.base; }
  void to_host(void* dst, size_t bytes) const {
    if (!bytes) return;
    vppb_img v = img; v.ncols = (int)bytes;
    vppb_check(vppb_download(&v, dst, (int64_t)bytes, 0, nullptr));
    vppb_check(vppb_sync(nullptr));
  }
  void from_host(const void* src, size_t bytes) const {
    if (!bytes) return;
    vppb_img v = img; v.ncols = (int)bytes;
    vppb_check(vppb_upload(&v, src, (int64_t)bytes, 0, nullptr));
    vppb_check(vppb_sync(nullptr));
  }
};
}  // namespace internals

template <typename... OPTS>
std::vector<vint2> fast9(const image2d<unsigned char>& A, int th, OPTS... opts_) {
  auto opts = s::D(opts_...);
  if (A.border() < 3) throw std::runtime_error("Image need a border of 3px at least for the FAST detector");  // fast.hpp:937-938
  image2d<unsigned char> mask = opts.get(s::_mask, image2d<unsigned char>());
  std::vector<int>* scores = opts.get(s::_scores, (std::vector<int>*)nullptr);
  const int block_size = opts.get(s::_block_size, 10);
  const int mode = opts.has(s::_local_maxima) ? VPPB_FAST_LOCAL_MAXIMA : (opts.has(s::_blockwise) ? VPPB_FAST_BLOCKWISE : VPPB_FAST_ALL);
  const int ring = opts.get(s::_ring, 0);
  // workspace and output buffers are kept between calls (one set per thread, regrown when an image or a result needs more):
  // a detector that runs every frame allocates nothing; the device work is queued without a host synchronisation
  // (vppb_fast9_u8_async), then the 4-byte count and the keypoints are read back once.
  struct cache_t { std::unique_ptr<internals::device_array> ws, kps, sc, cnt; size_t ws_bytes = 0; int capacity = 0; };
  static thread_local cache_t cache;
  const size_t ws_bytes = (size_t)vppb_fast9_workspace_bytes(A.nrows(), A.ncols(), block_size);
  if (cache.ws_bytes < ws_bytes) { cache.ws.reset(new internals::device_array(ws_bytes)); cache.ws_bytes = ws_bytes; }
  if (!cache.cnt) cache.cnt.reset(new internals::device_array(4));
  int capacity = std::max(std::max(1024, A.nrows() * A.ncols() / 16), cache.capacity), count = 0;
  std::vector<vint2> kps;
  for (;;) {
    if (cache.capacity < capacity) {
      cache.kps.reset(new internals::device_array((size_t)capacity * sizeof(vppb_int2)));
      cache.sc.reset(new internals::device_array((size_t)capacity * 4));
      cache.capacity = capacity;
    }
    vppb_check(vppb_fast9_u8_async(A.device_read(), th, mask.has_data() ? mask.device_read() : nullptr, mode, block_size, ring, cache.ws->ptr(), (int64_t)cache.ws_bytes,
                                   (vppb_int2*)cache.kps->ptr(), scores ? (int32_t*)cache.sc->ptr() : nullptr, cache.capacity, (int32_t*)cache.cnt->ptr(), nullptr));
    cache.cnt->to_host(&count, 4);
    if (count > cache.capacity) { capacity = count; continue; }  // keypoints beyond the capacity were dropped: run again with room for all
    kps.resize(count);
    cache.kps->to_host(kps.data(), (size_t)count * sizeof(vint2));
    if (scores) { scores->resize(count); cache.sc->to_host(scores->data(), (size_t)count * 4); }
    return kps;
  }
}

template <typename KPS>
void fast9_scores(const image2d<unsigned char>& A, int th, const KPS& keypoints, std::vector<int>& scores) {  // fast.hpp:643-652
  const size_t n = keypoints.size();
  scores.resize(n);
  internals::device_array dk(n * sizeof(vint2)), ds(n * 4);
  dk.from_host(keypoints.data(), n * sizeof(vint2));
  vppb_check(vppb_fast9_scores(A.device_read(), th, (const vppb_int2*)dk.ptr(), (int)n, (int32_t*)ds.ptr(), nullptr));
  ds.to_host(scores.data(), n * 4);
}
// local_maxima_filter(A, nbh_size) (fast.hpp:555-575; nbh_size is ignored by the reference as well): in place, every pixel that is not
// strictly greater than its 8 neighbours becomes zero - in the reference's serial raster order (the neighbours above / to the left
// have already been filtered).  unsigned char or int images with a border >= 1.
template <typename V>
void local_maxima_filter(const image2d<V>& A, int /*nbh_size*/ = 3) {
  static_assert(sizeof(V) == 1 || sizeof(V) == 4, "local_maxima_filter: unsigned char or int images");
  const int64_t bytes = vppb_local_maxima_filter_workspace_bytes(A.nrows(), A.ncols(), (int)sizeof(V));
  internals::device_array ws((size_t)bytes);
  vppb_check(vppb_local_maxima_filter(A.device_write(), ws.ptr(), bytes, nullptr));
  vppb_check(vppb_sync(nullptr));  // the workspace is released on return
}

// fast_detector9_blockwise_rank(A, th, block_size, max_point_per_block, mask, scores) (fast.hpp:801-886): (row, col, rank) of up to
// max_point_per_block (<= 16) keypoints per block, blocks in raster order, ranks by decreasing raw score.
inline std::vector<vint3> fast_detector9_blockwise_rank(const image2d<unsigned char>& A, int th, int block_size, int max_point_per_block,
                                                        const image2d<unsigned char>& mask = image2d<unsigned char>(), std::vector<int>* scores = nullptr,
                                                        int ring = 0) {
  if (A.border() < 3) throw std::runtime_error("Image need a border of 3px at least for the FAST detector");
  const int cells = ((A.nrows() + block_size - 1) / block_size) * ((A.ncols() + block_size - 1) / block_size);
  const int capacity = std::max(1, cells * max_point_per_block);
  const int64_t bytes = vppb_fast9_rank_workspace_bytes(A.nrows(), A.ncols(), block_size, max_point_per_block);
  internals::device_array ws((size_t)bytes), kps((size_t)capacity * 12), sc((size_t)capacity * 4);
  int count = 0;
  vppb_check(vppb_fast9_blockwise_rank_u8(A.device_read(), th, mask.has_data() ? mask.device_read() : nullptr, block_size, max_point_per_block, ring, ws.ptr(), bytes,
                                          (int32_t*)kps.ptr(), (int32_t*)sc.ptr(), capacity, &count, nullptr));
  std::vector<vint3> out(count);
  kps.to_host(out.data(), (size_t)count * sizeof(vint3));
  if (scores) { scores->resize(count); sc.to_host(scores->data(), (size_t)count * 4); }
  return out;
}

inline int fast9_score(const image2d<unsigned char>& A, int th, vint2 p) {  // fast.hpp:655-660
  std::vector<vint2> k(1, p); std::vector<int> s;
  fast9_scores(A, th, k, s);
  return s[0];
}

}  // namespace vpp
