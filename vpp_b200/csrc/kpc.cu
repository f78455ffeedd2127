// Device-resident keypoint_container<keypoint<int>> + keypoint trajectories for video_extruder.
// Reference: vpp/core/keypoint_container.hpp:22-200 (add / remove / move / compact / sync_attributes),
// vpp/core/keypoint_trajectory.hh:11-73, vpp/algorithms/video_extruder/video_extruder.hpp:45-133 (the update loop).
//
// The reference keeps the container on the host and walks it serially after every step of video_extruder_update; here
// the arrays (position, velocity, age, trajectory ring) live in HBM and every step of the loop is a kernel, so a frame
// needs no host round trip of keypoints or masks - only the 4-byte entry count crosses PCIe on detection frames (the host
// sizes its launches with it).  All steps reproduce the SERIAL semantics of the reference:
//   flow step    : entry i moves (velocity = new - old, age + 1) or is removed (age = 0)           (:45-56, container :136-167)
//   merge        : entries that share a keypoint_spacing cell: visiting them in index order, the first becomes the cell's
//                  holder; a later entry older than the holder removes it and takes over, a younger one is removed, an
//                  equally old one stays (:59-84).  Closed form per entry i with M = max age of its cell mates before it:
//                  removed  <=>  age_i < M,  or  (no mate before it or age_i > M) and a later mate is older than it.
//   score filter : fast9_score(frame2, th, position) < 3 removes the entry (:87-91)
//   mask         : 1 everywhere incl. the border, 0 in [-s, s)^2 around EVERY entry, dead ones included (:97-109)
//   add/compact  : detections are appended with age 1, dead entries leave, order kept; kept entries keep their trajectory,
//                  new ones start one at this frame (:111-118, keypoint_container.hpp:22-110)
//   trajectories : alive entries push their position and drop the oldest beyond max_trajectory_length, dead ones die (:122-133)
#include "common.cuh"

#include <limits.h>

#include <algorithm>

namespace vppb {

struct Kpc {
  int capacity, max_traj, n;  // n: entries (dead ones included), host copy
  // double-buffered entry arrays (compaction gathers from one set into the other)
  vppb_int2* pos[2];
  vppb_int2* vel[2];
  int* age[2];
  int* tstart[2];
  int* tlen[2];
  int* thead[2];               // ring position of the newest history entry
  unsigned char* talive[2];
  vppb_float2* thist[2];       // capacity x max_traj
  int cur;
  int* map;                    // compaction: new index of an entry or -1
  int* count;                  // device: entry count after the last add / compact
  int* cell_head;              // merge grid: head of the cell's list, next[] links
  int* next;
  unsigned char* removed;
  long long cells_cap;
};

__device__ __forceinline__ int c_div(int a, int b) { return a / b; }  // C++ integer division (truncation)

__global__ void k_kpc_flow(vppb_int2* pos, vppb_int2* vel, int* age, int n, const vppb_int2* newpos, const unsigned char* valid, int nrows, int ncols) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (!valid[i]) continue;
    const vppb_int2 p = newpos[i];
    if (p.r >= 0 && p.r < nrows && p.c >= 0 && p.c < ncols) {
      vel[i] = vppb_int2{p.r - pos[i].r, p.c - pos[i].c};
      pos[i] = p;
      age[i] += 1;
    } else {
      age[i] = 0;
    }
  }
}

__global__ void k_kpc_fill_int(int* a, long long n, int v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a[i] = v;
}

__global__ void k_kpc_link(const vppb_int2* pos, int n, int spacing, int stride, int* cell_head, int* next) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int cell = (c_div(pos[i].r, spacing) + 1) * stride + c_div(pos[i].c, spacing) + 1;
    next[i] = atomicExch(&cell_head[cell], i);
  }
}

__global__ void k_kpc_merge(const vppb_int2* pos, const int* age, int n, int spacing, int stride, const int* cell_head, const int* next, unsigned char* removed) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int cell = (c_div(pos[i].r, spacing) + 1) * stride + c_div(pos[i].c, spacing) + 1;
    const int a = age[i];
    int before_max = INT_MIN;
    bool any_before = false, later_older = false;
    for (int j = cell_head[cell]; j >= 0; j = next[j]) {
      if (j < i) { any_before = true; before_max = max(before_max, age[j]); }
      else if (j > i && age[j] > a) later_older = true;
    }
    const bool holder_once = !any_before || a > before_max;
    removed[i] = (unsigned char)((any_before && a < before_max) || (holder_once && later_older));
  }
}

__global__ void k_kpc_apply_removed(int* age, const unsigned char* removed, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (removed[i]) age[i] = 0;
}

__constant__ signed char c_true_ring[16][2] = {{-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}, {0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3}, {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}};

// fast.hpp:36-77 (true ring), entries with a score under min_score are removed
__global__ void k_kpc_score_filter(Img im, int th, int min_score, const vppb_int2* pos, int* age, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned char* p = im.base + (long long)pos[i].r * im.pitch + pos[i].c;
    const int v = *p;
    int sum_inf = 0, sum_sup = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const int diff = v - (int)p[(long long)c_true_ring[k][0] * im.pitch + c_true_ring[k][1]];
      if (diff < -th) sum_inf -= diff;
      else if (diff > th) sum_sup += diff;
    }
    if (max(sum_sup, sum_inf) < min_score) age[i] = 0;
  }
}

// one warp per entry zeroes its 2s x 2s square of the mask (the mask has a border of s pixels, so no clipping is needed)
__global__ void k_kpc_paint(Img mask, const vppb_int2* pos, int n, int s) {
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nwarps) {
    const int r = pos[i].r, c = pos[i].c;
    for (int k = lane; k < 4 * s * s; k += 32) {
      const int dr = k / (2 * s) - s, dc = k % (2 * s) - s;
      mask.base[(long long)(r + dr) * mask.pitch + (c + dc)] = 0;
    }
  }
}

__global__ void k_kpc_add(vppb_int2* pos, vppb_int2* vel, int* age, int n, const vppb_int2* det, const int* det_count, int capacity, int* count) {
  const int m = min(*det_count, capacity - n);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    pos[n + i] = det[i];
    vel[n + i] = vppb_int2{0, 0};
    age[n + i] = 1;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *count = n + m;
}

// exclusive scan of the alive flags of *count entries by ONE CTA: map[i] = new index or -1; count <- survivors
__global__ void __launch_bounds__(1024) k_kpc_scan(const int* age, int* map, int* count) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int n = *count;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = (i < n && age[i] > 0) ? 1 : 0;
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
    if (i < n) map[i] = v ? incl - 1 : -1;
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
  __syncthreads();
  if (threadIdx.x == 0) { map[n] = carry; }  // survivors, picked up by the gather kernel's thread 0
}

struct KpcArrays {
  vppb_int2 *pos, *vel;
  int *age, *tstart, *tlen, *thead;
  unsigned char* talive;
  vppb_float2* thist;
};

// gather the survivors into the other buffer set; entries that had no trajectory yet (index >= n_traj) start one at frame_id
__global__ void k_kpc_gather(KpcArrays src, KpcArrays dst, const int* map, int* count, int n_traj, int frame_id, int max_traj) {
  const int n = *count;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int j = map[i];
    if (j < 0) continue;
    dst.pos[j] = src.pos[i]; dst.vel[j] = src.vel[i]; dst.age[j] = src.age[i];
    if (i < n_traj) {
      dst.tstart[j] = src.tstart[i]; dst.tlen[j] = src.tlen[i]; dst.thead[j] = src.thead[i]; dst.talive[j] = src.talive[i];
      for (int k = 0; k < max_traj; k++) dst.thist[(long long)j * max_traj + k] = src.thist[(long long)i * max_traj + k];
    } else {
      dst.tstart[j] = frame_id; dst.tlen[j] = 0; dst.thead[j] = 0; dst.talive[j] = 1;
    }
  }
}
__global__ void k_kpc_set_count(const int* map, int* count) { *count = map[*count]; }

__global__ void k_kpc_traj(KpcArrays a, int n, int max_traj) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (a.age[i] > 0) {  // move_to: push front; pop the oldest beyond max_traj (keypoint_trajectory.hh)
      const int h = (a.thead[i] + max_traj - 1) % max_traj;  // the ring grows downwards: head = newest
      a.thist[(long long)i * max_traj + h] = vppb_float2{(float)a.pos[i].r, (float)a.pos[i].c};
      a.thead[i] = h;
      a.tlen[i] = min(a.tlen[i] + 1, max_traj);
    } else {
      a.talive[i] = 0;
    }
  }
}

__global__ void k_kpc_table(KpcArrays a, int n, int* out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    out[6 * i + 0] = a.pos[i].r; out[6 * i + 1] = a.pos[i].c; out[6 * i + 2] = a.age[i];
    out[6 * i + 3] = a.tstart[i]; out[6 * i + 4] = a.tlen[i]; out[6 * i + 5] = a.talive[i];
  }
}

static KpcArrays arrays(const Kpc* k, int which) {
  return KpcArrays{k->pos[which], k->vel[which], k->age[which], k->tstart[which], k->tlen[which], k->thead[which], k->talive[which], k->thist[which]};
}
static int grid_for(long long n, int threads = 256) {
  long long b = (n + threads - 1) / threads;
  const long long cap = (long long)sm_count() * 16;
  return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

}  // namespace vppb

using namespace vppb;

extern "C" {

int vppb_kpc_create(int32_t capacity, int32_t max_trajectory_length, void** handle) {
  VPPB_REQUIRE(handle && capacity > 0 && max_trajectory_length > 0, VPPB_E_ARG, "vppb_kpc_create: bad argument");
  Kpc* k = new Kpc();
  memset(k, 0, sizeof(*k));
  k->capacity = capacity; k->max_traj = max_trajectory_length;
  for (int w = 0; w < 2; w++) {
    VPPB_CUDA(cudaMalloc(&k->pos[w], (size_t)capacity * 8)); VPPB_CUDA(cudaMalloc(&k->vel[w], (size_t)capacity * 8));
    VPPB_CUDA(cudaMalloc(&k->age[w], (size_t)capacity * 4)); VPPB_CUDA(cudaMalloc(&k->tstart[w], (size_t)capacity * 4));
    VPPB_CUDA(cudaMalloc(&k->tlen[w], (size_t)capacity * 4)); VPPB_CUDA(cudaMalloc(&k->thead[w], (size_t)capacity * 4));
    VPPB_CUDA(cudaMalloc(&k->talive[w], (size_t)capacity));
    VPPB_CUDA(cudaMalloc(&k->thist[w], (size_t)capacity * max_trajectory_length * 8));
  }
  VPPB_CUDA(cudaMalloc(&k->map, ((size_t)capacity + 1) * 4));
  VPPB_CUDA(cudaMalloc(&k->next, (size_t)capacity * 4));
  VPPB_CUDA(cudaMalloc(&k->removed, (size_t)capacity));
  VPPB_CUDA(cudaMalloc(&k->count, 4));
  VPPB_CUDA(cudaMemset(k->count, 0, 4));
  *handle = k;
  return VPPB_OK;
}

int vppb_kpc_destroy(void* handle) {
  if (!handle) return VPPB_OK;
  Kpc* k = static_cast<Kpc*>(handle);
  for (int w = 0; w < 2; w++) {
    cudaFree(k->pos[w]); cudaFree(k->vel[w]); cudaFree(k->age[w]); cudaFree(k->tstart[w]); cudaFree(k->tlen[w]); cudaFree(k->thead[w]);
    cudaFree(k->talive[w]); cudaFree(k->thist[w]);
  }
  cudaFree(k->map); cudaFree(k->next); cudaFree(k->removed); cudaFree(k->count); cudaFree(k->cell_head);
  delete k;
  return VPPB_OK;
}

int32_t vppb_kpc_size(void* handle) { return handle ? static_cast<Kpc*>(handle)->n : 0; }
const vppb_int2* vppb_kpc_positions(void* handle) { Kpc* k = static_cast<Kpc*>(handle); return k ? k->pos[k->cur] : nullptr; }

int vppb_kpc_flow_update(void* handle, const vppb_int2* new_pos, const unsigned char* valid, int32_t nrows, int32_t ncols, void* stream) {
  VPPB_REQUIRE(handle, VPPB_E_ARG, "vppb_kpc_flow_update: NULL container");
  Kpc* k = static_cast<Kpc*>(handle);
  if (k->n == 0) return VPPB_OK;
  VPPB_REQUIRE(new_pos && valid, VPPB_E_ARG, "vppb_kpc_flow_update: NULL argument");
  k_kpc_flow<<<grid_for(k->n), 256, 0, as_stream(stream)>>>(k->pos[k->cur], k->vel[k->cur], k->age[k->cur], k->n, new_pos, valid, nrows, ncols);
  VPPB_LAUNCH_CHECK("vppb_kpc_flow_update");
  return VPPB_OK;
}

int vppb_kpc_merge(void* handle, int32_t nrows, int32_t ncols, int32_t spacing, void* stream) {
  VPPB_REQUIRE(handle && spacing > 0, VPPB_E_ARG, "vppb_kpc_merge: bad argument");
  Kpc* k = static_cast<Kpc*>(handle);
  if (k->n == 0) return VPPB_OK;
  const int stride = ncols / spacing + 3;
  const long long cells = (long long)(nrows / spacing + 3) * stride;
  if (cells > k->cells_cap) {
    if (k->cell_head) VPPB_CUDA(cudaFree(k->cell_head));
    VPPB_CUDA(cudaMalloc(&k->cell_head, (size_t)cells * 4));
    k->cells_cap = cells;
  }
  cudaStream_t st = as_stream(stream);
  k_kpc_fill_int<<<grid_for(cells), 256, 0, st>>>(k->cell_head, cells, -1);
  k_kpc_link<<<grid_for(k->n), 256, 0, st>>>(k->pos[k->cur], k->n, spacing, stride, k->cell_head, k->next);
  k_kpc_merge<<<grid_for(k->n), 256, 0, st>>>(k->pos[k->cur], k->age[k->cur], k->n, spacing, stride, k->cell_head, k->next, k->removed);
  k_kpc_apply_removed<<<grid_for(k->n), 256, 0, st>>>(k->age[k->cur], k->removed, k->n);
  VPPB_LAUNCH_CHECK("vppb_kpc_merge");
  return VPPB_OK;
}

int vppb_kpc_score_filter(void* handle, const vppb_img* img, int32_t th, int32_t min_score, void* stream) {
  VPPB_REQUIRE(handle && img && img->base && img->elem_bytes == 1, VPPB_E_ARG, "vppb_kpc_score_filter: bad argument");
  VPPB_REQUIRE(img->border >= 3, VPPB_E_BORDER, "vppb_kpc_score_filter: border %d < 3", img->border);
  Kpc* k = static_cast<Kpc*>(handle);
  if (k->n == 0) return VPPB_OK;
  k_kpc_score_filter<<<grid_for(k->n), 256, 0, as_stream(stream)>>>(view(img), th, min_score, k->pos[k->cur], k->age[k->cur], k->n);
  VPPB_LAUNCH_CHECK("vppb_kpc_score_filter");
  return VPPB_OK;
}

int vppb_kpc_paint_mask(void* handle, const vppb_img* mask, int32_t spacing, void* stream) {
  VPPB_REQUIRE(handle && mask && mask->base && mask->elem_bytes == 1 && spacing > 0, VPPB_E_ARG, "vppb_kpc_paint_mask: bad argument");
  VPPB_REQUIRE(mask->border >= spacing, VPPB_E_BORDER, "vppb_kpc_paint_mask: the mask needs a border of keypoint_spacing = %d pixels", spacing);
  Kpc* k = static_cast<Kpc*>(handle);
  const unsigned char one = 1;
  int rc = vppb_fill(mask, &one, 1, stream);  // fill_with_border(mask, 1)
  if (rc) return rc;
  if (k->n == 0) return VPPB_OK;
  const long long blocks = ((long long)k->n + 7) / 8;
  k_kpc_paint<<<(int)(blocks < (long long)sm_count() * 16 ? blocks : (long long)sm_count() * 16), 256, 0, as_stream(stream)>>>(view(mask), k->pos[k->cur], k->n, spacing);
  VPPB_LAUNCH_CHECK("vppb_kpc_paint_mask");
  return VPPB_OK;
}

// append the detections (device array + device count, as vppb_fast9_u8_async leaves them), compact, start trajectories for
// the newcomers; reads the new entry count back (4 bytes, one synchronisation) because the host sizes later launches with it
int vppb_kpc_add_and_compact(void* handle, const vppb_int2* detections, const int32_t* det_count_dev, int32_t max_detections, int32_t frame_id, void* stream) {
  VPPB_REQUIRE(handle && det_count_dev && (detections || max_detections == 0), VPPB_E_ARG, "vppb_kpc_add_and_compact: NULL argument");
  Kpc* k = static_cast<Kpc*>(handle);
  cudaStream_t st = as_stream(stream);
  const int cur = k->cur, n_traj = k->n;
  const long long upper = std::min<long long>((long long)k->n + max_detections, k->capacity);
  k_kpc_add<<<grid_for(std::max(1, max_detections)), 256, 0, st>>>(k->pos[cur], k->vel[cur], k->age[cur], k->n, detections, det_count_dev, k->capacity, k->count);
  k_kpc_scan<<<1, 1024, 0, st>>>(k->age[cur], k->map, k->count);
  k_kpc_gather<<<grid_for(std::max<long long>(1, upper)), 256, 0, st>>>(arrays(k, cur), arrays(k, cur ^ 1), k->map, k->count, n_traj, frame_id, k->max_traj);
  k_kpc_set_count<<<1, 1, 0, st>>>(k->map, k->count);
  VPPB_LAUNCH_CHECK("vppb_kpc_add_and_compact");
  int n = 0;
  VPPB_CUDA(cudaMemcpyAsync(&n, k->count, 4, cudaMemcpyDeviceToHost, st));
  VPPB_CUDA(cudaStreamSynchronize(st));
  k->cur = cur ^ 1;
  k->n = n;
  return VPPB_OK;
}

int vppb_kpc_trajectories_update(void* handle, void* stream) {
  VPPB_REQUIRE(handle, VPPB_E_ARG, "vppb_kpc_trajectories_update: NULL container");
  Kpc* k = static_cast<Kpc*>(handle);
  if (k->n == 0) return VPPB_OK;
  k_kpc_traj<<<grid_for(k->n), 256, 0, as_stream(stream)>>>(arrays(k, k->cur), k->n, k->max_traj);
  VPPB_LAUNCH_CHECK("vppb_kpc_trajectories_update");
  return VPPB_OK;
}

// rows of 6 ints per entry (row, col, age, trajectory start frame, trajectory length, trajectory alive) into a DEVICE buffer
int vppb_kpc_state_table(void* handle, int32_t* table_dev, void* stream) {
  VPPB_REQUIRE(handle, VPPB_E_ARG, "vppb_kpc_state_table: NULL container");
  Kpc* k = static_cast<Kpc*>(handle);
  if (k->n == 0) return VPPB_OK;
  VPPB_REQUIRE(table_dev, VPPB_E_ARG, "vppb_kpc_state_table: NULL output");
  k_kpc_table<<<grid_for(k->n), 256, 0, as_stream(stream)>>>(arrays(k, k->cur), k->n, table_dev);
  VPPB_LAUNCH_CHECK("vppb_kpc_state_table");
  return VPPB_OK;
}

}  // extern "C"
