// Multi-GPU plumbing of the row-tile sharding: CUDA IPC so that a rank can map its neighbours' tiles (the
// box kernel then reads the halo rows straight from peer memory over NVLink, vppb_box5x5_u8c3_tiles), and the single
// grouped NCCL halo exchange per frame set for the kernels that want the halo rows materialised in the tile's border.
// The reference has no multi-GPU path (its parallelism is OpenMP over rows, vpp/core/pixel_wise.hpp:85-105); this is
// the row-tile extension north_star asks for.  NCCL is bound at run time (dlopen of the libnccl.so.2 already in the
// process or on the loader path): the library itself has no link-time dependency on it.
#include "common.cuh"

#include <dlfcn.h>

namespace vppb {

// ---- the few NCCL declarations used (nccl.h 2.x ABI: ncclUniqueId = 128 bytes, ncclUint8 = 1, ncclSuccess = 0)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
typedef int ncclDataType_t;
struct Nccl {
  void* so = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static Nccl* nccl() {
  static Nccl n;
  static int state = 0;  // 0 = not tried, 1 = ok, -1 = unavailable
  if (state == 0) {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      n.so = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (n.so) break;
    }
    if (n.so) {
      n.GetUniqueId = (decltype(n.GetUniqueId))dlsym(n.so, "ncclGetUniqueId");
      n.CommInitRank = (decltype(n.CommInitRank))dlsym(n.so, "ncclCommInitRank");
      n.CommDestroy = (decltype(n.CommDestroy))dlsym(n.so, "ncclCommDestroy");
      n.GroupStart = (decltype(n.GroupStart))dlsym(n.so, "ncclGroupStart");
      n.GroupEnd = (decltype(n.GroupEnd))dlsym(n.so, "ncclGroupEnd");
      n.Send = (decltype(n.Send))dlsym(n.so, "ncclSend");
      n.Recv = (decltype(n.Recv))dlsym(n.so, "ncclRecv");
      n.GetErrorString = (decltype(n.GetErrorString))dlsym(n.so, "ncclGetErrorString");
    }
    state = (n.so && n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.GroupStart && n.GroupEnd && n.Send && n.Recv) ? 1 : -1;
  }
  return state == 1 ? &n : nullptr;
}

static int nccl_fail(ncclResult_t r, const char* what) {
  Nccl* n = nccl();
  set_error("NCCL error %d (%s) in %s", (int)r, (n && n->GetErrorString) ? n->GetErrorString(r) : "?", what);
  return VPPB_E_NCCL;
}

#define VPPB_NCCL(call)                                       \
  do {                                                        \
    ncclResult_t r__ = (call);                                \
    if (r__ != 0) return ::vppb::nccl_fail(r__, #call);       \
  } while (0)

}  // namespace vppb

using namespace vppb;

extern "C" {

// ------------------------------------------------------------------ CUDA IPC
int vppb_ipc_export(const vppb_img* img, void* handle64, int64_t* offset_out) {
  VPPB_REQUIRE(img && img->base && img->alloc && handle64 && offset_out, VPPB_E_ARG, "vppb_ipc_export: needs an image that owns its allocation (vppb_alloc)");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  VPPB_CUDA(cudaIpcGetMemHandle(&h, img->alloc));
  memcpy(handle64, &h, sizeof(h));
  *offset_out = (int64_t)(static_cast<const unsigned char*>(img->base) - static_cast<const unsigned char*>(img->alloc));
  return VPPB_OK;
}

int vppb_ipc_open(const void* handle64, int64_t offset, const vppb_img* geometry, vppb_img* out) {
  VPPB_REQUIRE(handle64 && geometry && out && offset >= 0, VPPB_E_ARG, "vppb_ipc_open: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* mapped = nullptr;
  VPPB_CUDA(cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess));
  *out = *geometry;
  out->alloc = mapped;  // what vppb_ipc_close unmaps
  out->base = static_cast<unsigned char*>(mapped) + offset;
  return VPPB_OK;
}

int vppb_ipc_close(vppb_img* img) {
  VPPB_REQUIRE(img, VPPB_E_ARG, "vppb_ipc_close: NULL");
  if (img->alloc) VPPB_CUDA(cudaIpcCloseMemHandle(img->alloc));
  img->alloc = nullptr;
  img->base = nullptr;
  return VPPB_OK;
}

// ------------------------------------------------------------------ NCCL communicator + halo exchange
int vppb_comm_unique_id(void* id128) {
  VPPB_REQUIRE(id128, VPPB_E_ARG, "vppb_comm_unique_id: NULL");
  Nccl* n = nccl();
  VPPB_REQUIRE(n, VPPB_E_NCCL, "libnccl.so.2 could not be loaded (%s)", dlerror() ? dlerror() : "symbols missing");
  ncclUniqueId id;
  VPPB_NCCL(n->GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return VPPB_OK;
}

int vppb_comm_init(const void* id128, int32_t rank, int32_t nranks, void** comm_out) {
  VPPB_REQUIRE(id128 && comm_out && nranks > 0 && rank >= 0 && rank < nranks, VPPB_E_ARG, "vppb_comm_init: bad argument");
  Nccl* n = nccl();
  VPPB_REQUIRE(n, VPPB_E_NCCL, "libnccl.so.2 could not be loaded");
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c = nullptr;
  VPPB_NCCL(n->CommInitRank(&c, nranks, id, rank));
  *comm_out = c;
  return VPPB_OK;
}

int vppb_comm_destroy(void* comm) {
  if (!comm) return VPPB_OK;
  Nccl* n = nccl();
  VPPB_REQUIRE(n, VPPB_E_NCCL, "libnccl.so.2 could not be loaded");
  VPPB_NCCL(n->CommDestroy(static_cast<ncclComm_t>(comm)));
  return VPPB_OK;
}

// One grouped send/recv with both neighbours for n row tiles: the `halo` top domain rows of every tile go to rank - 1 and
// land in its bottom border rows, the `halo` bottom rows go to rank + 1 and land in its top border rows.  The rows of a
// pitched tile (with their column border) are contiguous, so nothing is packed: NCCL reads and writes the images.
int vppb_halo_exchange(void* comm, int32_t rank, int32_t nranks, const vppb_img* imgs, int32_t n, int32_t halo, void* stream) {
  VPPB_REQUIRE(comm && imgs && n >= 0 && nranks > 0 && rank >= 0 && rank < nranks, VPPB_E_ARG, "vppb_halo_exchange: bad argument");
  Nccl* nc = nccl();
  VPPB_REQUIRE(nc, VPPB_E_NCCL, "libnccl.so.2 could not be loaded");
  for (int i = 0; i < n; i++) {
    VPPB_REQUIRE(imgs[i].base, VPPB_E_ARG, "vppb_halo_exchange: NULL image %d", i);
    VPPB_REQUIRE(halo > 0 && halo <= imgs[i].border && halo <= imgs[i].nrows, VPPB_E_BORDER, "vppb_halo_exchange: halo %d exceeds the border %d or the tile (image %d)", halo,
                 imgs[i].border, i);
  }
  if (n == 0 || nranks == 1) return VPPB_OK;
  ncclComm_t c = static_cast<ncclComm_t>(comm);
  cudaStream_t st = as_stream(stream);
  const int up = rank - 1, down = rank + 1;
  VPPB_NCCL(nc->GroupStart());
  for (int i = 0; i < n; i++) {
    const vppb_img* im = &imgs[i];
    long long bs = (long long)im->border * im->elem_bytes;
    if (im->align > 0 && bs % im->align) bs += im->align - (bs % im->align);
    unsigned char* row0 = static_cast<unsigned char*>(im->base) - bs;  // start of the buffer row that holds image row 0
    const size_t bytes = (size_t)halo * im->pitch;
    if (up >= 0) {
      VPPB_NCCL(nc->Send(row0, bytes, 1 /* ncclUint8 */, up, c, st));
      VPPB_NCCL(nc->Recv(row0 - (long long)halo * im->pitch, bytes, 1, up, c, st));
    }
    if (down < nranks) {
      VPPB_NCCL(nc->Send(row0 + (long long)(im->nrows - halo) * im->pitch, bytes, 1, down, c, st));
      VPPB_NCCL(nc->Recv(row0 + (long long)im->nrows * im->pitch, bytes, 1, down, c, st));
    }
  }
  VPPB_NCCL(nc->GroupEnd());
  return VPPB_OK;
}

}  // extern "C"
