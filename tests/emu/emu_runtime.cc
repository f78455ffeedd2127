// TEST INFRASTRUCTURE ONLY: the emulated machine behind tests/emu/cuda_runtime.h.
// One block at a time; each CUDA thread of the block is a ucontext fiber on the calling OS thread.
#include <cuda_runtime.h>

#include "tma.cuh"

#include <sys/mman.h>
#include <ucontext.h>

#include <map>
#include <vector>

thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace emu {

bool reverse_order = false;
unsigned shuffle_seed = 0;  // != 0: every scheduler round visits the threads of the block in a fresh pseudo-random order

namespace {
size_t stack_bytes() {  // per-fiber stack; VPPB_EMU_STACK_KB overrides (the ASan run uses small stacks: its swapcontext
  static size_t n = 0;  // interceptor pays per byte of stack)
  if (!n) {
    const char* e = getenv("VPPB_EMU_STACK_KB");
    n = (size_t)(e && atoi(e) >= 16 ? atoi(e) : 512) * 1024;
  }
  return n;
}

struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  bool done = true;
};
struct Collective {      // one in-flight warp collective per (warp, mask)
  unsigned arrived = 0;  // lanes that have deposited their value for the current generation
  unsigned long long gen = 0;
  uint64_t slot[32];
  uint64_t snap[32];     // values of the completed generation
  unsigned snap_from = 0;
};
struct Block {
  unsigned n = 0, alive = 0;
  std::vector<Fiber> fibers;
  std::vector<std::map<unsigned, Collective>> warps;
  std::vector<unsigned> warp_alive;  // live-lane mask per warp
  unsigned bar_arrived = 0;
  unsigned long long bar_gen = 0;
  unsigned long long progress = 0;   // bumped whenever anything completes; a round without a bump = deadlock
  unsigned current = 0;
  ucontext_t sched;
  void (*entry)(void*) = nullptr;
  void* arg = nullptr;
  struct MBar { uint32_t phase = 0, init = 0; int pending = 0; long long tx = 0; };
  std::map<uint64_t*, MBar> mbars;
  std::vector<unsigned> order;
};
Block g;
void* g_smem = nullptr;
size_t g_smem_cap = 0, g_smem_bytes = 0;

void mbar_maybe_complete(Block::MBar& m) {
  if (m.pending == 0 && m.tx == 0) {
    m.phase ^= 1u;
    m.pending = (int)m.init;
    g.progress++;
  }
}

void release_barrier_if_complete() {
  if (g.alive > 0 && g.bar_arrived == g.alive) {
    g.bar_arrived = 0;
    g.bar_gen++;
    g.progress++;
  }
}

void fiber_main() {
  g.entry(g.arg);
  const unsigned t = g.current;
  g.fibers[t].done = true;
  g.alive--;
  g.warp_alive[t / 32] &= ~(1u << (t % 32));
  g.progress++;
  release_barrier_if_complete();  // threads that exit no longer take part in __syncthreads
  swapcontext(&g.fibers[t].ctx, &g.sched);
}
}  // namespace

void yield() { swapcontext(&g.fibers[g.current].ctx, &g.sched); }

int lane_id() { return (int)(g.current % 32); }

void block_barrier() {
  const unsigned long long gen = g.bar_gen;
  g.bar_arrived++;
  g.progress++;  // an arrival is progress too: a round in which nobody even arrives anywhere new is the deadlock
  release_barrier_if_complete();
  while (g.bar_gen == gen) yield();
}

unsigned warp_exchange(unsigned mask, uint64_t mine, uint64_t out[32]) {
  const unsigned w = g.current / 32, lane = g.current % 32;
  if (!((mask >> lane) & 1u)) {
    fail("a warp collective is called with a mask that does not name the calling lane");
  }
  Collective& c = g.warps[w][mask];
  const unsigned long long gen = c.gen;
  c.slot[lane] = mine;
  c.arrived |= 1u << lane;
  g.progress++;
  for (;;) {
    if (c.gen != gen) break;
    const unsigned need = mask & g.warp_alive[w];  // exited lanes are not waited for
    if ((c.arrived & need) == need) {
      memcpy(c.snap, c.slot, sizeof(c.snap));
      c.snap_from = c.arrived;
      c.arrived = 0;
      c.gen++;
      g.progress++;
      break;
    }
    yield();
  }
  memcpy(out, c.snap, sizeof(c.snap));
  return c.snap_from;
}

void fail(const char* what) {  // a rule of the machine was broken: report (stderr + $VPPB_EMU_LOG, pytest captures stderr) and stop
  fprintf(stderr, "emu: %s (block %u thread %u)\n", what, blockIdx.x, g.current);
  if (const char* path = getenv("VPPB_EMU_LOG")) {
    if (FILE* f = fopen(path, "a")) {
      fprintf(f, "emu: %s (block %u thread %u)\n", what, blockIdx.x, g.current);
      fclose(f);
    }
  }
  abort();
}

void* dyn_smem() { return g_smem; }
void set_dyn_smem(size_t bytes) {
  if (bytes > g_smem_cap) {
    free(g_smem);
    if (posix_memalign(&g_smem, 1024, bytes)) fail("dynamic shared memory allocation");
    g_smem_cap = bytes;
  }
  g_smem_bytes = bytes;
}

// mbarrier (phase completes when the pending arrivals and the expected transaction bytes both reach zero)
void mbar_init(uint64_t* bar, uint32_t count) {
  Block::MBar& m = g.mbars[bar];
  m = Block::MBar();
  m.init = count;
  m.pending = (int)count;
}
void mbar_arrive(uint64_t* bar, uint32_t expect_tx) {
  auto it = g.mbars.find(bar);
  if (it == g.mbars.end()) fail("mbarrier used before mbarrier.init");
  it->second.tx += expect_tx;
  if (--it->second.pending < 0) fail("more arrivals than the mbarrier was initialised for");
  g.progress++;
  mbar_maybe_complete(it->second);
}
void mbar_complete_tx(uint64_t* bar, uint32_t bytes) {
  auto it = g.mbars.find(bar);
  if (it == g.mbars.end()) fail("TMA completes on an mbarrier that was not initialised");
  it->second.tx -= bytes;
  g.progress++;
  mbar_maybe_complete(it->second);
}
bool mbar_phase_done(uint64_t* bar, uint32_t parity) {
  auto it = g.mbars.find(bar);
  if (it == g.mbars.end()) return false;  // not initialised yet: another thread will do it (the wait spins)
  return it->second.phase != (parity & 1u);
}

void run_block(unsigned nthreads, void (*entry)(void*), void* arg) {
  static bool env_read = false;
  if (!env_read) {  // programs that cannot call vppb_emu_set_* (the C++ test binaries) pick the schedule from the environment
    env_read = true;
    if (const char* e = getenv("VPPB_EMU_SHUFFLE")) shuffle_seed = (unsigned)atoi(e);
    if (const char* e = getenv("VPPB_EMU_REVERSE")) reverse_order = atoi(e) != 0;
  }
  if (g.fibers.size() < nthreads) g.fibers.resize(nthreads);
  g.mbars.clear();
  g.n = g.alive = nthreads;
  g.entry = entry;
  g.arg = arg;
  g.bar_arrived = 0;
  const unsigned nwarps = (nthreads + 31) / 32;
  g.warps.assign(nwarps, {});
  g.warp_alive.assign(nwarps, 0);
  for (unsigned t = 0; t < nthreads; t++) {
    Fiber& f = g.fibers[t];
    if (!f.stack) {
      f.stack = static_cast<char*>(mmap(nullptr, stack_bytes(), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_STACK, -1, 0));
      if (f.stack == MAP_FAILED) { perror("emu: mmap"); abort(); }
    }
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = stack_bytes();
    f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, fiber_main, 0);
    f.done = false;
    g.warp_alive[t / 32] |= 1u << (t % 32);
  }
  while (g.alive > 0) {
    const unsigned long long before = g.progress;
    if (shuffle_seed) {
      if (g.order.size() != nthreads) { g.order.resize(nthreads); for (unsigned k = 0; k < nthreads; k++) g.order[k] = k; }
      for (unsigned k = nthreads; k > 1; k--) {  // Fisher-Yates with an xorshift generator
        shuffle_seed ^= shuffle_seed << 13; shuffle_seed ^= shuffle_seed >> 17; shuffle_seed ^= shuffle_seed << 5;
        std::swap(g.order[k - 1], g.order[shuffle_seed % k]);
      }
    }
    for (unsigned k = 0; k < nthreads; k++) {
      const unsigned t = shuffle_seed ? g.order[k] : (reverse_order ? nthreads - 1 - k : k);
      if (g.fibers[t].done) continue;
      g.current = t;
      threadIdx = dim3{t % blockDim.x, t / blockDim.x, 0};  // 2-D blocks: x fastest
      swapcontext(&g.sched, &g.fibers[t].ctx);
    }
    if (g.progress == before && g.alive > 0) {
      fail("deadlock: the remaining threads of the block wait at a barrier / warp collective / mbarrier nobody else will reach");
    }
  }
}

}  // namespace emu

// "driver": cuTensorMapEncodeTiled records its arguments, checking what the real one rejects
static CUresult emu_encode_tiled(CUtensorMap* map, CUtensorMapDataType dt, cuuint32_t rank, void* origin, const cuuint64_t* gdim,
                                 const cuuint64_t* gstride, const cuuint32_t* box, const cuuint32_t* estride, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
  static const int bytes_of[] = {1, 2, 4, 4, 8};
  if (rank != 2 || ((uintptr_t)origin % 16) != 0 || (gstride[0] % 16) != 0 || box[0] == 0 || box[1] == 0 || box[0] > 256 || box[1] > 256 ||
      estride[0] != 1 || estride[1] != 1 || (((uint64_t)box[0] * bytes_of[dt]) % 16) != 0)
    return 1;  // CUDA_ERROR_INVALID_VALUE
  emu::TensorMap2d t{static_cast<unsigned char*>(origin), (uint64_t)bytes_of[dt], gdim[0], gdim[1], gstride[0], box[0], box[1], 0x7e4503a9ull};
  static_assert(sizeof(t) <= sizeof(CUtensorMap), "record fits the opaque map");
  memcpy(map, &t, sizeof(t));
  return CUDA_SUCCESS;
}
cudaError_t cudaGetDriverEntryPoint(const char* symbol, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* q) {
  const bool ok = strcmp(symbol, "cuTensorMapEncodeTiled") == 0;
  *fn = ok ? reinterpret_cast<void*>(&emu_encode_tiled) : nullptr;
  *q = ok ? cudaDriverEntryPointSuccess : cudaDriverEntryPointSymbolNotFound;
  return cudaSuccess;
}

extern "C" void vppb_emu_set_reverse(int on) { emu::reverse_order = on != 0; }
extern "C" void vppb_emu_set_shuffle(unsigned seed) { emu::shuffle_seed = seed; }
