// TEST INFRASTRUCTURE ONLY: state of the CPU thread-by-thread emulation (see common.cuh in this directory).
#include "common.cuh"

#include <stdarg.h>

thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
namespace emu { bool reverse_order = false; }

namespace vppb {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("emulated CUDA error %d in %s", (int)e, what);
  return VPPB_E_CUDA;
}
}  // namespace vppb

extern "C" {
const char* vppb_last_error(void) { return vppb::g_err; }
void vppb_emu_set_reverse(int on) { emu::reverse_order = on != 0; }
}
