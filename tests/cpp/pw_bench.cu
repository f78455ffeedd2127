// Timing of the C++ host API's generic lowering against the named C-ABI kernel (GPU box only):
//   pixel_wise(A, B, C) | [=] VPP_KERNEL (int& a, int& b, int& c) { a = b + c; }     (benchmarks/image_add.cc:51-57)
//   vppb_pw_add_i32(A, B, C)
// at 3840 x 2160, 4 image triples cycled (398 MB > L2), CUDA events.  Prints one JSON line.
#include <cstdio>
#include <vector>
#include <vpp/vpp.hh>

using namespace vpp;

template <typename F>
static float time_us(F f, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f(); f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; i++) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  return ms * 1000.f / reps;
}

int main() {
  vppb_check(vppb_init(0));
  const int H = 2160, W = 3840, N = 4;
  std::vector<image2d<int>> A, B, C;
  for (int i = 0; i < N; i++) {
    A.emplace_back(H, W); B.emplace_back(H, W); C.emplace_back(H, W);
    fill(B[i], 1000 + i); fill(C[i], 7 * i);
  }
  const float us_lambda = time_us([&] { for (int i = 0; i < N; i++) pixel_wise(A[i], B[i], C[i]) | [=] VPP_KERNEL(int& a, int& b, int& c) { a = b + c; }; }, 20) / N;
  bool ok = true;
  for (int i = 0; i < N; i++) ok = ok && A[i](H - 1, W - 1) == 1000 + 8 * i && A[i](0, 0) == 1000 + 8 * i && A[i](1000, 1777) == 1000 + 8 * i;
  const float us_named = time_us([&] { for (int i = 0; i < N; i++) vppb_check(vppb_pw_add_i32(A[i].device_write(), B[i].device_read(), C[i].device_read(), nullptr)); }, 20) / N;
  // block_wise: 10 x 10 blocks over a 4K image, device callback (one launch) - the per-block sum written to the block
  image2d<int> blk(H, W);
  fill(blk, 1);
  const float us_block = time_us([&] {
    block_wise(vint2(10, 10), blk) | [=] VPP_KERNEL(block_view<int> b) {
      int s = 0;
      for (int r = 0; r < b.nrows(); r++)
        for (int c = 0; c < b.ncols(); c++) s += b(r, c);
      b(0, 0) = s;
    };
  }, 5);
  const double bytes = 12.0 * H * W;
  std::printf("{\"pixel_wise_lambda_add_4k_us\": %.2f, \"vppb_pw_add_i32_4k_us\": %.2f, \"lambda_GBps\": %.1f, \"named_GBps\": %.1f, \"lambda_over_named\": %.3f, "
              "\"block_wise_10x10_4k_device_us\": %.2f, \"results_ok\": %s}\n",
              us_lambda, us_named, bytes / us_lambda / 1e3, bytes / us_named / 1e3, us_lambda / us_named, us_block, ok ? "true" : "false");
  return ok ? 0 : 1;
}
