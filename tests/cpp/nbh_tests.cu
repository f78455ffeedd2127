// C++ host API added late in round 1: box_nbh2d<V, R, C> (the reference's tests/box_nbh2d.cc point form and the pixel_wise
// range form of benchmarks/box_5x5_filter.cc:163-172), window.hh, cast_to_float / zero, colorspace_conversions.hh
// (tests/colorspace_conversions.cc:8-23 + the fused ingest).  Exit code 0 = all asserts held.
#undef NDEBUG
#include <cassert>
#include <cstdio>
#include <iostream>
#include <vpp/vpp.hh>

using namespace vpp;

static void test_point_form() {  // tests/box_nbh2d.cc:8-29, statement for statement
  image2d<int> A(3, 3);
  auto nbh = box_nbh2d<int, 3, 3>(A, vint2{1, 1});

  fill(A, 1);  // device write after the accessor was made: the accessor must see it

  nbh.for_all([](int& p) { p = 2; });
  nbh.north() = 3;
  nbh.east() = 4;
  nbh.south() = 5;
  nbh.west() = 6;

  assert(A(0, 0) == 2);
  assert(A(0, 1) == 3);
  assert(A(0, 2) == 2);

  assert(A(1, 0) == 6);
  assert(A(1, 1) == 2);
  assert(A(1, 2) == 4);

  assert(A(2, 0) == 2);
  assert(A(2, 1) == 5);
  assert(A(2, 2) == 2);

  // and the device sees what the accessor wrote on the host
  image2d<int> B(3, 3);
  pixel_wise(B, A) | [=] VPP_KERNEL(int& b, int& a) { b = a * 10; };
  assert(B(1, 2) == 40 && B(2, 1) == 50 && B(0, 0) == 20);
}

static void test_range_form() {  // benchmarks/box_5x5_filter.cc:163-172
  image2d<int> A(40, 50, _border = 2), B(40, 50), B2(40, 50), N4(40, 50), V(40, 50);
  for (auto p : A.domain()) A(p) = (p[0] * 31 + p[1] * 17) % 1000;
  fill_border_mirror(A);

  auto Anbh = box_nbh2d<int, 5, 5>(A);
  pixel_wise(B, Anbh) | [=] VPP_KERNEL(int& b, box_nbh2d_kernel<int, 5, 5>& a_nbh) {
    int sum = 0;
    a_nbh.for_all([&sum](int& n) { sum += n; });
    b = sum / 25;
  };
  // the same filter through relative_access (benchmarks/box_5x5_filter2.cc:71-81): identical, border pixels included
  pixel_wise(B2, relative_access(A)) | [=] VPP_KERNEL(int& b, relative_access_kernel<int> a) {
    int sum = 0;
    for (int i = -2; i <= 2; i++)
      for (int j = -2; j <= 2; j++) sum += a(i, j);
    b = sum / 25;
  };
  for (auto p : B.domain()) assert(B(p) == B2(p));
  for (int r = 0; r < 40; r++)
    for (int c = 0; c < 50; c++) {
      int sum = 0;
      for (int d = -2; d <= 2; d++)
        for (int e = -2; e <= 2; e++) sum += A(r + d, c + e);  // border rows / columns are addressable on the host mirror
      assert(B(r, c) == sum / 25);
    }

  // named neighbours, accessor taken by value, 3x3 window
  pixel_wise(N4, box_nbh2d<int, 3, 3>(A)) | [=] VPP_KERNEL(int& o, box_nbh2d_kernel<int, 3, 3> n) {
    o = n.north() * 1000 + n.south() * 100 + n.east() * 10 + n.west() - n(0, 0);
  };
  for (int r = 0; r < 40; r++)
    for (int c = 0; c < 50; c++) assert(N4(r, c) == A(r - 1, c) * 1000 + A(r + 1, c) * 100 + A(r, c + 1) * 10 + A(r, c - 1) - A(r, c));

  // writing through the accessor: every pixel stamps itself (window 1x1 reaches only the centre)
  fill(V, 0);
  pixel_wise(box_nbh2d<int, 1, 1>(V), V.domain()) | [=] VPP_KERNEL(box_nbh2d_kernel<int, 1, 1>& n, vint2 p) {
    n.for_all([p](int& v) { v = p[0] * 100 + p[1]; });
  };
  for (auto p : V.domain()) assert(V(p) == p[0] * 100 + p[1]);
}

static void test_windows() {  // vpp/core/window.hh:11-62
  assert(c9.size() == 9 && c8.size() == 8 && c5.size() == 5 && c4.size() == 4);
  // the reference's member order (raster)
  const int e8[8][2] = {{-1, -1}, {-1, 0}, {-1, 1}, {0, -1}, {0, 1}, {1, -1}, {1, 0}, {1, 1}};
  const int e5[5][2] = {{-1, 0}, {0, -1}, {0, 0}, {0, 1}, {1, 0}};
  const int e4[4][2] = {{-1, 0}, {0, -1}, {0, 1}, {1, 0}};
  for (int i = 0; i < 8; i++) assert(c8[i][0] == e8[i][0] && c8[i][1] == e8[i][1]);
  for (int i = 0; i < 5; i++) assert(c5[i][0] == e5[i][0] && c5[i][1] == e5[i][1]);
  for (int i = 0; i < 4; i++) assert(c4[i][0] == e4[i][0] && c4[i][1] == e4[i][1]);
  for (int i = 0; i < 9; i++) assert(c9[i][0] == i / 3 - 1 && c9[i][1] == i % 3 - 1);
  int host_sum = 0;
  foreach(c9, [&](vint2 n) { host_sum += 10 * n[0] + n[1] + 11; });
  assert(host_sum == 9 * 11);

  image2d<int> A(20, 30, _border = 1), S4(20, 30), S8(20, 30);
  for (auto p : A.domain()) A(p) = p[0] * 37 + p[1] * 3;
  fill_border_mirror(A);
  pixel_wise(S4, relative_access(A)) | [=] VPP_KERNEL(int& s, relative_access_kernel<int> a) {
    int acc = 0;
    foreach(c4, [&](vint2 n) { acc += a(n); });
    s = acc;
  };
  pixel_wise(S8, relative_access(A)) | [=] VPP_KERNEL(int& s, relative_access_kernel<int> a) {
    int acc = 0, k = 0;
    foreach(c8, [&](vint2 n) { acc += a(n) * (++k); });  // order-sensitive
    s = acc;
  };
  for (int r = 0; r < 20; r++)
    for (int c = 0; c < 30; c++) {
      assert(S4(r, c) == A(r - 1, c) + A(r, c - 1) + A(r, c + 1) + A(r + 1, c));
      int acc = 0;
      for (int i = 0; i < 8; i++) acc += A(r + e8[i][0], c + e8[i][1]) * (i + 1);
      assert(S8(r, c) == acc);
    }
  static_assert(std::is_same<cast_to_float<int>, float>::value && std::is_same<cast_to_float<vuchar3>, vfloat3>::value, "cast_to_float");
  vint2 z = zero<vint2>();
  assert(z[0] == 0 && z[1] == 0 && (int)zero<int>() == 0);
}

static void test_colorspace() {  // tests/colorspace_conversions.cc:8-23 + the fused ingest
  image2d<vuchar3> i1(100, 100);
  unsigned char i = 0;
  for (vint2 p : i1.domain()) { i1(p) = vuchar3{i, i, i}; i++; }
  image2d<vuchar1> i2 = rgb_to_graylevel<vuchar1>(i1);
  i = 0;
  for (vint2 p : i1.domain()) { assert(i2(p)[0] == i); i++; }

  // truncating mean of three different channels, border converted too (domain_with_border), RGBA ignores alpha
  image2d<vuchar3> c3(37, 53, _border = 2);
  image2d<vuchar4> c4(37, 53, _border = 2);
  for (vint2 p : c3.domain()) {
    c3(p) = vuchar3{(unsigned char)(p[0] * 7 + p[1]), (unsigned char)(p[1] * 5), (unsigned char)(255 - p[0])};
    c4(p) = vuchar4{c3(p)[0], c3(p)[1], c3(p)[2], (unsigned char)(p[0] ^ p[1])};
  }
  fill_border_mirror(c3);
  fill_border_mirror(c4);
  image2d<unsigned char> g3 = rgb_to_graylevel<unsigned char>(c3), g4 = rgb_to_graylevel<unsigned char>(c4);
  assert(g3.border() == 2 && g3.alignment() == c3.alignment());
  for (int r = -2; r < 39; r++)
    for (int c = -2; c < 55; c++) {
      const int e = ((int)c3(r, c)[0] + (int)c3(r, c)[1] + (int)c3(r, c)[2]) / 3;
      assert(g3(r, c) == e && g4(r, c) == e);
    }
  // the generic (pixel_wise) path: int channels, int gray level - same arithmetic, any magnitude
  image2d<vint3> wide(9, 11);
  for (vint2 p : wide.domain()) wide(p) = vint3{p[0] * 1000, p[1] * 1000 + 1, 7};
  image2d<int> gw = rgb_to_graylevel<int>(wide);
  for (vint2 p : wide.domain()) assert(gw(p) == (p[0] * 1000 + p[1] * 1000 + 1 + 7) / 3);
  // ingest: clone(_border = 3) + fill_border_mirror + rgb_to_graylevel in one launch
  image2d<vuchar3> frame(37, 53);
  for (vint2 p : frame.domain()) frame(p) = c3(p);
  image2d<unsigned char> ing = ingest_rgb_frame(frame, 3);
  auto ref = clone(frame, _border = 3);
  fill_border_mirror(ref);
  image2d<unsigned char> gref = rgb_to_graylevel<unsigned char>(ref);
  assert(ing.border() == 3);
  for (int r = -3; r < 40; r++)
    for (int c = -3; c < 56; c++) assert(ing(r, c) == gref(r, c));
  std::puts("colorspace ok");
}

int main() {
  vppb_check(vppb_init(0));
  test_windows();
  test_colorspace();
  test_point_form();
  test_range_form();
  std::printf("ALL OK\n");
  return 0;
}
