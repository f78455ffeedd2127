// C++ host API tests: the reference's own tests (tests/imageNd.cc, image2d.cc, pixel_wise.cc,
// block_wise.cc, fill.cc, border.cc, sum.cc) rewritten against vpp_b200/include/vpp with device
// kernels (VPP_KERNEL lambdas capturing by value).  Exit code 0 = all asserts held.
#undef NDEBUG
#include <cassert>
#include <cstdio>
#include <iostream>
#include <vpp/vpp.hh>

using namespace vpp;

template <typename V, typename U>
bool equals(image2d<V>& v, image2d<U>& u) {
  for (auto p : v.domain())
    if (v(p) != u(p)) return false;
  return true;
}

static void test_imageNd() {  // tests/imageNd.cc
  imageNd<int, 2> img_test({2, 3});
  std::vector<int> dims = {100, 200};
  imageNd<int, 2> img(dims);
  assert(img.domain().size(0) == dims[0]);
  assert(img.domain().size(1) == dims[1]);
  assert(&(*img.begin()) == &img(0, 0));
  for (int r = 0; r < img.domain().size(0); r++)
    for (int c = 0; c < img.domain().size(1); c++) assert(img.coords_to_offset(vint2(r, c)) == img.pitch() * r + c * (int)sizeof(int));
  for (int r = 0; r < 100; r++)
    for (int c = 0; c < 200; c++) { img(vint2(r, c)) = r * c; img(r, c) = r * c; }
  for (int r = 0; r < 99; r++)
    for (int c = 0; c < 199; c++) assert(img(vint2(r, c)) == r * c);

  int align_size = 256;
  imageNd<int, 2> img2(dims, _border = 1, _aligned = align_size);
  assert(!(img2.pitch() % align_size));                                   // tests/imageNd.cc:48-49 (device pitch)
  assert(!((unsigned long long)img2.device_read()->base % align_size));   // first pixel aligned (device address)

  int i = 0;
  for (auto& p : img) p = i++;
  auto img_clone = clone(img);
  auto img_clone_border = clone(img, _border = 3);
  assert(img.domain() == img_clone.domain());
  assert(img.domain() == img_clone_border.domain());
  assert(img_clone_border.border() == 3);
  for (auto p : img.domain()) { assert(img(p) == img_clone(p)); assert(img(p) == img_clone_border(p)); }

  {  // subimage
    vint2 p1(10, 10), p2(12, 15);
    auto sub = img | box2d(p1, p2);
    assert(&sub(0, 0) == &img(p1));
    assert(sub.nrows() == (p2[0] - p1[0] + 1));
    assert(sub.ncols() == (p2[1] - p1[1] + 1));
    fill(sub, 7);  // device write through the view
    assert(img(11, 12) == 7 && img(9, 12) == 9 * 200 + 12);
  }
  {  // linear interpolation KAT (tests/imageNd.cc:87-107)
    image2d<vuchar1> test(2, 2, _border = 1);
    vuchar1 m; m[0] = 2;
    vint1 a = cast<vint1>(m);
    assert(a[0] == 2);
    test(0, 0)[0] = 0; test(0, 1)[0] = 10; test(1, 0)[0] = 20; test(1, 1)[0] = 30;
    int v1 = (10 + 20 + 30) / 4.f;
    assert(test.linear_interpolate(vfloat2(0.5, 0.5))[0] == v1);
  }
  {  // move
    image2d<int> i1(10, 10);
    image2d<int> i2 = std::move(i1);
    assert(i2.has_data());
    assert(!i1.has_data());
  }
}

static void test_pixel_wise() {  // tests/pixel_wise.cc
  image2d<int> img2(10, 10, _border = 1);
  pixel_wise(img2) | [=] VPP_KERNEL(int& p) { p = 42; };
  for (auto p : img2.domain()) assert(img2(p) == 42);
  fill(img2, 0);
  pixel_wise(img2) | [=] VPP_KERNEL(int& p) { p = 43; };
  for (auto p : img2.domain()) assert(img2(p) == 43);

  // on a domain: the kernel receives the coordinates
  image2d<vint2> coords(10, 10);
  pixel_wise(coords.domain(), coords) | [=] VPP_KERNEL(vint2 p, vint2& o) { o = p; };
  for (auto p : coords.domain()) assert(coords(p) == p);

  fill_with_border(img2, 0);
  // row forward (tests/pixel_wise.cc:33-39)
  fill(img2, 1);
  pixel_wise(img2, relative_access(img2))(_left_to_right) | [=] VPP_KERNEL(int& o, relative_access_kernel<int> nbh) { o = o + nbh(0, -1); };
  for (auto p : img2.domain()) assert(img2(p) == p[1] + 1);
  // row backward
  fill(img2, 1);
  pixel_wise(img2, relative_access(img2))(_right_to_left) | [=] VPP_KERNEL(int& o, relative_access_kernel<int> nbh) { o = o + nbh(0, 1); };
  for (auto p : img2.domain()) assert(img2(p) == (img2.ncols() - p[1]));
  // col forward
  fill(img2, 1);
  pixel_wise(img2, relative_access(img2), img2.domain())(_top_to_bottom) |
      [=] VPP_KERNEL(int& o, relative_access_kernel<int> nbh, vint2) { o = o + nbh(-1, 0); };
  for (auto p : img2.domain()) assert(img2(p) == (p[0] + 1));
  // col backward
  fill(img2, 1);
  pixel_wise(img2, relative_access(img2), img2.domain())(_bottom_to_top) |
      [=] VPP_KERNEL(int& o, relative_access_kernel<int> nbh, vint2) { o = o + nbh(1, 0); };
  for (auto p : img2.domain()) assert(img2(p) == (img2.nrows() - p[0]));
  // serial raster order (_no_threads): prefix count over the whole image
  fill_with_border(img2, 0);
  pixel_wise(img2, relative_access(img2), img2.domain())(_no_threads) |
      [=] VPP_KERNEL(int& o, relative_access_kernel<int> nbh, vint2 p) { o = (p[1] == 0 ? (p[0] == 0 ? 0 : nbh(-1, 9)) : nbh(0, -1)) + 1; };
  for (auto p : img2.domain()) assert(img2(p) == p[0] * 10 + p[1] + 1);
  // image construction from a value-returning kernel (tests/pixel_wise.cc:62-64)
  auto img3 = pixel_wise(img2) | [=] VPP_KERNEL(int& x) { return x; };
  for (auto p : img2.domain()) assert(img2(p) == img3(p));

  // the hot loop of benchmarks/image_add.cc:51-57
  image2d<int> A(64, 96), B(64, 96), C(64, 96);
  for (auto p : B.domain()) { B(p) = p[0] * 1000 + p[1]; C(p) = 7 * p[1] - p[0]; }
  pixel_wise(A, B, C) | [=] VPP_KERNEL(int& a, int& b, int& c) { a = b + c; };
  for (auto p : A.domain()) assert(A(p) == B(p) + C(p));
  // the 5x5 box of benchmarks/box_5x5_filter2.cc:71-81 through relative_access
  image2d<int> S(40, 50, _border = 2), D(40, 50);
  for (auto p : S.domain()) S(p) = (p[0] * 31 + p[1] * 17) % 1000;
  fill_border_mirror(S);
  pixel_wise(D, relative_access(S)) | [=] VPP_KERNEL(int& b, relative_access_kernel<int> a) {
    int sum = 0;
    for (int i = -2; i <= 2; i++)
      for (int j = -2; j <= 2; j++) sum += a(i, j);
    b = sum / 25;
  };
  for (int r = 5; r < 35; r++)
    for (int c = 5; c < 45; c++) {
      int sum = 0;
      for (int d = -2; d <= 2; d++)
        for (int e = -2; e <= 2; e++) sum += S(r + d, c + e);
      assert(D(r, c) == sum / 25);
    }
}

struct add_k { VPP_KERNEL void operator()(int& a, const int& b, const int& c) const { a = b + c; } };

// The vectorised lowering of the default traversal (16 bytes per thread and image, written back only when changed) against
// plain host loops: ragged widths, mixed element sizes, a box range, a sub-image (rows not 16-byte aligned: scalar path), the
// same image twice (must stay one memory), read-only inputs keep their host mirror.
static void test_pixel_wise_vectorised() {
  const int sizes[][2] = {{1, 1}, {3, 5}, {17, 33}, {64, 96}, {50, 131}, {9, 260}};
  for (auto& sz : sizes) {
    const int nr = sz[0], nc = sz[1];
    image2d<int> A(nr, nc), B(nr, nc), C(nr, nc);
    for (auto p : B.domain()) { B(p) = p[0] * 1000 + p[1]; C(p) = 7 * p[1] - p[0]; A(p) = -1; }
    pixel_wise(A, B, C) | [=] VPP_KERNEL(int& a, int& b, int& c) { a = b + c; };
    for (auto p : A.domain()) { assert(A(p) == B(p) + C(p)); assert(B(p) == p[0] * 1000 + p[1]); }
    // u8 -> int with the coordinates (16 pixels per thread: 16 B of u8, 64 B of int)
    image2d<unsigned char> U(nr, nc);
    image2d<int> W(nr, nc, _border = 2);
    for (auto p : U.domain()) U(p) = (unsigned char)(p[0] * 7 + p[1] * 3);
    fill_with_border(W, 5);
    pixel_wise(W, U, W.domain()) | [=] VPP_KERNEL(int& w, unsigned char& u, vint2 p) { w = 2 * u + p[0] * 100000 + p[1]; };
    for (auto p : W.domain_with_border())
      assert(W(p) == (W.has(p) ? 2 * (int)U(p) + p[0] * 100000 + p[1] : 5));
    // in place through two ranges naming the same pixels
    pixel_wise(A, A) | [=] VPP_KERNEL(int& x, int& y) { x = x + 1; y = y * 2; };
    for (auto p : A.domain()) assert(A(p) == (B(p) + C(p) + 1) * 2);
    // vint2 (8 bytes) and a value-returning kernel
    image2d<vint2> V2(nr, nc);
    pixel_wise(V2, V2.domain()) | [=] VPP_KERNEL(vint2& v, vint2 p) { v = vint2(p[1], -p[0]); };
    auto S = pixel_wise(V2, B) | [=] VPP_KERNEL(vint2& v, int& b) { return v[0] - v[1] + b; };
    for (auto p : S.domain()) assert(S(p) == p[1] + p[0] + B(p));
    if (nr > 4 && nc > 9) {  // sub-image: unaligned rows
      auto sub = A | box2d(vint2(1, 3), vint2(nr - 2, nc - 4));
      auto subB = B | box2d(vint2(1, 3), vint2(nr - 2, nc - 4));
      pixel_wise(sub, subB) | [=] VPP_KERNEL(int& a, int& b) { a = -b; };
      for (auto p : A.domain()) {
        const bool in = p[0] >= 1 && p[0] <= nr - 2 && p[1] >= 3 && p[1] <= nc - 4;
        assert(A(p) == (in ? -B(p) : (B(p) + C(p) + 1) * 2));
      }
    }
  }
  // a kernel that takes its inputs by const reference does not invalidate their host mirror (no download on the next host read)
  image2d<int> A(8, 40), B(8, 40), C(8, 40);
  for (auto p : B.domain()) { B(p) = p[1]; C(p) = p[0]; }
  pixel_wise(A, B, C) | add_k();
  for (auto p : A.domain()) assert(A(p) == p[0] + p[1]);
}

// block_wise with a device callback: one launch for all blocks (block_view<V>), same tiling as the host form
static void test_block_wise_device() {
  image2d<int> im(10, 23, _border = 1), idx(10, 23);
  fill_border_with_value(im, 2);
  fill(im, 0);
  block_wise(vint2(3, 4), im, idx, im.domain()) | [=] VPP_KERNEL(block_view<int> b, block_view<int> id, box2d d) {
    for (int r = 0; r < b.nrows(); r++)
      for (int c = 0; c < b.ncols(); c++) { b(r, c) = 1 + b.nrows() * 10 + b.ncols(); id(r, c) = d.p1()[0] * 100 + d.p1()[1]; }
  };
  for (auto p : im.domain_with_border()) {
    if (!im.has(p)) { assert(im(p) == 2); continue; }
    const int br = p[0] / 3, bc = p[1] / 4;
    const int h = std::min(3, 10 - br * 3), w = std::min(4, 23 - bc * 4);
    assert(im(p) == 1 + h * 10 + w);
    assert(idx(p) == br * 3 * 100 + bc * 4);
  }
  // ordered traversal: a running counter across the blocks, raster order and its reverse
  image2d<int> cnt(1, 1), img(4, 6);
  for (int rev = 0; rev < 2; rev++) {
    fill(cnt, 0);
    const block_view<int> counter{(unsigned char*)cnt.device_write()->base, cnt.device_write()->pitch, 1, 1, vint2(0, 0)};
    auto body = [=] VPP_KERNEL(block_view<int> b) {
      const int k = counter(0, 0)++;
      for (int r = 0; r < b.nrows(); r++)
        for (int c = 0; c < b.ncols(); c++) b(r, c) = k;
    };
    if (rev) block_wise(vint2(2, 2), img)(_bottom_to_top, _right_to_left) | body;
    else block_wise(vint2(2, 2), img)(_no_threads) | body;
    for (auto p : img.domain()) {
      const int k = (p[0] / 2) * 3 + p[1] / 2;
      assert(img(p) == (rev ? 5 - k : k));
    }
  }
}

static void test_block_wise() {  // tests/block_wise.cc
  image2d<int> img(4, 4);
  vint2 b(2, 2);
  auto test_dependency = [&](int* ref_data, auto dep, int dim) {
    image2d<int> ref(img.domain(), _data = (int*)ref_data, _pitch = 4 * sizeof(int));
    int cols[2] = {1, 1};
    block_wise(b, img, img, img.domain())(dep) | [&](image2d<int> I, image2d<int> J, box2d d) {
      int& cpt = cols[d.p1()[dim] / 2];
      fill(I, cpt);
      cpt++;
    };
    assert(equals(ref, img));
  };
  { int ref_data[] = {1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2}; test_dependency(ref_data, _top_to_bottom, 1); }
  fill(img, 9);
  { int ref_data[] = {2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1}; test_dependency(ref_data, _bottom_to_top, 1); }
  fill(img, 9);
  { int ref_data[] = {1, 1, 2, 2, 1, 1, 2, 2, 1, 1, 2, 2, 1, 1, 2, 2}; test_dependency(ref_data, _left_to_right, 0); }
  fill(img, 9);
  { int ref_data[] = {2, 2, 1, 1, 2, 2, 1, 1, 2, 2, 1, 1, 2, 2, 1, 1}; test_dependency(ref_data, _right_to_left, 0); }
  {  // blocks cover the whole image and never touch the border (tests/block_wise.cc:98-112)
    image2d<int> im(10, 10, _border = 1);
    fill_border_with_value(im, 2);
    fill(im, 0);
    block_wise(vint2(3, 3), im) | [](image2d<int> si) { fill(si, 1); };
    for (auto p : im.domain_with_border()) {
      if (im.has(p)) assert(im(p) == 1);
      else assert(im(p) == 2);
    }
    int rows = 0;
    row_wise(im) | [&](image2d<int> row) { assert(row.nrows() == 1 && row.ncols() == 10); rows++; };
    assert(rows == 10);
  }
}

static void test_fill_border_sum() {  // tests/fill.cc, tests/border.cc, tests/sum.cc
  imageNd<int, 2> img({100, 200});
  fill(img, 42);
  for (auto& v : img) assert(v == 42);

  image2d<int> img1(5, 10, _border = 2, _aligned = 16);
  fill_with_border(img1, 42);
  for (auto p : img1.domain_with_border()) assert(img1(p) == 42);
  fill(img1, 5);
  fill_border_with_value(img1, 6);
  for (auto p : img1.domain_with_border()) assert(img1(p) == (img1.domain().has(p) ? 5 : 6));
  fill_with_border(img1, 0);
  pixel_wise(img1.domain(), img1) | [=] VPP_KERNEL(vint2 p, int& v) { v = (p[0] + p[1]) % 10; };
  fill_border_closest(img1);
  for (auto p : img1.domain_with_border()) {
    int cc = std::max(std::min(img1.ncols() - 1, p[1]), 0);
    int cr = std::max(std::min(img1.nrows() - 1, p[0]), 0);
    assert(img1(p) == (cc + cr) % 10);
  }
  fill_border_mirror(img1);
  for (auto p : img1.domain_with_border()) {
    int cr = p[0] < 0 ? -p[0] - 1 : (p[0] >= 5 ? 9 - p[0] : p[0]);
    int cc = p[1] < 0 ? -p[1] - 1 : (p[1] >= 10 ? 19 - p[1] : p[1]);
    assert(img1(p) == (cc + cr) % 10);
  }
  image2d<char> ci(100, 200);
  int s = 0;
  char k = 0;
  for (char& c : ci) { c = k++; s += c; }
  assert(sum(ci) == s);
}

int main() {
  vppb_check(vppb_init(0));
  test_imageNd(); std::puts("imageNd ok");
  test_pixel_wise(); std::puts("pixel_wise ok");
  test_pixel_wise_vectorised(); std::puts("pixel_wise vectorised ok");
  test_block_wise(); std::puts("block_wise ok");
  test_block_wise_device(); std::puts("block_wise device ok");
  test_fill_border_sum(); std::puts("fill/border/sum ok");
  std::puts("ALL OK");
  return 0;
}
