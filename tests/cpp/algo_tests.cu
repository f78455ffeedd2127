// C++ host API tests for the algorithms: the reference's LK integration test (tests/pyrlk.cc:14-50)
// on the committed fixture, plus FAST9 / Scharr / pyramid sanity checks through the reference's names.
#undef NDEBUG
#include <cassert>
#include <cstdio>
#include <fstream>
#include <vpp/vpp.hh>
#include <vpp/algorithms/fast_detector/fast.hh>
#include <vpp/algorithms/filters/scharr.hh>
#include <vpp/algorithms/lucas_kanade.hh>
#include <vpp/algorithms/optical_flow.hh>
#include <vpp/algorithms/pyrlk/pyrlk_match.hh>
#include <vpp/algorithms/pyrlk/lk.hh>
#include <vpp/algorithms/lbp/lbp_transform.hh>

using namespace vpp;

static image2d<uint8_t> load_u8(const std::string& path, int nr, int nc) {
  image2d<uint8_t> img(nr, nc);
  std::ifstream f(path, std::ios::binary);
  assert(f.good());
  for (int r = 0; r < nr; r++) f.read((char*)&img(r, 0), nc);
  return img;
}

struct kp_t { vfloat2 position; int age = 1; bool alive() const { return age > 0; } };
struct kp_container {
  std::vector<kp_t> k;
  int size() const { return (int)k.size(); }
  kp_t& operator[](int i) { return k[i]; }
  void remove(int i) { k[i].age = 0; }
  void move(int i, vfloat2 p) { k[i].position = p; k[i].age++; }
};

int main(int argc, char** argv) {
  vppb_check(vppb_init(0));
  std::string gold = argc > 1 ? argv[1] : "tests/golden";

  {  // tests/pyrlk.cc:14-50
    image2d<uint8_t> i1 = load_u8(gold + "/pyrlk_i1_100x100.u8", 100, 100), i2 = load_u8(gold + "/pyrlk_i2_100x100.u8", 100, 100);
    std::vector<vfloat2> keypoints;
    keypoints.push_back(vfloat2(50, 50));
    int calls = 0;
    lucas_kanade(i1, i2, _keypoints = keypoints, _niterations = 50, _winsize = 5, _min_ev = 0.001, _delta = 0.01, _nscales = 2,
                 _flow = [&](vfloat2 p, vfloat2 f, float d) {
                   assert(p == vfloat2(50.f, 50.f));
                   assert((f - vfloat2(2.f, 2.f)).norm() < 0.05);
                   std::printf("lucas_kanade flow (%f, %f) dist %f\n", f[0], f[1], d);
                   calls++;
                 });
    assert(calls == 1);

    // pyrlk_match on the same scene (float gradient, as benchmarks/pyrlk_opencv_comparison.cc:49-60 builds them)
    pyramid2d<uint8_t> prev(i1, 2, 2, _border = 4), next(i2, 2, 2, _border = 4);
    pyramid2d<vfloat2> grad(i1.domain(), 2, 2, _border = 4);
    scharr(prev[0], grad[0]);
    grad.propagate_level0();
    kp_container kc;
    kc.k.resize(2);
    kc.k[0].position = vfloat2(50, 50);
    kc.k[1].position = vfloat2(10, 10);  // flat area: rejected by the min eigenvalue test -> removed
    pyrlk_match(prev, grad, next, kc, lk_match_point_square_win<5>(), 0.001f, 1e9f, 50, 0.01f);
    assert(kc.k[0].alive() && (kc.k[0].position - vfloat2(52.f, 52.f)).norm() < 0.1);
    assert(!kc.k[1].alive());
  }
  {  // FAST9: a bright square on black has exactly its 4 corners as local maxima candidates; border rule
    image2d<uint8_t> img(64, 64, _border = 3);
    fill(img, 0);
    fill(img, (uint8_t)255, box2d(vint2(20, 20), vint2(40, 40)));
    fill_border_mirror(img);
    std::vector<int> scores;
    auto kps = fast9(img, 20, _scores = &scores, _ring = 1);
    assert(!kps.empty() && kps.size() == scores.size());
    for (size_t i = 1; i < kps.size(); i++) assert(kps[i - 1][0] < kps[i][0] || (kps[i - 1][0] == kps[i][0] && kps[i - 1][1] < kps[i][1]));
    bool has_corner = false;
    for (auto p : kps) has_corner |= (p == vint2(20, 20));
    assert(has_corner);
    auto lm = fast9(img, 20, _local_maxima, _ring = 1);
    assert(!lm.empty() && lm.size() <= kps.size());
    auto bw = fast9(img, 20, _blockwise, _block_size = 8, _ring = 1);
    assert(!bw.empty() && bw.size() <= kps.size());
    assert(fast9_score(img, 20, vint2(20, 20)) > 0);
    bool threw = false;
    try { fast9(image2d<uint8_t>(10, 10, _border = 2), 10); } catch (std::runtime_error&) { threw = true; }  // fast.hpp:937-938
    assert(threw);
  }
  {  // pyramid level sizes (pyramid.hh:140) and Scharr of a ramp
    pyramid2d<uint8_t> p(make_box2d(1080, 1920), 3, 2, _border = 2);
    assert(p[1].nrows() == 541 && p[1].ncols() == 961 && p[2].nrows() == 271 && p[2].ncols() == 481);
    image2d<uint8_t> ramp(32, 32, _border = 1);
    pixel_wise(ramp.domain_with_border(), ramp) | [=] VPP_KERNEL(vint2 q, uint8_t& v) { v = (uint8_t)(4 * q[1] + 8); };
    image2d<vint2> g(32, 32);
    scharr(ramp, g);
    for (auto q : g.domain()) assert(g(q) == vint2(0, 4));  // d/drow = 0, d/dcol = (3+10+3)*8/32
  }
  {  // semi_dense_optical_flow: a textured frame shifted by (3,-2) -> most cells report that flow
    image2d<uint8_t> f1(121, 161), f2(121, 161);
    for (auto p : f1.domain()) {
      auto tex = [](int r, int c) { return (uint8_t)(128 + 60 * std::sin(r * 0.37) * std::cos(c * 0.29) + 50 * std::sin((r + 2 * c) * 0.11)); };
      f1(p) = tex(p[0], p[1]);
      f2(p) = tex(p[0] - 3, p[1] + 2);
    }
    image2d<uint8_t> g = clone(f1, _border = 3);
    fill_border_mirror(g);
    auto kps = fast9(g, 5, _blockwise, _block_size = 6);
    assert(kps.size() > 50);
    int calls = 0, good = 0, last = -1;
    semi_dense_optical_flow(kps, [&](int i, vint2 pos, int d) { assert(i > last); last = i; calls++; vint2 f = pos - kps[i]; good += (f == vint2(3, -2)); (void)d; },
                            f1, f2, _winsize = 9, _nscales = 3, _patchsize = 5, _propagation = 2);
    std::printf("semi_dense_optical_flow: %d callbacks, %d with flow (3,-2)\n", calls, good);
    assert(calls > 50 && good * 10 >= calls * 7);
  }
  {  // tests/lbp.cc:9-41 (the reference's own test, verbatim values)
    image2d<unsigned char> V(3, 3, _border = 1);
    image2d<unsigned char> lbp(3, 3);
    V(1, 1) = 1;
    V(0, 0) = 0; V(0, 1) = 2; V(0, 2) = 2;
    V(1, 0) = 2; V(1, 2) = 0;
    V(2, 0) = 2; V(2, 1) = 0; V(2, 2) = 2;
    unsigned char x = 0b10101110;
    lbp_transform(V, lbp);
    assert(lbp(1, 1) == x);
    assert(lbp_hamming_distance(0b01010101, 0b01010101) == 0);
    assert(lbp_hamming_distance(0b11010101, 0b01010101) == 1);
    assert(lbp_hamming_distance(0b11111111, 0b00000000) == 8);
  }
  {  // local_maxima_filter (fast.hpp:555-575), serial in-place semantics: on a row that decreases to the right every pixel's left
     // neighbour is larger - but it has been zeroed before the pixel is looked at, so every second pixel survives
    image2d<int> S(3, 8, _border = 1);
    fill_with_border(S, 0);
    for (int c = 0; c < 8; c++) S(1, c) = 100 - c;
    local_maxima_filter(S, 3);
    for (int c = 0; c < 8; c++) assert(S(1, c) == ((c % 2 == 0) ? 100 - c : 0));
    for (int c = 0; c < 8; c++) assert(S(0, c) == 0 && S(2, c) == 0);
  }
  {  // fast_detector9_blockwise_rank (fast.hpp:801-886): ranks per block ascend, scores descend, every point is a detected corner
    image2d<uint8_t> img(96, 128, _border = 3);
    fill(img, 40);
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 8; j++) fill(img, (uint8_t)(120 + 10 * ((i + j) % 5)), box2d(vint2(16 * i + 3, 16 * j + 3), vint2(16 * i + 11, 16 * j + 10)));  // 48 bright rectangles
    fill_border_mirror(img);
    std::vector<int> sc;
    auto r3 = fast_detector9_blockwise_rank(img, 20, 16, 3, image2d<unsigned char>(), &sc, 1);
    auto all = fast9(img, 20, _ring = 1);
    assert(r3.size() > 10 && sc.size() == r3.size());
    for (size_t i = 0; i < r3.size(); i++) {
      assert(r3[i][2] >= 0 && r3[i][2] < 3 && sc[i] > 0);
      bool found = false;
      for (auto& k : all) found = found || (k[0] == r3[i][0] && k[1] == r3[i][1]);
      assert(found);
      if (r3[i][2] > 0) assert(i > 0 && r3[i - 1][2] == r3[i][2] - 1 && sc[i - 1] >= sc[i]);
    }
  }
  {  // oriented_lk_match_point_square_win (lk.hh:180-317): the tests/pyrlk.cc scene moved by (2, 2); axis-aligned directions
    image2d<uint8_t> i1 = load_u8(gold + "/pyrlk_i1_100x100.u8", 100, 100), i2 = load_u8(gold + "/pyrlk_i2_100x100.u8", 100, 100);
    image2d<uint8_t> a = clone(i1, _border = 3), b = clone(i2, _border = 3);
    fill_border_mirror(a); fill_border_mirror(b);
    image2d<vfloat2> g(100, 100, _border = 3);
    scharr(a, g);
    fill_border_mirror(g);
    oriented_lk_match_point_square_win<9> matcher;
    auto m = matcher(vfloat2(50, 50), vfloat2(1.5f, 1.5f), a, b, g, 0.0001f, 30, 0.01f, 1.f, vfloat2(0, 1), vfloat2(0, 1));
    std::printf("oriented LK: flow (%f, %f) err %f\n", m.first[0], m.first[1], m.second);
    assert(std::fabs(m.first[0] - 2.f) < 0.3f && std::fabs(m.first[1] - 2.f) < 0.3f && m.second < 1.f);
    auto out = matcher(vfloat2(50, 50), vfloat2(60.f, 0.f), a, b, g, 0.0001f, 30, 0.01f, 100.f, vfloat2(0, 1), vfloat2(0, 1));
    assert(out.second > 1e30f);  // driven out of the domain shrunk by 3: ((0,0), FLT_MAX)
  }
  std::puts("ALL OK");
  return 0;
}
