// box2d: inclusive integer box [p1, p2] with raster iteration (reference: vpp/core/boxNd.hh:11-149).
#pragma once
#include <vpp/core/vector.hh>

namespace vpp {

class border {
 public:
  explicit border(int n) : size_(n) {}
  int size() const { return size_; }
 private:
  int size_;
};

class box2d {
 public:
  typedef vint2 coord_type;
  struct iterator {
    vint2 p; int c0, c1;
    VPP_HD const vint2& operator*() const { return p; }
    iterator& operator++() { next(); return *this; }
    void next() { if (p[1] == c1) { p[1] = c0; p[0]++; } else p[1]++; }
    bool operator!=(const iterator& o) const { return p != o.p; }
    bool operator==(const iterator& o) const { return p == o.p; }
  };
  VPP_HD box2d() : p1_(0, 0), p2_(-1, -1) {}
  VPP_HD box2d(vint2 p1, vint2 p2) : p1_(p1), p2_(p2) {}
  VPP_HD bool has(const vint2& p) const { return p[0] >= p1_[0] && p[0] <= p2_[0] && p[1] >= p1_[1] && p[1] <= p2_[1]; }
  VPP_HD const vint2& p1() const { return p1_; }
  VPP_HD const vint2& p2() const { return p2_; }
  VPP_HD const vint2& first_point_coordinates() const { return p1_; }
  VPP_HD const vint2& last_point_coordinates() const { return p2_; }
  VPP_HD int size(int d) const { return p2_[d] - p1_[d] + 1; }
  VPP_HD int nrows() const { return size(0); }
  VPP_HD int ncols() const { return size(1); }
  iterator begin() const { return iterator{p1_, p1_[1], p2_[1]}; }
  iterator end() const { return iterator{vint2(p2_[0] + 1, p1_[1]), p1_[1], p2_[1]}; }
 private:
  vint2 p1_, p2_;
};

inline bool operator==(const box2d& a, const box2d& b) { return a.p1() == b.p1() && a.p2() == b.p2(); }
inline bool operator!=(const box2d& a, const box2d& b) { return !(a == b); }
inline box2d make_box2d(int nr, int nc) { return box2d(vint2(0, 0), vint2(nr - 1, nc - 1)); }
inline box2d operator-(const box2d& b, const border& bd) { return box2d(vint2(b.p1()[0] + bd.size(), b.p1()[1] + bd.size()), vint2(b.p2()[0] - bd.size(), b.p2()[1] - bd.size())); }
inline box2d operator+(const box2d& b, const border& bd) { return box2d(vint2(b.p1()[0] - bd.size(), b.p1()[1] - bd.size()), vint2(b.p2()[0] + bd.size(), b.p2()[1] + bd.size())); }
inline box2d operator|(const box2d&, const box2d& b) { return b; }

}  // namespace vpp
