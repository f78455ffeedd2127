// iod::array_view(size, f): read-only view whose i-th element is f(i) (used by video_extruder.hpp:46-47)
#pragma once
#include <iod/symbol.hh>
namespace iod {
template <typename F>
struct array_view_ {
  int size_; F f;
  int size() const { return size_; }
  auto operator[](int i) const { return f(i); }
};
template <typename F>
array_view_<F> array_view(int size, F f) { return array_view_<F>{size, f}; }
}  // namespace iod
