// keypoint<C> and keypoint_container<P, F>: the set of tracked keypoints with a 2-D index image
// (reference: vpp/core/keypoint_container.hh:13-90, keypoint_container.hpp:11-200).  Host-side bookkeeping: the
// index image lives on the host mirror of an image2d<int> (border 10 as in the reference) and never touches the GPU.
//
// Semantics kept from the reference, because video_extruder depends on them:
//  - remove(i) only marks the keypoint dead (age 0); it stays in the container until compact();
//  - move(i, p) increments the age, so a dead keypoint that is moved again is alive again (keypoint_container.hpp:136-149);
//  - compact() keeps the live keypoints in order and records old -> new indices, which sync_attributes() uses to carry
//    a parallel attribute vector (the trajectories) along; entries beyond the attribute vector get `new_value`.
#pragma once
#include <cassert>
#include <vector>

#include <vpp/core/image2d.hh>

namespace vpp {

template <typename C>
struct keypoint {
  keypoint() : age(0) {}
  keypoint(vector<C, 2> pos) : position(pos), velocity(0, 0), age(1) {}

  vector<C, 2> position;
  vector<C, 2> velocity;
  int age;

  void die() { age = 0; }
  bool alive() const { return age > 0; }
};

template <typename P, typename F>
struct keypoint_container {
  typedef P keypoint_type;
  typedef F feature_type;
  typedef std::vector<P> keypoint_vector_type;
  typedef std::vector<F> feature_vector_type;

  keypoint_container(const box2d& d) : index2d_(d, _border = 10), compact_has_run_(false) {  // keypoint_container.hpp:11-19
    index2d_.host_fill_with_border(-1);
    keypoint_vector_.reserve((size_t)(d.nrows() * d.ncols()) / 10);
    feature_vector_.reserve((size_t)(d.nrows() * d.ncols()) / 10);
  }

  // drop the dead keypoints, keeping the order of the live ones (keypoint_container.hpp:22-53)
  void compact() {
    compact_has_run_ = true;
    matches_.assign(keypoint_vector_.size(), -1);
    size_t out = 0;
    for (size_t i = 0; i < keypoint_vector_.size(); i++) {
      if (!keypoint_vector_[i].alive()) continue;
      keypoint_vector_[out] = keypoint_vector_[i];
      feature_vector_[out] = feature_vector_[i];
      index2d_(cast<vint2>(keypoint_vector_[i].position)) = (int)out;
      matches_[i] = (int)out;
      out++;
    }
    keypoint_vector_.resize(out);
    feature_vector_.resize(out);
  }

  void prepare_matching() {  // keypoint_container.hpp:55-62
    compact_has_run_ = false;
    index2d_.host_fill_with_border(-1);
    std::fill(matches_.begin(), matches_.end(), -1);
  }

  struct no_op {
    template <typename T>
    void operator()(T&) {}
  };

  // Bring a parallel attribute vector in line with the container (keypoint_container.hpp:64-101): after compact(),
  // attribute i follows keypoint i to its new index (die_fun gets the attributes of removed keypoints); new keypoints
  // receive new_value.
  template <typename T, typename D = no_op>
  void sync_attributes(T& v, typename T::value_type new_value = typename T::value_type(), D die_fun = D()) const {
    const size_t nparts = keypoint_vector_.size();
    if (!compact_has_run_) { v.resize(nparts, new_value); return; }
    T tmp(nparts, new_value);
    for (size_t i = 0; i < matches_.size(); i++) {
      const int ni = matches_[i];
      if (ni >= 0) {
        if (i < v.size()) tmp[ni] = std::move(v[i]);
      } else if (i < v.size()) {
        die_fun(v[i]);
      }
    }
    v.swap(tmp);
  }
  template <typename T, typename U>
  void sync_attributes(T& container, typename T::value_type new_value, std::vector<U>& dead_vector) const {
    sync_attributes(container, new_value, [&dead_vector](typename T::value_type& x) { dead_vector.push_back(std::move(x)); });
  }

  void add(const keypoint_type& p, const feature_type& f = feature_type()) {  // keypoint_container.hpp:114-121
    index2d_(cast<vint2>(p.position)) = (int)keypoint_vector_.size();
    keypoint_vector_.push_back(p);
    feature_vector_.push_back(f);
  }
  void add(const vfloat2& p) { add(keypoint_type(cast<decltype(keypoint_type().position)>(p))); }

  void remove(int i) {  // keypoint_container.hpp:135-142
    assert(i >= 0 && i < size());
    keypoint_vector_[i].die();
    int& index = index2d_(cast<vint2>(keypoint_vector_[i].position));
    if (index == i) index = -1;
  }
  void remove(vint2 position) { assert(has(position)); remove(index2d_(position)); }

  template <typename T>
  void move(int i, T position) {  // keypoint_container.hpp:153-166
    assert(i >= 0 && i < size());
    keypoint_type& kp = keypoint_vector_[i];
    kp.velocity = position - kp.position;
    kp.position = position;
    kp.age++;
    index2d_(cast<vint2>(kp.position)) = i;
  }
  void update(unsigned i, const keypoint_type& p, const feature_type& f) {
    assert((int)i < size());
    keypoint_vector_[i] = p;
    feature_vector_[i] = f;
    index2d_(cast<vint2>(p.position)) = (int)i;
  }
  void update_index(unsigned i, const vint2& p) { index2d_(p) = (int)i; }

  keypoint_vector_type& keypoints() { return keypoint_vector_; }
  const keypoint_vector_type& keypoints() const { return keypoint_vector_; }
  image2d<int>& index2d() { return index2d_; }
  const image2d<int>& index2d() const { return index2d_; }
  int index_of(const vint2& p) const { return index2d_(p); }
  bool has(vint2 p) const { return index2d_(p) >= 0; }
  int size() const { return (int)keypoint_vector_.size(); }

  keypoint_type& operator[](unsigned i) { return keypoint_vector_[i]; }
  const keypoint_type& operator[](unsigned i) const { return keypoint_vector_[i]; }
  keypoint_type& operator()(vint2 p) { return keypoint_vector_[index2d_(p)]; }
  const keypoint_type& operator()(vint2 p) const { return keypoint_vector_[index2d_(p)]; }

 private:
  std::vector<int> matches_;
  image2d<int> index2d_;
  keypoint_vector_type keypoint_vector_;
  feature_vector_type feature_vector_;
  bool compact_has_run_;
};

}  // namespace vpp
