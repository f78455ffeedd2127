// keypoint_trajectory: the positions a tracked keypoint went through, newest first
// (reference: vpp/core/keypoint_trajectory.hh:11-73).  Host-side bookkeeping of video_extruder.
#pragma once
#include <cassert>
#include <deque>

#include <vpp/core/vector.hh>

namespace vpp {

struct keypoint_trajectory {
  keypoint_trajectory() : start_frame_(0), alive_(true) {}
  keypoint_trajectory(int frame_cpt) : start_frame_(frame_cpt), alive_(true) {}

  void die() { alive_ = false; }
  bool alive() const { return alive_; }

  int size() const { return (int)history_.size(); }
  vfloat2 position() const { assert(size() > 0); return history_.front(); }  // the current (newest) position
  vfloat2 operator[](unsigned i) const { return history_[i]; }               // i frames ago
  vfloat2 position_at_frame(int frame_cpt) const {
    assert(frame_cpt >= start_frame_ && frame_cpt < start_frame_ + size());
    return history_[size() - 1 - (frame_cpt - start_frame_)];
  }
  void move_to(vfloat2 p) { history_.push_front(p); }
  void pop_oldest_position() { history_.pop_back(); }

  const std::deque<vfloat2>& positions() const { return history_; }
  int start_frame() const { return start_frame_; }
  int end_frame() const { return start_frame_ + size() - 1; }

 private:
  int start_frame_;
  bool alive_;
  std::deque<vfloat2> history_;
};

}  // namespace vpp
