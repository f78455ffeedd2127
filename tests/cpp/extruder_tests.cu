// C++ host API: keypoint_container / keypoint_trajectory unit checks and video_extruder_init / video_extruder_update
// (vpp/algorithms/video_extruder.hh) against the tables the REFERENCE's own video_extruder_update produced on the
// committed frames (tests/golden/make_video_extruder_fixture.py).  Exit code 0 = all asserts held.
#undef NDEBUG
#include <cassert>
#include <cstdio>
#include <fstream>
#include <string>
#include <vpp/vpp.hh>
#include <vpp/algorithms/video_extruder.hh>

using namespace vpp;

static void test_trajectory() {  // keypoint_trajectory.hh
  keypoint_trajectory t(4);
  assert(t.alive() && t.size() == 0 && t.start_frame() == 4);
  t.move_to(vfloat2(1, 2));
  t.move_to(vfloat2(3, 4));
  t.move_to(vfloat2(5, 6));
  assert(t.size() == 3 && t.end_frame() == 6);
  assert(t.position()[0] == 5 && t[0][1] == 6 && t[2][0] == 1);
  assert(t.position_at_frame(4)[0] == 1 && t.position_at_frame(6)[1] == 6);
  t.pop_oldest_position();
  assert(t.size() == 2 && t[1][0] == 3);
  t.die();
  assert(!t.alive());
}

static void test_container() {  // keypoint_container.hpp:11-200
  keypoint_container<keypoint<int>, int> c(make_box2d(40, 60));
  assert(c.size() == 0 && !c.has(vint2(3, 3)) && c.index_of(vint2(-10, -10)) == -1 && c.index_of(vint2(49, 69)) == -1);  // border of 10, filled with -1
  c.add(keypoint<int>(vint2(5, 7)));
  c.add(keypoint<int>(vint2(20, 30)));
  c.add(keypoint<int>(vint2(39, 59)));
  assert(c.size() == 3 && c.has(vint2(20, 30)) && c.index_of(vint2(39, 59)) == 2 && c(vint2(5, 7)).age == 1);
  c.move(1, vint2(22, 29));
  assert(c[1].age == 2 && c[1].velocity[0] == 2 && c[1].velocity[1] == -1 && c.index_of(vint2(22, 29)) == 1);
  assert(c.index_of(vint2(20, 30)) == 1);  // the old cell keeps its stale entry until prepare_matching()
  c.remove(0);
  assert(!c[0].alive() && c.size() == 3 && !c.has(vint2(5, 7)));
  c.move(0, vint2(6, 7));  // a dead keypoint that is moved is alive again (age 0 -> 1)
  assert(c[0].alive() && c[0].age == 1);
  c.remove(vint2(6, 7));
  std::vector<int> attr = {100, 101};  // shorter than the container: the third keypoint has no attribute yet
  std::vector<int> dead;
  c.compact();
  assert(c.size() == 2 && c[0].position[0] == 22 && c[1].position[0] == 39 && c.index_of(vint2(22, 29)) == 0 && c.index_of(vint2(39, 59)) == 1);
  c.sync_attributes(attr, -7, dead);
  assert(attr.size() == 2 && attr[0] == 101 && attr[1] == -7 && dead.size() == 1 && dead[0] == 100);
  c.prepare_matching();
  assert(!c.has(vint2(22, 29)));
  c.add(keypoint<int>(vint2(1, 1)));
  c.sync_attributes(attr, 55);  // no compact since prepare_matching: plain resize
  assert(attr.size() == 3 && attr[2] == 55 && attr[0] == 101);
  // the index image is host-side bookkeeping but still an image2d<int>
  assert(c.index2d().border() == 10 && c.index2d().nrows() == 40 && c.index2d()(1, 1) == 2);
}

static std::vector<image2d<unsigned char>> load_frames(const std::string& path, int nf, int nr, int nc) {
  std::ifstream f(path, std::ios::binary);
  assert(f.good());
  std::vector<image2d<unsigned char>> frames;
  for (int k = 0; k < nf; k++) {
    image2d<unsigned char> img(nr, nc, _border = 3);  // fast9 / fast9_score need the radius-3 ring
    for (int r = 0; r < nr; r++) f.read((char*)&img(r, 0), nc);
    fill_border_mirror(img);
    frames.push_back(img);
  }
  return frames;
}

static void run_case(const std::vector<image2d<unsigned char>>& frames, int nframes, int detector_th, const std::string& expected_path) {
  std::ifstream f(expected_path, std::ios::binary);
  assert(f.good());
  std::vector<int> expected;
  f.seekg(0, std::ios::end);
  const size_t bytes = (size_t)f.tellg();
  f.seekg(0);
  expected.assign(bytes / 4, 0);
  f.read((char*)expected.data(), (std::streamsize)bytes);
  const int n = (int)(bytes / 24);

  auto ctx = video_extruder_init(frames[0].domain());
  assert(ctx.frame_id == -1);
  for (int k = 1; k < nframes; k++)
    video_extruder_update(ctx, frames[k - 1], frames[k], _detector_th = detector_th, _keypoint_spacing = 10, _detector_period = 3,
                          _max_trajectory_length = 5, _nscales = 3, _winsize = 9, _propagation = 2);
  assert(ctx.frame_id == nframes - 2);
  std::printf("video_extruder %d frames th %d: %d keypoints (reference: %d)\n", nframes, detector_th, ctx.keypoints.size(), n);
  assert(ctx.keypoints.size() == n && (int)ctx.trajectories.size() == n);
  int tracked = 0, dead = 0, shared = 0;
  for (int i = 0; i < n; i++) {
    const int* e = &expected[6 * i];
    assert(ctx.keypoints[i].position[0] == e[0] && ctx.keypoints[i].position[1] == e[1]);
    assert(ctx.keypoints[i].age == e[2]);
    assert(ctx.trajectories[i].start_frame() == e[3] && ctx.trajectories[i].size() == e[4] && (int)ctx.trajectories[i].alive() == e[5]);
    if (ctx.keypoints[i].alive()) {  // the newest trajectory point is the keypoint itself
      assert(ctx.trajectories[i].position()[0] == (float)e[0] && ctx.trajectories[i].position()[1] == (float)e[1]);
      // the index image is lossy, as in the reference: a second keypoint that lands on the pixel takes the entry over and
      // clears it when it is removed (keypoint_container.hpp:135-142,153-166) - count the entries that still point back
      shared += ctx.keypoints.index_of(ctx.keypoints[i].position) != i;
    }
    tracked += e[2] > 1;
    dead += e[2] == 0;
  }
  assert(tracked > 10);
  std::printf("  %d tracked across frames, %d dead but not yet compacted, %d live keypoints whose index entry points elsewhere\n", tracked, dead, shared);
}

int main(int argc, char** argv) {
  vppb_check(vppb_init(0));
  const std::string gold = argc > 1 ? argv[1] : "tests/golden";
  test_trajectory();
  test_container();
  std::printf("keypoint_trajectory / keypoint_container ok\n");
  auto frames = load_frames(gold + "/video_extruder_frames_9x121x161.u8", 9, 121, 161);
  run_case(frames, 7, 4, gold + "/video_extruder_expected_7f_th4.i32");
  run_case(frames, 9, 5, gold + "/video_extruder_expected_9f_th5.i32");
  std::printf("ALL OK\n");
  return 0;
}
