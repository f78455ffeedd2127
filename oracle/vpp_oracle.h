/*
 * vpp_oracle.h — CPU restatement of the Video++ (matt-42/vpp) dense-pixel hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the checker the CUDA path is compared against; nothing in
 * the product (vpp_b200/) may include, link or call it.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Each function restates, in plain C, the arithmetic of the reference template it cites
 * (file:line relative to the reference tree, commit 8413773a).  The reference itself needs
 * Eigen3 and iod, neither of which is vendored nor installed here; oracle/ref_shim builds the
 * reference's own headers against a minimal stand-in for those two libraries (oracle/_ref) and
 * tests/test_oracle_vs_ref.py pins this restatement to it.
 *
 * Documented deviations from the reference (all remove undefined behaviour):
 *   - the low-pass temp image border read by subsample2 for even parent sizes is uninitialised
 *     in the reference (pyramid.hh:179-181); here out-of-domain rows/cols of the temp are defined
 *     by mirroring (the value fill_border_mirror would have produced).
 *   - LK as[]/gs[] entries of out-of-domain window pixels are zero instead of uninitialised.
 *
 * Build: parity  gcc -O2 -ffp-contract=off -fno-fast-math            (serial, deterministic)
 *        timing  gcc -O3 -march=native -fopenmp -DNDEBUG -ffp-contract=off  (benchmarks/CMakeLists.txt:10,18)
 */
#ifndef VPP_ORACLE_H_
#define VPP_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* host image descriptor, same meaning as vppb_img but `base` is a HOST pointer to pixel (0,0) */
typedef struct vo_img {
  unsigned char* base;
  int32_t nrows, ncols, pitch, border, elem;
} vo_img;

typedef struct vo_int2 { int32_t r, c; } vo_int2;
typedef struct vo_float2 { float r, c; } vo_float2;

/* imageNd.hpp:151-196 */
int vo_layout(int nrows, int ncols, int elem, int border, int align, int* pitch, int64_t* total, int64_t* origin);

/* pixel_wise.hpp:69-165 with the kernels of benchmarks/image_add.cc:51-57, fill.hh, copy.hh, sum.hh */
void vo_pw_add_i32(const vo_img* a, const vo_img* b, const vo_img* c);
void vo_fill(const vo_img* img, const void* value, int with_border);
void vo_copy(const vo_img* src, const vo_img* dst, int with_border);
/* rgb_to_graylevel (colorspace_conversions.hh:10-47): in elem 3 or 4 (u8 channels), out u8 */
void vo_rgb_to_graylevel(const vo_img* in, const vo_img* out);
void vo_fill_border_value(const vo_img* img, const void* value);
void vo_fill_border_mirror(const vo_img* img);
void vo_fill_border_closest(const vo_img* img);
int64_t vo_sum_i32(const vo_img* img, int is_signed);

/* benchmarks/box_5x5_filter2.cc:71-81, examples/box_filter.cc:23-32; channels = 1 or 3 (u8), or int32 */
void vo_box5x5_u8(const vo_img* in, const vo_img* out, int channels);
void vo_box5x5_i32(const vo_img* in, const vo_img* out);

/* scharr.hh:46-87; out elem 8 bytes: as_float 0 -> vint2 (float quotient truncated), 1 -> vfloat2 */
void vo_scharr_u8(const vo_img* in, const vo_img* out, int as_float);
/* pyramid.hh:12-59 + :62-81: out = subsample2(antialiasing_lowpass_filter(in)); kind 0 u8, 1 vint2, 2 vfloat2 */
void vo_lowpass_sub2(const vo_img* in, const vo_img* out, int kind);
/* full-resolution low-pass only (pyramid.hh:12-59), out same domain as in */
void vo_lowpass(const vo_img* in, const vo_img* out, int kind);

/* fast.hpp:253-508 (ring = 0: as implemented, a4/a12 taken on row r-3; ring = 1: true ring == fast.hpp:79-112),
 * mask semantics fast.hpp:310-317; modes fast.hpp:663-673, 889-928, 745-799.
 * Returns the number of keypoints written (raster order), or -(needed) if capacity is too small. */
int vo_fast9_u8(const vo_img* img, int th, const vo_img* mask, int mode, int block_size, int ring, vo_int2* kps,
                int32_t* scores, int capacity);
int vo_fast9_score(const vo_img* img, int th, int r, int c); /* fast.hpp:36-77 */

/* imageNd.hpp:280-300 on u8 (returns truncated uchar) / vint2 / vfloat2 */
int vo_interp_u8(const vo_img* img, float pr, float pc);

typedef struct vo_lk_params {
  int32_t nlevels, min_scale, winsize, max_iter, grad_is_float, err_mode, gate_on_max_err;
  float min_ev, delta, max_err, factor, pred_div;
} vo_lk_params;
/* lucas_kanade.hpp:12-184 / lk.hh:42-175 / pyrlk_match.hh:15-55 (see include/vppb.h for the parameter meaning) */
void vo_lk_match_u8(const vo_img* prev, const vo_img* next, const vo_img* grad, const vo_lk_params* p,
                    const vo_float2* kps, const vo_float2* prediction, int n, vo_float2* flow_out, float* err_out);

/* semi_dense_optical_flow.hpp:46-214 + gradient_descent.hh:10-89, serial semantics (oracle/vpp_oracle_sdof.c) */
/* kitti::flow_error_stats, evaluation/utils/kitti.hh:75-134 */
int vo_flow_error_stats(const vo_img* flow, const vo_img* ref, float* out, const vo_img* errors_map);
void vo_semi_dense_flow(const vo_img* i1, const vo_img* i2, const vo_int2* kps, int n, int winsize, int nscales, int min_scale,
                        int propagation, int patchsize, vo_int2* out_pos, int32_t* out_dist, unsigned char* out_valid);

/* lbp_transform.hh:7-38; fast.hpp:555-575 (serial order, in place); fast.hpp:710-740 + 801-886; lk.hh:180-317 */
void vo_lbp_u8(const vo_img* in, const vo_img* out);
void vo_local_maxima_filter(const vo_img* img);
int vo_fast9_blockwise_rank(const vo_img* img, int th, const vo_img* mask, int block_size, int max_points, int ring, int32_t* kps3, int32_t* scores,
                            int capacity);
void vo_lk_match_oriented_u8(const vo_img* A, const vo_img* B, const vo_img* Ag, int grad_is_float, int winsize, float min_ev_th, int max_iter,
                             float delta, float max_step_norm, const vo_float2* kps, const vo_float2* prediction, const vo_float2* dir1,
                             const vo_float2* dir2, int n, vo_float2* flow_out, float* err_out);

int vo_num_threads(void);
void vo_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
