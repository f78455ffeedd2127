// lbp_transform(A, B) (reference: vpp/algorithms/lbp/lbp_transform.hh:7-38) and lbp_hamming_distance (lbp/lbp_distance.hh): bit k of
// B(r, c) = neighbour k of A(r, c) > A(r, c), neighbours in raster order without the centre.  A needs a border >= 1, filled by the
// caller (the reference reads it as it is).  unsigned char images go to the CUDA kernel behind vppb_lbp_u8.
#pragma once
#include <vpp/core/image2d.hh>

namespace vpp {

inline void lbp_transform(const image2d<unsigned char>& A, image2d<unsigned char>& B) {
  vppb_check(vppb_lbp_u8(A.device_read(), B.device_write(), nullptr));
}

inline int lbp_hamming_distance(unsigned char a, unsigned char b) {
  unsigned x = (unsigned)(a ^ b);
  int n = 0;
  for (; x; x &= x - 1) n++;
  return n;
}

}  // namespace vpp
