// Shadow of vpp/algorithms/epipolar_geometry.hh for the reference build: semi_dense_optical_flow.hpp
// calls epipole_right(F) unconditionally but only uses the result on the compile-time-off
// _epipolar_flow path; the real header needs Eigen::EigenSolver.
#pragma once
#include <vpp/vpp.hh>
namespace vpp { inline vfloat2 epipole_right(const Eigen::Matrix3f&) { return vfloat2(0.f, 0.f); } }
