// Shadow of vpp/algorithms/optical_flow.hh for the reference build: only the semi-dense flow (the dense /
// sparse variants pull OpenCV and are not on the hot path).
#pragma once
#include <vpp/vpp.hh>
#include <vpp/algorithms/optical_flow/semi_dense_optical_flow.hpp>
