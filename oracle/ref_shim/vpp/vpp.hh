// Shadow of vpp/vpp.hh for the reference build: the same core headers (taken from /root/reference
// through the second -I), minus draw / colorspace / keypoint-trajectory / patch headers that the hot
// path does not use and that would need a larger Eigen surface.
#pragma once
#include <vpp/core/imageNd.hh>
#include <vpp/core/image2d.hh>
#include <vpp/core/imageNd_iterator.hh>
#include <vpp/core/boxNd_iterator.hh>
#include <vpp/core/vector.hh>
#include <vpp/core/boxNd.hh>
#include <vpp/core/relative_accessor.hh>
#include <vpp/core/pixel_wise.hh>
#include <vpp/core/block_wise.hh>
#include <vpp/core/copy.hh>
#include <vpp/core/clone.hh>
#include <vpp/core/fill.hh>
#include <vpp/core/zero.hh>
#include <vpp/core/cast_to_float.hh>
#include <vpp/core/pyramid.hh>
#include <vpp/core/sum.hh>
