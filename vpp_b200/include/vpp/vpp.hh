// Umbrella header of the B200-native dense-pixel path (reference: vpp/vpp.hh:3-37, hot-path subset).
#pragma once
#include <vpp/core/boxNd.hh>
#include <vpp/core/clone.hh>
#include <vpp/core/colorspace_conversions.hh>
#include <vpp/core/copy.hh>
#include <vpp/core/fill.hh>
#include <vpp/core/image2d.hh>
#include <vpp/core/imageNd.hh>
#include <vpp/core/pixel_wise.hh>
#include <vpp/core/block_wise.hh>
#include <vpp/core/pyramid.hh>
#include <vpp/core/relative_accessor.hh>
#include <vpp/core/sum.hh>
#include <vpp/core/symbols.hh>
#include <vpp/core/vector.hh>
