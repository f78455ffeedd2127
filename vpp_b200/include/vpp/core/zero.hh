// zero<V> (reference: vpp/core/zero.hh:7-26) lives with the pixel types.
#pragma once
#include <vpp/core/vector.hh>
