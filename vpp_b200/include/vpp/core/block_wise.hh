#pragma once
#include <vpp/core/pixel_wise.hh>
