#pragma once
#include <vpp/core/image2d.hh>
