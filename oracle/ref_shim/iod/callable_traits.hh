#pragma once
#include <iod/symbol.hh>
