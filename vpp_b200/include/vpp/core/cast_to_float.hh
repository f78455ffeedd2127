// cast_to_float<V>: float for scalars, vector<float, N> for vector<T, N> (reference: vpp/core/cast_to_float.hh:9-22).
#pragma once
#include <vpp/core/vector.hh>

namespace vpp {

template <typename V> struct cast_to_float_ { typedef float ret; };
template <typename T, unsigned N> struct cast_to_float_<vector<T, N>> { typedef vector<float, N> ret; };
template <typename V> using cast_to_float = typename cast_to_float_<V>::ret;

}  // namespace vpp
