// Pixel value types: vector<T,N> (vuchar3, vint2, vfloat2, ...), cast<U>, plus_promotion.
// Reference: vpp/core/vector.hh:10-109 (there an alias of Eigen::Matrix<T,N,1>); here a POD usable from
// host and device code with the same arithmetic (component-wise ops, C++ truncating casts).
#pragma once
#include <cmath>
#include <type_traits>

#if defined(__CUDACC__)
#define VPP_HD __host__ __device__
#else
#define VPP_HD
#endif
// annotation a pixel_wise kernel (lambda or functor call operator) needs to run on the GPU
#define VPP_KERNEL VPP_HD

namespace vpp {

template <typename T, unsigned N>
struct vector {
  typedef T Scalar;
  enum { SizeAtCompileTime = N };
  T v[N];

  VPP_HD vector() {}
  template <typename A, typename B>
  VPP_HD vector(A a, B b) { static_assert(N == 2, "2 components"); v[0] = T(a); v[1] = T(b); }
  template <typename A, typename B, typename C>
  VPP_HD vector(A a, B b, C c) { static_assert(N == 3, "3 components"); v[0] = T(a); v[1] = T(b); v[2] = T(c); }
  template <typename A, typename B, typename C, typename D>
  VPP_HD vector(A a, B b, C c, D d) { static_assert(N == 4, "4 components"); v[0] = T(a); v[1] = T(b); v[2] = T(c); v[3] = T(d); }
  VPP_HD T& operator[](int i) { return v[i]; }
  VPP_HD const T& operator[](int i) const { return v[i]; }
  VPP_HD static vector Zero() { vector r; for (unsigned i = 0; i < N; i++) r.v[i] = T(0); return r; }
  VPP_HD static vector Ones() { vector r; for (unsigned i = 0; i < N; i++) r.v[i] = T(1); return r; }
  template <typename U>
  VPP_HD vector<U, N> cast() const { vector<U, N> r; for (unsigned i = 0; i < N; i++) r.v[i] = U(v[i]); return r; }
  VPP_HD float norm() const { float s = 0; for (unsigned i = 0; i < N; i++) s += float(v[i]) * float(v[i]); return sqrtf(s); }
  VPP_HD vector& operator+=(const vector& o) { for (unsigned i = 0; i < N; i++) v[i] += o.v[i]; return *this; }
  VPP_HD vector& operator-=(const vector& o) { for (unsigned i = 0; i < N; i++) v[i] -= o.v[i]; return *this; }
  template <typename S> VPP_HD vector& operator*=(S s) { for (unsigned i = 0; i < N; i++) v[i] = T(v[i] * s); return *this; }
};

#define VPP_VEC_BINOP(op)                                                                                   \
  template <typename T, unsigned N>                                                                         \
  VPP_HD vector<T, N> operator op(const vector<T, N>& a, const vector<T, N>& b) {                             \
    vector<T, N> r; for (unsigned i = 0; i < N; i++) r.v[i] = T(a.v[i] op b.v[i]); return r; }
VPP_VEC_BINOP(+)
VPP_VEC_BINOP(-)
#undef VPP_VEC_BINOP
template <typename T, unsigned N, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
VPP_HD vector<T, N> operator*(const vector<T, N>& a, S s) { vector<T, N> r; for (unsigned i = 0; i < N; i++) r.v[i] = T(a.v[i] * s); return r; }
template <typename T, unsigned N, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
VPP_HD vector<T, N> operator*(S s, const vector<T, N>& a) { return a * s; }
template <typename T, unsigned N, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
VPP_HD vector<T, N> operator/(const vector<T, N>& a, S s) { vector<T, N> r; for (unsigned i = 0; i < N; i++) r.v[i] = T(a.v[i] / s); return r; }
template <typename T, unsigned N>
VPP_HD bool operator==(const vector<T, N>& a, const vector<T, N>& b) { for (unsigned i = 0; i < N; i++) if (!(a.v[i] == b.v[i])) return false; return true; }
template <typename T, unsigned N>
VPP_HD bool operator!=(const vector<T, N>& a, const vector<T, N>& b) { return !(a == b); }

#define VPP_ALIAS_DECL(T1, T2)                                                      \
  template <unsigned N> using v##T2 = vector<T1, N>;                                 \
  typedef v##T2<1> v##T2##1; typedef v##T2<2> v##T2##2; typedef v##T2<3> v##T2##3; \
  typedef v##T2<4> v##T2##4;
VPP_ALIAS_DECL(char, char)
VPP_ALIAS_DECL(short, short)
VPP_ALIAS_DECL(int, int)
VPP_ALIAS_DECL(float, float)
VPP_ALIAS_DECL(double, double)
VPP_ALIAS_DECL(unsigned char, uchar)
VPP_ALIAS_DECL(unsigned short, ushort)
VPP_ALIAS_DECL(unsigned int, uint)
#undef VPP_ALIAS_DECL

// vector.hh:36-50
template <typename T> struct plus_promotion_ { typedef decltype(T() + T()) type; };
template <typename X, unsigned N> struct plus_promotion_<vector<X, N>> { typedef vector<decltype(X() + X()), N> type; };
template <typename T> using plus_promotion = typename plus_promotion_<T>::type;

// vector.hh:55-109: cast<U>(v)
template <typename U, typename V> struct cast_ { VPP_HD static U run(const V& v) { return U(v); } };
template <typename US, typename VS, unsigned N> struct cast_<vector<US, N>, vector<VS, N>> {
  VPP_HD static vector<US, N> run(const vector<VS, N>& v) { return v.template cast<US>(); } };
template <typename U, typename VS> struct cast_<U, vector<VS, 1>> {  // size-1 vector -> scalar (vector.hh:70-78)
  static_assert(std::is_arithmetic<U>::value, "cast: vector<.,1> converts to a scalar or to another vector<.,1>");
  VPP_HD static U run(const vector<VS, 1>& v) { return U(v[0]); } };
template <typename US, typename VS> struct cast_<vector<US, 1>, vector<VS, 1>> {
  VPP_HD static vector<US, 1> run(const vector<VS, 1>& v) { return v.template cast<US>(); } };
template <typename U, typename V> VPP_HD U cast(const V& v) { return cast_<U, V>::run(v); }

template <typename V> struct zero { VPP_HD operator V() { return V(0); } };
template <typename T, unsigned N> struct zero<vector<T, N>> { VPP_HD operator vector<T, N>() { return vector<T, N>::Zero(); } };

}  // namespace vpp
