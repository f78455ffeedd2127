// Minimal stand-in for the subset of iod (github.com/matt-42/iod) used by the Video++ headers:
// symbols `_name`, `_name = value`, iod::D(...), sio::has / get / member access, static_if, has_symbol.
// TEST INFRASTRUCTURE (see Eigen/Core in this directory).
#pragma once
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

namespace iod {

struct variable_base {};  // makes iod an associated namespace of every `_sym = value` (unqualified D(...) calls)

template <typename S>
struct symbol {
  typedef S symbol_type;
  template <typename V>
  constexpr auto operator=(V&& v) const {
    return typename S::template variable_type<typename std::decay<V>::type>{std::forward<V>(v)};
  }
};

template <typename... T>
struct sio;

namespace internal {
template <typename S, typename... T> struct find_symbol;
template <typename S> struct find_symbol<S> { typedef void type; enum { value = 0 }; };
template <typename S, typename T1, typename... T>
struct find_symbol<S, T1, T...> {
  enum { match = std::is_same<S, typename T1::symbol_type>::value };
  typedef typename std::conditional<match, T1, typename find_symbol<S, T...>::type>::type type;
  enum { value = match || find_symbol<S, T...>::value };
};
// a bare symbol passed to D() is a flag: it becomes variable_type<bool>{true}
template <typename A, typename Enable = void> struct to_member { typedef A type; static A make(const A& a) { return a; } };
template <typename A>
struct to_member<A, typename std::enable_if<std::is_base_of<symbol<A>, A>::value>::type> {
  typedef typename A::template variable_type<bool> type;
  static type make(const A&) { return type{true}; }
};
}  // namespace internal

template <typename... T>
struct sio : public T... {
  template <typename... U, typename = typename std::enable_if<sizeof...(U) == sizeof...(T) && (sizeof...(U) > 0)>::type>
  sio(const U&... t) : T(t)... {}
  sio() {}
  template <typename S> static constexpr bool has(const S&) { return internal::find_symbol<S, T...>::value; }
  template <typename S, typename D>
  auto get(const S&, const D& dflt) const { return get_(static_cast<S*>(nullptr), dflt, std::integral_constant<bool, internal::find_symbol<S, T...>::value>()); }
 private:
  template <typename S, typename D> auto get_(S*, const D&, std::true_type) const {
    typedef typename internal::find_symbol<S, T...>::type M;
    return static_cast<const M*>(this)->value();
  }
  template <typename S, typename D> D get_(S*, const D& dflt, std::false_type) const { return dflt; }
};

template <typename... A>
auto D(const A&... a) { return sio<typename internal::to_member<A>::type...>(internal::to_member<A>::make(a)...); }

template <typename O, typename S> struct has_symbol;
template <typename... T, typename S> struct has_symbol<sio<T...>, S> { enum { value = internal::find_symbol<S, T...>::value }; };

template <bool C, typename F, typename G, typename... A>
auto static_if_(std::true_type, F f, G, A&&... a) { return f(std::forward<A>(a)...); }
template <bool C, typename F, typename G, typename... A>
auto static_if_(std::false_type, F, G g, A&&... a) { return g(std::forward<A>(a)...); }
template <bool C, typename F, typename G, typename... A>
auto static_if(F f, G g, A&&... a) { return static_if_<C>(std::integral_constant<bool, C>(), f, g, std::forward<A>(a)...); }

}  // namespace iod

#define iod_define_symbol(NAME)                                        \
  namespace s {                                                        \
  struct _##NAME##_t : iod::symbol<_##NAME##_t> {                      \
    using iod::symbol<_##NAME##_t>::operator=;                         \
    constexpr _##NAME##_t() {}                                         \
    template <typename T>                                              \
    struct variable_type : iod::variable_base {                        \
      typedef _##NAME##_t symbol_type;                                 \
      typedef T value_type;                                            \
      variable_type() {}                                               \
      variable_type(const T& v) : NAME(v) {}                           \
      const T& value() const { return NAME; }                          \
      T NAME;                                                          \
    };                                                                 \
  };                                                                   \
  static constexpr _##NAME##_t _##NAME{};                              \
  }

#define iod_define_number_symbol(N)
