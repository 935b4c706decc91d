// ORACLE (test infrastructure only) -- forward-mode dual numbers.
// Stands in for CppAD: the reference obtains every model Jacobian by taping the templated model code on
// CppAD::AD<CppAD::cg::CG<double>> (ocs2_core/src/automatic_differentation/CppAdInterface.cpp:151-187); running the same
// templated code on Dual<N> yields the same derivatives up to round-off.
#pragma once
#include <cmath>

namespace orc {

using std::cos;
using std::sin;
using std::sqrt;

template <int N>
struct Dual {
  double v;
  double d[N];
  Dual() : v(0.0) {
    for (int i = 0; i < N; ++i) d[i] = 0.0;
  }
  Dual(double x) : v(x) {  // NOLINT implicit on purpose
    for (int i = 0; i < N; ++i) d[i] = 0.0;
  }
  static Dual variable(double x, int idx) {
    Dual r(x);
    r.d[idx] = 1.0;
    return r;
  }
};

template <int N>
inline Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = a.v + b.v;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <int N>
inline Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = a.v - b.v;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <int N>
inline Dual<N> operator-(const Dual<N>& a) {
  Dual<N> r;
  r.v = -a.v;
  for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
  return r;
}
template <int N>
inline Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = a.v * b.v;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
template <int N>
inline Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  const double inv = 1.0 / b.v;
  r.v = a.v * inv;
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r;
}
#define ORC_DUAL_MIXED(op)                                                   \
  template <int N>                                                           \
  inline Dual<N> operator op(const Dual<N>& a, double b) { return a op Dual<N>(b); } \
  template <int N>                                                           \
  inline Dual<N> operator op(double a, const Dual<N>& b) { return Dual<N>(a) op b; }
ORC_DUAL_MIXED(+)
ORC_DUAL_MIXED(-)
ORC_DUAL_MIXED(*)
ORC_DUAL_MIXED(/)
#undef ORC_DUAL_MIXED
template <int N>
inline Dual<N>& operator+=(Dual<N>& a, const Dual<N>& b) {
  a = a + b;
  return a;
}
template <int N>
inline Dual<N>& operator-=(Dual<N>& a, const Dual<N>& b) {
  a = a - b;
  return a;
}
template <int N>
inline Dual<N> sin(const Dual<N>& a) {
  Dual<N> r;
  r.v = std::sin(a.v);
  const double c = std::cos(a.v);
  for (int i = 0; i < N; ++i) r.d[i] = c * a.d[i];
  return r;
}
template <int N>
inline Dual<N> cos(const Dual<N>& a) {
  Dual<N> r;
  r.v = std::cos(a.v);
  const double s = -std::sin(a.v);
  for (int i = 0; i < N; ++i) r.d[i] = s * a.d[i];
  return r;
}
template <int N>
inline Dual<N> sqrt(const Dual<N>& a) {
  Dual<N> r;
  r.v = std::sqrt(a.v);
  const double h = 0.5 / r.v;
  for (int i = 0; i < N; ++i) r.d[i] = h * a.d[i];
  return r;
}
inline double value(double x) { return x; }
template <int N>
inline double value(const Dual<N>& x) {
  return x.v;
}

}  // namespace orc
