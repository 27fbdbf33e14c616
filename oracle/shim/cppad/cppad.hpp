// oracle/shim/cppad — NOT CppAD: AD<double> is a plain double wrapper (values only, no derivatives), enough
// to EVALUATE FG_EVAL::operator() (src/model_predictive_control.cpp:199-252) and to compile update() (:69-81),
// which calls CppAD::tan on a float.  TEST INFRASTRUCTURE ONLY (oracle/_ref).
#ifndef CRB_SHIM_CPPAD_
#define CRB_SHIM_CPPAD_
#include <cmath>
#include <string>
#include <vector>
#define CPPAD_TESTVECTOR(T) std::vector<T>
namespace CppAD {
template <typename B>
struct AD {
  B v;
  AD() : v(0) {}
  AD(B x) : v(x) {}
  AD(int x) : v((B)x) {}
  AD(float x) : v((B)x) {}
  AD& operator+=(const AD& o) { v += o.v; return *this; }
};
template <typename B> AD<B> operator+(const AD<B>& a, const AD<B>& b) { return AD<B>(a.v + b.v); }
template <typename B> AD<B> operator-(const AD<B>& a, const AD<B>& b) { return AD<B>(a.v - b.v); }
template <typename B> AD<B> operator*(const AD<B>& a, const AD<B>& b) { return AD<B>(a.v * b.v); }
template <typename B> AD<B> operator/(const AD<B>& a, const AD<B>& b) { return AD<B>(a.v / b.v); }
#define CRB_AD_MIXED(op)                                                                     \
  template <typename B> AD<B> operator op(const AD<B>& a, double b) { return AD<B>(a.v op b); } \
  template <typename B> AD<B> operator op(double a, const AD<B>& b) { return AD<B>(a op b.v); }
CRB_AD_MIXED(+) CRB_AD_MIXED(-) CRB_AD_MIXED(*) CRB_AD_MIXED(/)
#undef CRB_AD_MIXED
template <typename B> AD<B> pow(const AD<B>& a, int n) { return AD<B>(std::pow(a.v, n)); }
template <typename B> AD<B> cos(const AD<B>& a) { return AD<B>(std::cos(a.v)); }
template <typename B> AD<B> sin(const AD<B>& a) { return AD<B>(std::sin(a.v)); }
template <typename B> AD<B> tan(const AD<B>& a) { return AD<B>(std::tan(a.v)); }
// CppAD re-exports the standard math functions for base types: CppAD::tan(float) is std::tan(float)
inline float tan(float x) { return std::tan(x); }
inline double tan(double x) { return std::tan(x); }
template <typename B> B Value(const AD<B>& a) { return a.v; }
}  // namespace CppAD
#endif
