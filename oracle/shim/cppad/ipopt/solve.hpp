// oracle/shim/cppad/ipopt — NOT the CppAD/IPOPT bridge: solve() does not solve anything.  It evaluates the
// caller's FG_EVAL once at the initial point (so the functor is instantiated exactly as the reference
// instantiates it), returns that point and reports failure.  mpc_solve() compiled against it is therefore
// NOT a usable solver and no test treats it as one; it exists so that the translation unit links.
#ifndef CRB_SHIM_CPPAD_IPOPT_
#define CRB_SHIM_CPPAD_IPOPT_
#include <string>
#include "../cppad.hpp"
namespace CppAD { namespace ipopt {
template <typename Dvector>
struct solve_result {
  enum status_type { not_defined, success, maxiter_exceeded, unknown };
  status_type status;
  Dvector x;
  solve_result() : status(not_defined) {}
};
template <typename Dvector, typename FG_eval>
void solve(const std::string&, const Dvector& xi, const Dvector&, const Dvector&, const Dvector& gl,
           const Dvector&, FG_eval& fg_eval, solve_result<Dvector>& solution) {
  typename FG_eval::ADvector fg(1 + gl.size()), vars(xi.size());
  for (size_t i = 0; i < xi.size(); ++i) vars[i] = xi[i];
  fg_eval(fg, vars);
  solution.x = xi;
  solution.status = solve_result<Dvector>::unknown;
}
} }  // namespace CppAD::ipopt
#endif
