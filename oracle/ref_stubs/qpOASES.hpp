// Stand-in for <qpOASES.hpp> (TEST INFRASTRUCTURE ONLY).
// qpOASES (master@326a651) is not vendored by the reference and not installed here.  This facade keeps
// the SQProblem call surface balance_controller.cpp uses (bc.cpp:84, 95, 177-216) and forwards the QP
// -- in exactly the form the reference assembled it: row-major H, A, two-sided lbA/ubA, null lb/ub --
// to the oracle's dense Goldfarb-Idnani solver.  So a run of oracle/_ref pins every line of the
// reference's own assembly/epilogue code, but NOT qpOASES' arithmetic (parity for the solve stays unpinned).
#ifndef QPB_QPOASES_STANDIN
#define QPB_QPOASES_STANDIN
#include <vector>
extern "C" int orc_qp_solve(int n, int m, const double* Q, const double* c, const double* C, const double* lb,
                            const double* ub, int max_iter, double* x, double* lam, int* iters);
namespace qpOASES
{
typedef double real_t;
typedef int int_t;
enum returnValue { SUCCESSFUL_RETURN = 0, RET_INIT_FAILED = 33, RET_HOTSTART_FAILED = 57 };
enum PrintLevel { PL_NONE = 0 };
class Options {};
class SQProblem
{
public:
  SQProblem(int_t nV, int_t nC) : nV_(nV), nC_(nC), x_(nV, 0.0) {}
  void setPrintLevel(PrintLevel) {}
  bool isInitialised() const { return initialised_; }
  bool isSolved() const { return solved_; }
  returnValue init(const real_t* H, const real_t* g, const real_t* A, const real_t* lb, const real_t* ub,
                   const real_t* lbA, const real_t* ubA, int_t& nWSR, real_t* cputime = nullptr)
  {
    initialised_ = true;
    return solve(H, g, A, lb, ub, lbA, ubA, nWSR, cputime) ? SUCCESSFUL_RETURN : RET_INIT_FAILED;
  }
  returnValue hotstart(const real_t* H, const real_t* g, const real_t* A, const real_t* lb, const real_t* ub,
                       const real_t* lbA, const real_t* ubA, int_t& nWSR, real_t* cputime = nullptr)
  {
    return solve(H, g, A, lb, ub, lbA, ubA, nWSR, cputime) ? SUCCESSFUL_RETURN : RET_HOTSTART_FAILED;
  }
  returnValue getPrimalSolution(real_t* xOpt) const
  {
    for (int_t i = 0; i < nV_; i++) xOpt[i] = x_[i];
    return SUCCESSFUL_RETURN;
  }
  int_t lastIterations() const { return iters_; }

private:
  bool solve(const real_t* H, const real_t* g, const real_t* A, const real_t* lb, const real_t* ub, const real_t* lbA,
             const real_t* ubA, int_t& nWSR, real_t*)
  {
    if (lb != nullptr || ub != nullptr) return false;  // the reference passes no variable bounds (bc.cpp:165-166)
    int iters = 0;
    const int st = orc_qp_solve(nV_, nC_, H, g, A, lbA, ubA, nWSR, x_.data(), nullptr, &iters);
    nWSR = iters;
    iters_ = iters;
    solved_ = (st == 0);
    return solved_;
  }
  int_t nV_, nC_;
  std::vector<real_t> x_;
  bool initialised_ = false, solved_ = false;
  int_t iters_ = 0;
};
}  // namespace qpOASES
#endif
