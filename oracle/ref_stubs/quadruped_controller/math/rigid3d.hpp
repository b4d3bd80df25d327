// Stand-in for quadruped_controller/math/rigid3d.hpp (TEST INFRASTRUCTURE ONLY).
// The reference header wraps Drake v0.26 / Eigen, neither installed here.  Only what the balance
// controller uses is declared: skew_symmetric (rigid3d.cpp:61-74) and Rotation3d(mat) +
// angleAxisTotal() (rigid3d.cpp:177-179, 198-203), implemented in oracle/ref_glue.cpp on top of the
// oracle's restatement of Eigen's matrix -> quaternion -> angle-axis conversion.
#ifndef RIGID3D_HPP
#define RIGID3D_HPP
#include <tuple>
#include <armadillo>
namespace quadruped_controller
{
namespace math
{
using std::tuple;
using arma::mat;
using arma::vec;
using arma::vec3;

mat skew_symmetric(const vec3& x);

class Quaternion
{
};

class Rotation3d
{
public:
  Rotation3d() : R_(arma::eye(3, 3)) {}
  Rotation3d(const mat& R) : R_(R) {}
  vec angleAxisTotal() const;
  mat matrix() const { return R_; }

private:
  mat R_;
};
}  // namespace math
}  // namespace quadruped_controller
#endif
