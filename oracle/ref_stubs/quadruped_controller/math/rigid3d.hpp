// Stand-in for quadruped_controller/math/rigid3d.hpp (TEST INFRASTRUCTURE ONLY).
// The reference header wraps Drake v0.26 / Eigen, neither installed here.  Only what the balance
// controller uses is declared: skew_symmetric (rigid3d.cpp:61-74) and Rotation3d(mat) +
// angleAxisTotal() (rigid3d.cpp:177-179, 198-203), implemented in oracle/ref_glue.cpp on top of the
// oracle's restatement of Eigen's matrix -> quaternion -> angle-axis conversion.
#ifndef RIGID3D_HPP
#define RIGID3D_HPP
#include <cstdlib>
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

class Rotation3d;
// Declared so that trajectory.cpp (integrate_twist_yaw, Drake-backed in the reference) compiles; the glue never
// calls these paths and every member aborts if reached.
class Quaternion
{
public:
  Quaternion() {}
  Quaternion(double, const vec3&) { std::abort(); }
  mat matrix() const { std::abort(); }
  vec3 eulerAngles() const { std::abort(); }
};

class Rotation3d
{
public:
  Rotation3d() : R_(arma::eye(3, 3)) {}
  Rotation3d(const mat& R) : R_(R) {}
  Rotation3d(double, double, double) { std::abort(); }
  vec angleAxisTotal() const;
  mat matrix() const { return R_; }

private:
  mat R_;
};
class Transform3d
{
public:
  Transform3d() {}
  Transform3d(const Rotation3d&, const vec3&) { std::abort(); }
  Transform3d(const mat&, const vec3&) { std::abort(); }
  Transform3d operator*(const Transform3d&) const { std::abort(); }
  Quaternion getQuaternion() const { std::abort(); }
  mat adjoint() const { std::abort(); }
};
struct Pose
{
  Pose() {}
  Pose(const mat&, const vec3&) { std::abort(); }
  explicit Pose(const Transform3d&) { std::abort(); }
  Transform3d transform() const { std::abort(); }
  vec3 position;
  Quaternion orientation;
};
}  // namespace math
}  // namespace quadruped_controller
#endif
