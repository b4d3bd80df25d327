// balance_controller.hpp -- C++ shim that keeps the reference's class API on top of libqpb200.so.
//
// Mirrors (paths relative to the reference's quadruped_controller/):
//   include/quadruped_controller/balance_controller.hpp:66-180   class BalanceController
//   include/quadruped_controller/kinematics.hpp:31-113           class QuadrupedKinematics (hot-path members)
//   include/quadruped_controller/types.hpp:91-127                LegState, GaitMap, FootholdMap, ForceMap, ...
//   include/quadruped_controller/gait.hpp / src/.../gait.cpp:24-34  make_stance_gait()
// so that commander_node.cpp:337-338, 383-384 and 507-512 compile unchanged against it.
//
// With Armadillo present the reference's own mat/vec types are used; without it (this build
// image) a minimal column-major stand-in with the same element access is provided so the shim
// logic can be compiled and tested.  All arithmetic happens in the CUDA library: there is no
// CPU fallback, a failed call logs and returns an empty map exactly like the reference does when
// qpOASES fails (balance_controller.cpp:182-216).
#ifndef QPB_BALANCE_CONTROLLER_SHIM_HPP
#define QPB_BALANCE_CONTROLLER_SHIM_HPP

#include <qpb200.h>

#include <cstdio>
#include <cstring>
#include <initializer_list>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#if defined(QPB_USE_ARMADILLO) || __has_include(<armadillo>)
#include <armadillo>
#define QPB_HAVE_ARMA 1
#endif

namespace quadruped_controller
{
#ifdef QPB_HAVE_ARMA
using arma::mat;
using arma::vec;
using arma::vec3;
#else
/** Column-major dense matrix with Armadillo's element access (stand-in when Armadillo is absent). */
class mat
{
public:
  mat() = default;
  mat(unsigned int r, unsigned int c) : n_rows(r), n_cols(c), d_(static_cast<size_t>(r) * c, 0.0) {}
  mat(std::initializer_list<std::initializer_list<double>> rows)
  {
    n_rows = static_cast<unsigned int>(rows.size());
    n_cols = n_rows ? static_cast<unsigned int>(rows.begin()->size()) : 0;
    d_.assign(static_cast<size_t>(n_rows) * n_cols, 0.0);
    unsigned int i = 0;
    for (const auto& r : rows)
    {
      unsigned int j = 0;
      for (double v : r) (*this)(i, j++) = v;
      i++;
    }
  }
  double& operator()(unsigned int i, unsigned int j) { return d_.at(static_cast<size_t>(j) * n_rows + i); }
  double operator()(unsigned int i, unsigned int j) const { return d_.at(static_cast<size_t>(j) * n_rows + i); }
  unsigned int n_rows = 0, n_cols = 0;

private:
  std::vector<double> d_;
};

class vec
{
public:
  vec() = default;
  explicit vec(unsigned int n) : d_(n, 0.0) {}
  vec(std::initializer_list<double> v) : d_(v) {}
  explicit vec(const std::vector<double>& v) : d_(v) {}
  double& operator()(unsigned int i) { return d_.at(i); }
  double operator()(unsigned int i) const { return d_.at(i); }
  unsigned int size() const { return static_cast<unsigned int>(d_.size()); }

private:
  std::vector<double> d_;
};
typedef vec vec3;

inline mat eye(unsigned int r, unsigned int c)
{
  mat m(r, c);
  for (unsigned int i = 0; i < r && i < c; i++) m(i, i) = 1.0;
  return m;
}
#endif

/** types.hpp:76-88 */
struct LegJointStates
{
  LegJointStates() : q(make3()), qdot(make3()) {}
  LegJointStates(const vec3& q_, const vec3& qdot_) : q(q_), qdot(qdot_) {}
  vec3 q;
  vec3 qdot;

private:
  static vec3 make3()
  {
#ifdef QPB_HAVE_ARMA
    return vec3(arma::fill::zeros);
#else
    return vec3(3);
#endif
  }
};

/** types.hpp:91-95 */
enum LegState
{
  swing = 0,
  stance = 1
};

typedef std::map<std::string, std::pair<LegState, double>> GaitMap;  // types.hpp:100
typedef std::map<std::string, vec3> FootholdMap;                     // types.hpp:108
typedef std::map<std::string, vec3> ForceMap;                        // types.hpp:119
typedef std::map<std::string, vec3> TorqueMap;                       // types.hpp:122
typedef std::map<std::string, LegJointStates> JointStatesMap;        // types.hpp:127

/** gait.cpp:24-34 */
inline GaitMap make_stance_gait()
{
  GaitMap gait_map;
  gait_map.emplace("RL", std::make_pair(LegState::stance, 0.0));
  gait_map.emplace("FL", std::make_pair(LegState::stance, 0.0));
  gait_map.emplace("RR", std::make_pair(LegState::stance, 0.0));
  gait_map.emplace("FR", std::make_pair(LegState::stance, 0.0));
  return gait_map;
}

namespace detail
{
inline vec3 to_vec3(const double* p)
{
  vec3 v(3);
  v(0) = p[0];
  v(1) = p[1];
  v(2) = p[2];
  return v;
}
inline void log_error(const char* who, const char* what)
{
  std::fprintf(stderr, "[ERROR] [%s]: %s (%s)\n", who, what, qpb_last_error());
}
/** Position of a leg name in the library's fixed RL, FL, RR, FR layout (commander_node.cpp:61). */
inline int canonical_leg(const std::string& name)
{
  static const char* names[4] = { "RL", "FL", "RR", "FR" };
  for (int i = 0; i < 4; i++)
    if (name == names[i]) return i;
  return -1;
}
}  // namespace detail

/** @brief Reactive optimal control strategy (GPU-backed drop-in for balance_controller.hpp:66-180) */
class BalanceController
{
public:
  /** Same arguments as balance_controller.hpp:85-88. */
  BalanceController(double mu, double mass, double fzmin, double fzmax, const mat& Ib, const mat& S, const mat& W,
                    const vec& kff, const vec& kp_p, const vec& kd_p, const vec& kp_w, const vec& kd_w,
                    const std::vector<std::string>& leg_names, int device = 0)
    : leg_names_(leg_names)
  {
    qpb_params p;
    qpb_default_params(&p);  // geometry + clamp defaults; everything below is overwritten
    p.mu = mu;
    p.mass = mass;
    p.fzmin = fzmin;
    p.fzmax = fzmax;
    for (unsigned int i = 0; i < 3; i++)
      for (unsigned int j = 0; j < 3; j++) p.Ib[3 * i + j] = Ib(i, j);
    for (unsigned int i = 0; i < 6; i++)
      for (unsigned int j = 0; j < 6; j++) p.S[6 * i + j] = S(i, j);
    // The QP variables follow leg_names_ (balance_controller.cpp:109-116); the library's layout is
    // RL, FL, RR, FR.  W is permuted accordingly so any leg order gives the reference's answer.
    if (leg_names_.size() != 4) throw std::invalid_argument("BalanceController: exactly 4 leg names required");
    int perm[4];
    for (int i = 0; i < 4; i++)
    {
      perm[i] = detail::canonical_leg(leg_names_[i]);
      if (perm[i] < 0) throw std::invalid_argument("BalanceController: leg names must be RL, FL, RR, FR");
    }
    for (unsigned int i = 0; i < 12; i++)
      for (unsigned int j = 0; j < 12; j++)
        p.W[12 * (3 * perm[i / 3] + i % 3) + (3 * perm[j / 3] + j % 3)] = W(i, j);
    for (unsigned int i = 0; i < 6; i++) p.kff[i] = kff(i);
    for (unsigned int i = 0; i < 3; i++)
    {
      p.kp_p[i] = kp_p(i);
      p.kd_p[i] = kd_p(i);
      p.kp_w[i] = kp_w(i);
      p.kd_w[i] = kd_w(i);
    }
    if (qpb_create(&p, device, &handle_) != QPB_SUCCESS)
    {
      detail::log_error("Balance Controller", "failed to create the GPU balance controller");
      handle_ = nullptr;
    }
  }
  ~BalanceController() { qpb_destroy(handle_); }
  BalanceController(const BalanceController&) = delete;
  BalanceController& operator=(const BalanceController&) = delete;

  /** Same arguments and return value as balance_controller.hpp:104-107. */
  ForceMap control(const mat& Rwb, const mat& Rwb_d, const vec& x, const vec& xdot, const vec& w, const vec& x_d,
                   const vec& xdot_d, const vec& w_d, const FootholdMap& foot_map,
                   const GaitMap& gait_map = make_stance_gait()) const
  {
    ForceMap force_map;
    qpb_state_rec s;
    pack(s, Rwb, Rwb_d, x, xdot, w, x_d, xdot_d, w_d, foot_map, gait_map);
    qpb_out_rec o;
    if (!handle_ || qpb_control_batch_host(handle_, 1, &s, &o) != QPB_SUCCESS)
    {
      detail::log_error("Balance Controller", "GPU balance QP call failed");
      working_set_ = 0;
      return force_map;
    }
    remember(o);
    if (o.status != QPB_OK)
    {
      detail::log_error("Balance Controller", "Balance Controller QP Solver Failed");
      return force_map;  // empty map, balance_controller.cpp:212-216
    }
    for (const auto& leg_name : leg_names_)
      if (gait_map.at(leg_name).first == LegState::stance)  // stance legs only, :223-228
        force_map.emplace(leg_name, detail::to_vec3(&o.grf_body[3 * detail::canonical_leg(leg_name)]));
    return force_map;
  }

  /** Extension: control() and jacobianTransposeControl() from the same kernel launch. */
  std::pair<ForceMap, TorqueMap> controlWithTorques(const mat& Rwb, const mat& Rwb_d, const vec& x, const vec& xdot,
                                                    const vec& w, const vec& x_d, const vec& xdot_d, const vec& w_d,
                                                    const FootholdMap& foot_map, const JointStatesMap& joint_states_map,
                                                    const GaitMap& gait_map = make_stance_gait()) const
  {
    std::pair<ForceMap, TorqueMap> result;
    qpb_state_rec s;
    pack(s, Rwb, Rwb_d, x, xdot, w, x_d, xdot_d, w_d, foot_map, gait_map);
    for (const auto& leg_name : leg_names_)
    {
      const int c = detail::canonical_leg(leg_name);
      const auto& js = joint_states_map.at(leg_name);
      for (unsigned int k = 0; k < 3; k++) s.q[3 * c + k] = js.q(k);
    }
    qpb_out_rec o;
    if (!handle_ || qpb_control_batch_host(handle_, 1, &s, &o) != QPB_SUCCESS)
    {
      detail::log_error("Balance Controller", "GPU balance QP call failed");
      working_set_ = 0;
      return result;
    }
    remember(o);
    if (o.status != QPB_OK)
    {
      detail::log_error("Balance Controller", "Balance Controller QP Solver Failed");
      return result;
    }
    for (const auto& leg_name : leg_names_)
      if (gait_map.at(leg_name).first == LegState::stance)
      {
        const int c = detail::canonical_leg(leg_name);
        result.first.emplace(leg_name, detail::to_vec3(&o.grf_body[3 * c]));
        result.second.emplace(leg_name, detail::to_vec3(&o.tau[3 * c]));
      }
    return result;
  }

private:
  void pack(qpb_state_rec& s, const mat& Rwb, const mat& Rwb_d, const vec& x, const vec& xdot, const vec& w,
            const vec& x_d, const vec& xdot_d, const vec& w_d, const FootholdMap& foot_map,
            const GaitMap& gait_map) const
  {
    std::memset(&s, 0, sizeof(s));
    for (unsigned int i = 0; i < 3; i++)
    {
      for (unsigned int j = 0; j < 3; j++)
      {
        s.Rwb[3 * i + j] = Rwb(i, j);
        s.Rwb_d[3 * i + j] = Rwb_d(i, j);
      }
      s.x[i] = x(i);
      s.xdot[i] = xdot(i);
      s.w[i] = w(i);
      s.x_d[i] = x_d(i);
      s.xdot_d[i] = xdot_d(i);
      s.w_d[i] = w_d(i);
    }
    for (const auto& leg_name : leg_names_)
    {
      const int c = detail::canonical_leg(leg_name);
      const vec3& p = foot_map.at(leg_name);  // throws std::out_of_range like :115
      for (unsigned int k = 0; k < 3; k++) s.feet[3 * c + k] = p(k);
      s.contact[c] = (gait_map.at(leg_name).first == LegState::stance) ? 1 : 0;  // :223, :312
    }
    // the previous tick's final working set: the library starts its active-set method there (qpb200.h, pad[0:4])
    std::memcpy(s.pad, &working_set_, sizeof(working_set_));
  }

  /** Keep the working set this tick ended on for the next one -- what SQProblem::hotstart does inside the reference's
   *  mutable QPSolver_ (balance_controller.hpp:161, balance_controller.cpp:176-199).  A failed solve starts cold. */
  void remember(const qpb_out_rec& o) const
  {
    working_set_ = 0;
    if (o.status == QPB_OK) std::memcpy(&working_set_, o.pad, sizeof(working_set_));
  }

  mutable uint32_t working_set_ = 0;  // bit 31 set = bits 0..23 hold a working set (as the library returns it)
  qpb_handle* handle_ = nullptr;
  std::vector<std::string> leg_names_;
};

/** @brief Kinematic model (GPU-backed hot-path members of kinematics.hpp:31-113) */
class QuadrupedKinematics
{
public:
  explicit QuadrupedKinematics(int device = 0)
  {
    qpb_params p;
    qpb_default_params(&p);  // the geometry kinematics.cpp:23-47 hard-codes
    if (qpb_create(&p, device, &handle_) != QPB_SUCCESS)
    {
      detail::log_error("Kinematics", "failed to create the GPU kinematics");
      handle_ = nullptr;
    }
  }
  ~QuadrupedKinematics() { qpb_destroy(handle_); }
  QuadrupedKinematics(const QuadrupedKinematics&) = delete;
  QuadrupedKinematics& operator=(const QuadrupedKinematics&) = delete;

  /** kinematics.hpp:60 / kinematics.cpp:105-115 */
  FootholdMap forwardKinematics(const JointStatesMap& joint_states_map) const
  {
    FootholdMap foot_hold_map;
    double q[12] = { 0 }, feet[12];
    for (const auto& [leg_name, js] : joint_states_map)
    {
      const int c = detail::canonical_leg(leg_name);
      if (c < 0) throw std::out_of_range("QuadrupedKinematics: unknown leg " + leg_name);  // link_map_.at, :84
      for (unsigned int k = 0; k < 3; k++) q[3 * c + k] = js.q(k);
    }
    if (!handle_ || qpb_fk_batch_host(handle_, 1, q, feet) != QPB_SUCCESS)
    {
      detail::log_error("Kinematics", "GPU forward kinematics failed");
      return foot_hold_map;
    }
    for (const auto& kv : joint_states_map)
      foot_hold_map.emplace(kv.first, detail::to_vec3(&feet[3 * detail::canonical_leg(kv.first)]));
    return foot_hold_map;
  }

  /** kinematics.hpp:106-107 / kinematics.cpp:218-231: only legs present in force_map get a torque. */
  TorqueMap jacobianTransposeControl(const JointStatesMap& joint_states_map, const ForceMap& force_map) const
  {
    TorqueMap torque_map;
    double q[12] = { 0 }, f[12] = { 0 }, tau[12];
    uint8_t present[4] = { 0, 0, 0, 0 };
    for (const auto& [leg_name, force] : force_map)
    {
      const int c = detail::canonical_leg(leg_name);
      if (c < 0) throw std::out_of_range("QuadrupedKinematics: unknown leg " + leg_name);
      const auto& js = joint_states_map.at(leg_name);
      for (unsigned int k = 0; k < 3; k++)
      {
        q[3 * c + k] = js.q(k);
        f[3 * c + k] = force(k);
      }
      present[c] = 1;
    }
    if (force_map.empty()) return torque_map;
    if (!handle_ || qpb_jt_batch_host(handle_, 1, q, f, present, tau) != QPB_SUCCESS)
    {
      detail::log_error("Kinematics", "GPU Jacobian-transpose map failed");
      return torque_map;
    }
    for (const auto& kv : force_map)
      torque_map.emplace(kv.first, detail::to_vec3(&tau[3 * detail::canonical_leg(kv.first)]));
    return torque_map;
  }

private:
  qpb_handle* handle_ = nullptr;
};

}  // namespace quadruped_controller
#endif
