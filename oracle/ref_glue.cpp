// ref_glue.cpp -- flat C entry points around the REFERENCE's own classes (TEST INFRASTRUCTURE ONLY).
//
// oracle/ref_build.sh compiles, from the sources where they lie under /root/reference/quadruped_controller,
//   src/quadruped_controller/balance_controller.cpp, kinematics.cpp, gait.cpp, math/numerics.cpp
// against the stand-in headers in oracle/ref_stubs/ (Armadillo, qpOASES, ROS console, rigid3d) and links
// this file.  What that pins: every line of the reference's PD target, dynamics, QP assembly, row-major
// marshalling, friction cone/bounds, world->body epilogue, stance-only maps, leg Jacobian and torque map
// (SURVEY 8a rows a1-a4, a6, a8, a10-a14, a16, a17).  What it does not pin: qpOASES' and Drake/Eigen's
// arithmetic, which the stand-ins forward to the oracle (orc_qp_solve, orc_angle_axis_total).
#include <memory>
#include <cstring>
#include <thread>
#include <vector>

#include <quadruped_controller/balance_controller.hpp>
#include <quadruped_controller/kinematics.hpp>

#include "qpb_oracle.h"

int qpb_ref_error_count = 0;

namespace quadruped_controller
{
namespace math
{
// restates rigid3d.cpp:61-74 (the reference file itself needs Drake and is not compiled)
mat skew_symmetric(const vec3& x)
{
  mat skew(3, 3, arma::fill::zeros);
  skew(0, 1) = -x(2);
  skew(0, 2) = x(1);
  skew(1, 2) = -x(0);
  skew(1, 0) = x(2);
  skew(2, 0) = -x(1);
  skew(2, 1) = x(0);
  return skew;
}

// rigid3d.cpp:198-203 -> drake RotationMatrix::ToAngleAxis -> Eigen::AngleAxisd(Matrix3d)
vec Rotation3d::angleAxisTotal() const
{
  double R[9], aa[3];
  for (unsigned i = 0; i < 3; i++)
    for (unsigned j = 0; j < 3; j++) R[3 * i + j] = R_(i, j);
  orc_angle_axis_total(R, aa);
  return { aa[0], aa[1], aa[2] };
}
}  // namespace math
}  // namespace quadruped_controller

using namespace quadruped_controller;

namespace
{
const std::vector<std::string> kLegs = { "RL", "FL", "RR", "FR" };  // commander_node.cpp:61

struct Cached
{
  orc_params params;
  std::unique_ptr<BalanceController> bc;
  std::unique_ptr<QuadrupedKinematics> kin;
};
thread_local Cached g_cache;  // control() mutates its solver: one controller per thread (bc.hpp:161)

mat to_mat(const double* a, unsigned r, unsigned c)
{
  mat m(r, c);
  for (unsigned i = 0; i < r; i++)
    for (unsigned j = 0; j < c; j++) m(i, j) = a[i * c + j];
  return m;
}
vec to_vec(const double* a, unsigned n)
{
  vec v(n);
  for (unsigned i = 0; i < n; i++) v(i) = a[i];
  return v;
}

void ensure(const orc_params* p)
{
  if (g_cache.bc && std::memcmp(&g_cache.params, p, sizeof(orc_params)) == 0) return;
  g_cache.params = *p;
  // commander_node.cpp:337-338
  g_cache.bc.reset(new BalanceController(p->mu, p->mass, p->fzmin, p->fzmax, to_mat(p->Ib, 3, 3), to_mat(p->S, 6, 6),
                                         to_mat(p->W, 12, 12), to_vec(p->kff, 6), to_vec(p->kp_p, 3), to_vec(p->kd_p, 3),
                                         to_vec(p->kp_w, 3), to_vec(p->kd_w, 3), kLegs));
  g_cache.kin.reset(new QuadrupedKinematics());
}
}  // namespace

extern "C" {

/** One state through the reference's control() + jacobianTransposeControl() (commander_node.cpp:507-512).
 *  Returns the number of legs in the returned ForceMap; out->status = 0 if that equals the stance count. */
int ref_control(const orc_params* p, const orc_state* s, orc_out* out)
{
  ensure(p);
  std::memset(out, 0, sizeof(*out));
  FootholdMap foot_map;
  GaitMap gait_map;
  JointStatesMap joint_states_map;
  int n_stance = 0;
  for (unsigned leg = 0; leg < 4; leg++)
  {
    foot_map.emplace(kLegs[leg], vec3({ s->feet[3 * leg], s->feet[3 * leg + 1], s->feet[3 * leg + 2] }));
    const bool st = s->contact[leg] != 0;
    n_stance += st ? 1 : 0;
    gait_map.emplace(kLegs[leg], std::make_pair(st ? LegState::stance : LegState::swing, 0.0));
    LegJointStates js;
    for (unsigned k = 0; k < 3; k++) js.q(k) = s->q[3 * leg + k];
    joint_states_map.emplace(kLegs[leg], js);
  }
  const ForceMap force_map = g_cache.bc->control(to_mat(s->Rwb, 3, 3), to_mat(s->Rwb_d, 3, 3), to_vec(s->x, 3),
                                                 to_vec(s->xdot, 3), to_vec(s->w, 3), to_vec(s->x_d, 3),
                                                 to_vec(s->xdot_d, 3), to_vec(s->w_d, 3), foot_map, gait_map);
  const TorqueMap torque_map = g_cache.kin->jacobianTransposeControl(joint_states_map, force_map);
  for (unsigned leg = 0; leg < 4; leg++)
  {
    const auto f = force_map.find(kLegs[leg]);
    if (f == force_map.end()) continue;
    const auto t = torque_map.find(kLegs[leg]);
    for (unsigned k = 0; k < 3; k++)
    {
      out->grf_body[3 * leg + k] = f->second(k);
      out->tau[3 * leg + k] = t->second(k);
    }
  }
  out->status = (static_cast<int>(force_map.size()) == n_stance) ? 0 : 2;
  return static_cast<int>(force_map.size());
}

void ref_control_batch(const orc_params* p, const orc_state* s, long long n, orc_out* out, int nthreads)
{
  if (nthreads <= 1)
  {
    for (long long i = 0; i < n; i++) ref_control(p, &s[i], &out[i]);
    return;
  }
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads; t++)
    pool.emplace_back([=]() {
      for (long long i = n * t / nthreads; i < n * (t + 1) / nthreads; i++) ref_control(p, &s[i], &out[i]);
    });
  for (auto& th : pool) th.join();
}

/** forwardKinematics(JointStatesMap) (kinematics.cpp:105-115), legs RL FL RR FR */
void ref_forward_kinematics(const double* q12, double* feet12)
{
  QuadrupedKinematics kin;
  JointStatesMap jsm;
  for (unsigned leg = 0; leg < 4; leg++)
  {
    LegJointStates js;
    for (unsigned k = 0; k < 3; k++) js.q(k) = q12[3 * leg + k];
    jsm.emplace(kLegs[leg], js);
  }
  const FootholdMap fm = kin.forwardKinematics(jsm);
  for (unsigned leg = 0; leg < 4; leg++)
    for (unsigned k = 0; k < 3; k++) feet12[3 * leg + k] = fm.at(kLegs[leg])(k);
}

/** legJacobian (kinematics.cpp:162-188), row-major */
void ref_leg_jacobian(int leg, const double* q3, double* J9)
{
  QuadrupedKinematics kin;
  const arma::mat33 J = kin.legJacobian(kLegs[leg], arma::vec3({ q3[0], q3[1], q3[2] }));
  for (unsigned i = 0; i < 3; i++)
    for (unsigned j = 0; j < 3; j++) J9[3 * i + j] = J(i, j);
}

int ref_error_count(void) { return qpb_ref_error_count; }

}  // extern "C"

// ---- swing-leg half of the control tick: the reference's own legInverseKinematics, legJacobianInverse and
//      JointController::control, driven exactly as commander_node.cpp:482-533 drives them -----------------------
#include <quadruped_controller/joint_controller.hpp>

extern "C" {

void ref_leg_inverse_kinematics(int leg, const double* foothold3, double* q3)
{
  QuadrupedKinematics kin;
  const arma::vec3 q = kin.legInverseKinematics(kLegs[leg], arma::vec3({ foothold3[0], foothold3[1], foothold3[2] }));
  for (unsigned k = 0; k < 3; k++) q3[k] = q(k);
}

void ref_leg_jacobian_inverse(int leg, const double* q3, double* Jinv9)
{
  QuadrupedKinematics kin;
  const arma::mat33 Ji = kin.legJacobianInverse(kLegs[leg], arma::vec3({ q3[0], q3[1], q3[2] }));
  for (unsigned i = 0; i < 3; i++)
    for (unsigned j = 0; j < 3; j++) Jinv9[3 * i + j] = Ji(i, j);
}

int ref_tick(const orc_params* p, const orc_joint_gains* g, const orc_state* s, const orc_swing* sw, orc_out* out)
{
  ensure(p);
  std::memset(out, 0, sizeof(*out));
  const JointController joint_controller(arma::vec3({ g->kff[0], g->kff[1], g->kff[2] }), arma::vec3({ g->kp[0], g->kp[1], g->kp[2] }),
                                         arma::vec3({ g->kd[0], g->kd[1], g->kd[2] }));  // commander_node.cpp:341
  const mat Rwb = to_mat(s->Rwb, 3, 3);
  const vec x = to_vec(s->x, 3);
  FootholdMap foot_map;
  GaitMap gait_map;
  JointStatesMap joint_states_map;
  int n_stance = 0;
  for (unsigned leg = 0; leg < 4; leg++)
  {
    foot_map.emplace(kLegs[leg], vec3({ s->feet[3 * leg], s->feet[3 * leg + 1], s->feet[3 * leg + 2] }));
    const bool st = s->contact[leg] != 0;
    n_stance += st ? 1 : 0;
    gait_map.emplace(kLegs[leg], std::make_pair(st ? LegState::stance : LegState::swing, 0.0));
    LegJointStates js;
    for (unsigned k = 0; k < 3; k++)
    {
      js.q(k) = s->q[3 * leg + k];
      js.qdot(k) = sw->qdot[3 * leg + k];
    }
    joint_states_map.emplace(kLegs[leg], js);
  }
  // commander_node.cpp:482-500
  JointStatesMap swing_leg_js_map;
  for (const auto& [leg_name, leg_state] : gait_map)
  {
    if (leg_state.first == LegState::swing)
    {
      unsigned leg = 0;
      while (kLegs[leg] != leg_name) leg++;
      FootState foot_state(vec3({ sw->foot_ref_pos[3 * leg], sw->foot_ref_pos[3 * leg + 1], sw->foot_ref_pos[3 * leg + 2] }),
                           vec3({ sw->foot_ref_vel[3 * leg], sw->foot_ref_vel[3 * leg + 1], sw->foot_ref_vel[3 * leg + 2] }));
      foot_state.position = Rwb.t() * foot_state.position - x;
      foot_state.velocity = Rwb.t() * foot_state.velocity;
      const vec3 q = g_cache.kin->legInverseKinematics(leg_name, foot_state.position);
      const vec3 qdot = g_cache.kin->legJacobianInverse(leg_name, q) * foot_state.velocity;
      swing_leg_js_map.emplace(leg_name, LegJointStates(q, qdot));
    }
  }
  // :503-515
  const TorqueMap swing_torque_map = joint_controller.control(swing_leg_js_map, joint_states_map);
  const ForceMap force_map = g_cache.bc->control(Rwb, to_mat(s->Rwb_d, 3, 3), x, to_vec(s->xdot, 3), to_vec(s->w, 3),
                                                 to_vec(s->x_d, 3), to_vec(s->xdot_d, 3), to_vec(s->w_d, 3), foot_map, gait_map);
  TorqueMap torque_map = g_cache.kin->jacobianTransposeControl(joint_states_map, force_map);
  torque_map.insert(swing_torque_map.begin(), swing_torque_map.end());
  for (unsigned leg = 0; leg < 4; leg++)
  {
    const auto f = force_map.find(kLegs[leg]);
    if (f != force_map.end())
      for (unsigned k = 0; k < 3; k++) out->grf_body[3 * leg + k] = f->second(k);
    const auto t = torque_map.find(kLegs[leg]);
    if (t != torque_map.end())
      for (unsigned k = 0; k < 3; k++)
      {
        double tau = t->second(k);
        if (p->clamp_tau) tau = std::fmin(std::fmax(tau, p->tau_min), p->tau_max);  // arma::clamp, :526
        out->tau[3 * leg + k] = tau;
      }
  }
  out->status = (static_cast<int>(force_map.size()) == n_stance) ? 0 : 2;
  return static_cast<int>(torque_map.size());
}

void ref_tick_batch(const orc_params* p, const orc_joint_gains* g, const orc_state* s, const orc_swing* sw, long long n,
                    orc_out* out, int nthreads)
{
  if (nthreads <= 1)
  {
    for (long long i = 0; i < n; i++) ref_tick(p, g, &s[i], &sw[i], &out[i]);
    return;
  }
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads; t++)
    pool.emplace_back([=]() {
      for (long long i = n * t / nthreads; i < n * (t + 1) / nthreads; i++) ref_tick(p, g, &s[i], &sw[i], &out[i]);
    });
  for (auto& th : pool) th.join();
}

}  // extern "C"
