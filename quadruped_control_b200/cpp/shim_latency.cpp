// shim_latency.cpp -- per-call latency of the drop-in classes for ONE robot (BASELINE config 1): the reference's tick
// sequence control() + jacobianTransposeControl() (commander_node.cpp:507-512), forwardKinematics() (:383-384) and the
// shim's single-launch controlWithTorques().  Prints one JSON line of mean microseconds per call.
#include <balance_controller.hpp>

#include <chrono>
#include <cstdio>

using namespace quadruped_controller;

template <class F>
static double mean_us(F&& f, int iters)
{
  for (int i = 0; i < 200; i++) f();
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < iters; i++) f();
  return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / iters;
}

int main()
{
  const std::vector<std::string> leg_names = { "RL", "FL", "RR", "FR" };
  mat Ib = eye(3, 3);
  Ib(0, 0) = 0.011253;
  Ib(1, 1) = 0.036203;
  Ib(2, 2) = 0.042673;
  mat S = eye(6, 6);
  const double sd[6] = { 1.0, 1.0, 1.0, 10.0, 10.0, 5.0 };
  for (unsigned int i = 0; i < 6; i++) S(i, i) = sd[i];
  mat W = eye(12, 12);
  for (unsigned int i = 0; i < 12; i++) W(i, i) = 1e-5;
  const vec kff{ 0.0, 0.0, 0.15, 0.0, 0.0, 0.0 };
  const vec kp_p{ 100.0, 100.0, 100.0 }, kd_p{ 50.0, 50.0, 50.0 };
  const vec kp_w{ 5000.0, 5000.0, 5000.0 }, kd_w{ 500.0, 500.0, 500.0 };
  const BalanceController bc(0.6, 11.0, 10.0, 120.0, Ib, S, W, kff, kp_p, kd_p, kp_w, kd_w, leg_names);
  const QuadrupedKinematics kin;
  JointStatesMap js_map;
  const double hip[4] = { 0.056, 0.056, -0.056, -0.056 };
  for (int i = 0; i < 4; i++)
  {
    LegJointStates js;
    js.q(0) = hip[i];
    js.q(1) = 0.90;
    js.q(2) = -1.94;
    js_map.emplace(leg_names[i], js);
  }
  const mat Rwb = eye(3, 3), Rwb_d = eye(3, 3);
  // a disturbed state so the QP needs working-set changes (about 10 here) like BASELINE config 2
  const vec x{ 0.01, -0.02, 0.25 }, x_d{ 0.0, 0.0, 0.26 }, xdot{ 0.3, -0.2, 0.1 }, w{ 0.2, -0.3, 0.1 }, zero{ 0.0, 0.0, 0.0 };
  const FootholdMap feet = kin.forwardKinematics(js_map);
  ForceMap f;
  TorqueMap t;
  const int iters = 3000;
  const double us_control = mean_us([&] { f = bc.control(Rwb, Rwb_d, x, xdot, w, x_d, zero, zero, feet); }, iters);
  const double us_jt = mean_us([&] { t = kin.jacobianTransposeControl(js_map, f); }, iters);
  const double us_fk = mean_us([&] { (void)kin.forwardKinematics(js_map); }, iters);
  const double us_fused = mean_us([&] { (void)bc.controlWithTorques(Rwb, Rwb_d, x, xdot, w, x_d, zero, zero, feet, js_map); }, iters);
  std::printf("{\"robots\": 1, \"iters\": %d, \"control_us\": %.2f, \"jacobianTransposeControl_us\": %.2f, \"forwardKinematics_us\": %.2f, "
              "\"controlWithTorques_us\": %.2f, \"stance_legs\": %zu, \"fz_RL\": %.6f}\n",
              iters, us_control, us_jt, us_fk, us_fused, f.size(), f.empty() ? 0.0 : f.at("RL")(2));
  return 0;
}
