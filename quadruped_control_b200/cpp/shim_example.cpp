// shim_example.cpp -- the reference's call sequence (commander_node.cpp:337-338, 383-384, 507-512)
// written against the GPU-backed shim.  Prints the BASELINE config-1 result (4-foot stance pose) as
// one JSON line; tests/test_cpp_shim.py compares it with the CPU oracle.
#include <balance_controller.hpp>

#include <cstdio>

using namespace quadruped_controller;

int main()
{
  const std::vector<std::string> leg_names = { "RL", "FL", "RR", "FR" };  // commander_node.cpp:61
  // mit_cheetah_config.yaml:66-99
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

  const BalanceController balance_controller(0.8, 11.0, 10.0, 120.0, Ib, S, W, kff, kp_p, kd_p, kp_w, kd_w, leg_names);
  const QuadrupedKinematics kinematics;

  JointStatesMap joint_states_map;  // gait_visualizer.yaml:47-50
  const double hip[4] = { 0.056, 0.056, -0.056, -0.056 };
  for (int i = 0; i < 4; i++)
  {
    LegJointStates js;
    js.q(0) = hip[i];
    js.q(1) = 0.90;
    js.q(2) = -1.94;
    joint_states_map.emplace(leg_names[i], js);
  }
  const mat Rwb = eye(3, 3), Rwb_d = eye(3, 3);
  const vec x{ 0.0, 0.0, 0.2429 }, x_d{ 0.0, 0.0, 0.26 }, zero{ 0.0, 0.0, 0.0 };

  const FootholdMap foot_actual_map = kinematics.forwardKinematics(joint_states_map);
  GaitMap gait_map = make_stance_gait();
  gait_map.at("FR").first = LegState::swing;  // exercise the stance-only return
  const ForceMap force_map = balance_controller.control(Rwb, Rwb_d, x, zero, zero, x_d, zero, zero, foot_actual_map, gait_map);
  const TorqueMap torque_map = kinematics.jacobianTransposeControl(joint_states_map, force_map);
  const ForceMap force_all = balance_controller.control(Rwb, Rwb_d, x, zero, zero, x_d, zero, zero, foot_actual_map);

  std::printf("{");
  auto dump = [](const char* key, const std::map<std::string, vec3>& m, bool last) {
    std::printf("\"%s\": {", key);
    size_t k = 0;
    for (const auto& [name, v] : m)
      std::printf("\"%s\": [%.17g, %.17g, %.17g]%s", name.c_str(), v(0), v(1), v(2), (++k < m.size()) ? ", " : "");
    std::printf("}%s", last ? "" : ", ");
  };
  dump("feet", foot_actual_map, false);
  dump("force_3stance", force_map, false);
  dump("torque_3stance", torque_map, false);
  dump("force_4stance", force_all, true);
  std::printf("}\n");
  return (force_map.size() == 3 && torque_map.size() == 3 && force_all.size() == 4) ? 0 : 1;
}
