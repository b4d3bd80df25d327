// qpb_plan.cuh -- the caller code either side of the hot path, batched:
//   plan_kernel        FootPlanner::singleFoot (foot_planner.cpp:76-104) on stance->swing transitions and the sextic
//                      swing-foot reference of FootTrajectory / FootTrajectoryManager (trajectory.cpp:220-254, 300-307,
//                      323-324, 366-388), as the tick uses them at commander_node.cpp:429-461, 482-488
//   adapt_kernel       stateCallback + jointCallback + forwardKinematics (commander_node.cpp:127-187, 383-384)
//   torque_cmd_kernel  JointTorqueCmd assembly (commander_node.cpp:517-533)
// All three are HBM-bound record transforms: one thread per robot, every record touched once.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/qpb200.h"

namespace qpb {

// The 7x7 system of trajectory.cpp:256-296 has the closed-form solution a0 = p0, a1 = a2 = 0 and, with
// D = p_final - p_start, C = p_center - p_start:  a3 = 64C - 22D, a4 = -192C + 81D, a5 = 192C - 90D, a6 = -64C + 32D
// (the reference solves it numerically by LU each time; the results agree to rounding).
__device__ __forceinline__ void sextic_eval(double p0, double pc, double pf, double t, double& pos, double& vel) {
  const double D = pf - p0, C = pc - p0;
  const double a3 = 64.0 * C - 22.0 * D, a4 = -192.0 * C + 81.0 * D, a5 = 192.0 * C - 90.0 * D, a6 = -64.0 * C + 32.0 * D;
  const double t2 = t * t, t3 = t2 * t;
  pos = p0 + t3 * (a3 + t * (a4 + t * (a5 + t * a6)));
  vel = t2 * (3.0 * a3 + t * (4.0 * a4 + t * (5.0 * a5 + t * (6.0 * a6))));
}

__global__ void plan_kernel(const qpb_plan_params* __restrict__ PP, const qpb_state_rec* __restrict__ states,
                            qpb_plan_rec* __restrict__ plans, qpb_swing_rec* __restrict__ swing, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const qpb_state_rec& s = states[i];
  const uint32_t contact = *reinterpret_cast<const uint32_t*>(s.contact);
  if ((contact & 0x01010101u) == 0x01010101u) return;  // all four feet in stance: nothing to plan or reference
  qpb_plan_rec& pl = plans[i];
  const uint32_t replan = *reinterpret_cast<const uint32_t*>(pl.replan);
  const double stance_phase = PP->t_stance / (PP->t_swing + PP->t_stance);  // trajectory.cpp:300-307
  const double slope = 1.0 / (1.0 - stance_phase), y0 = 1.0 - slope;
  double R[9], x[3];
#pragma unroll
  for (int k = 0; k < 9; k++) R[k] = s.Rwb[k];
#pragma unroll
  for (int k = 0; k < 3; k++) x[k] = s.x[k];
  uint32_t cleared = replan;
  for (int leg = 0; leg < 4; leg++) {
    if ((contact >> (8 * leg)) & 0xffu) continue;
    double p0[3], pf[3];
    if ((replan >> (8 * leg)) & 0xffu) {
      // commander_node.cpp:436-461 + foot_planner.cpp:76-104
      const double f0 = s.feet[3 * leg], f1 = s.feet[3 * leg + 1], f2 = s.feet[3 * leg + 2];
      const double h0 = PP->thigh_offset[3 * leg], h1 = PP->thigh_offset[3 * leg + 1], h2 = PP->thigh_offset[3 * leg + 2];
      double pcf[3], pth[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        pcf[k] = R[3 * k] * f0 + R[3 * k + 1] * f1 + R[3 * k + 2] * f2;
        pth[k] = (R[3 * k] * h0 + R[3 * k + 1] * h1 + R[3 * k + 2] * h2) + x[k];
      }
      const double w0 = s.w[0], w1 = s.w[1], w2 = s.w[2];
      const double tang[3] = { w1 * pcf[2] - w2 * pcf[1], w2 * pcf[0] - w0 * pcf[2], w0 * pcf[1] - w1 * pcf[0] };
      const double half = PP->t_stance / 2.0, lip = 0.5 * sqrt(x[2] / PP->g);
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const double xd = s.xdot[k];
        const double lin = half * xd + PP->k_raibert * (xd - s.xdot_d[k]);
        pf[k] = ((pth[k] + lin) + half * tang[k]) + lip * xd;
        p0[k] = pcf[k] + x[k];
      }
      pf[2] = 0.0;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        pl.p_start[3 * leg + k] = p0[k];
        pl.p_final[3 * leg + k] = pf[k];
      }
      cleared &= ~(0xffu << (8 * leg));
    } else {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        p0[k] = pl.p_start[3 * leg + k];
        pf[k] = pl.p_final[3 * leg + k];
      }
    }
    double t = slope * pl.phase[leg] + y0;  // trajectory.cpp:373
    t = fmin(fmax(t, 0.0), 1.0);
    const double pc[3] = { (p0[0] + pf[0]) / 2.0, (p0[1] + pf[1]) / 2.0, PP->height };  // trajectory.cpp:323-324
#pragma unroll
    for (int k = 0; k < 3; k++) {
      double pos, vel;
      sextic_eval(p0[k], pc[k], pf[k], t, pos, vel);
      swing[i].foot_ref_pos[3 * leg + k] = pos;
      swing[i].foot_ref_vel[3 * leg + k] = vel;
    }
  }
  if (cleared != replan) *reinterpret_cast<uint32_t*>(pl.replan) = cleared;
}

// One thread per robot.  The records are 16-byte aligned (ABI contract) and every field group written here starts on a
// 16-byte boundary, so the state / swing records and the joint message are moved as double2 (half the L2 requests of
// 8-byte accesses; the kernel is request-bound, not DRAM-bound).
__global__ void adapt_kernel(const qpb_params* __restrict__ P, const qpb_com_msg* __restrict__ com,
                             const qpb_joint_msg* __restrict__ joints, qpb_state_rec* __restrict__ states,
                             qpb_swing_rec* __restrict__ swing, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const qpb_com_msg& c = com[i];  // 104-byte stride: 8-byte loads
  double2* sv = reinterpret_cast<double2*>(&states[i]);
  // stateCallback (commander_node.cpp:167-187): Quaternion(w, x, y, z).rotation().matrix(), i.e. Drake's
  // RotationMatrix(Eigen::Quaterniond): the 2/|q|^2 form, no normalisation of q
  const double x = c.orientation[0], y = c.orientation[1], z = c.orientation[2], w = c.orientation[3];
  const double two = 2.0 / (w * w + x * x + y * y + z * z);
  const double sx = two * x, sy = two * y, sz = two * z;
  const double swx = sx * w, swy = sy * w, swz = sz * w, sxx = sx * x, sxy = sy * x, sxz = sz * x, syy = sy * y, syz = sz * y,
               szz = sz * z;
  // Rwb = doubles 0..8 of the record
  sv[0] = make_double2(1.0 - syy - szz, sxy - swz);
  sv[1] = make_double2(sxz + swy, sxy + swz);
  sv[2] = make_double2(1.0 - sxx - szz, syz - swx);
  sv[3] = make_double2(sxz - swy, syz + swx);
  states[i].Rwb[8] = 1.0 - sxx - syy;
  // x, xdot, w = doubles 18..26
  sv[9] = make_double2(c.position[0], c.position[1]);
  sv[10] = make_double2(c.position[2], c.linear[0]);
  sv[11] = make_double2(c.linear[1], c.linear[2]);
  sv[12] = make_double2(c.angular[0], c.angular[1]);
  states[i].w[2] = c.angular[2];
  // jointCallback (commander_node.cpp:127-165): message index 4 * joint + leg; forwardKinematics (kinematics.cpp:81-103)
  const double2* jv = reinterpret_cast<const double2*>(&joints[i]);
  double pos[12], vel[12];
#pragma unroll
  for (int k = 0; k < 6; k++) {
    const double2 a = __ldg(jv + k), b = __ldg(jv + 6 + k);
    pos[2 * k] = a.x;
    pos[2 * k + 1] = a.y;
    vel[2 * k] = b.x;
    vel[2 * k + 1] = b.y;
  }
  double q[12], feet[12], qd[12];
#pragma unroll
  for (int leg = 0; leg < 4; leg++) {
    const double t1 = pos[leg], t2 = pos[4 + leg], t3 = pos[8 + leg];
    q[3 * leg] = t1;
    q[3 * leg + 1] = t2;
    q[3 * leg + 2] = t3;
    qd[3 * leg] = vel[leg];
    qd[3 * leg + 1] = vel[4 + leg];
    qd[3 * leg + 2] = vel[8 + leg];
    const double l1 = P->link[3 * leg], l2 = P->link[3 * leg + 1], l3 = P->link[3 * leg + 2];
    double s1, c1, s2, c2, s23, c23;
    sincos(t1, &s1, &c1);
    sincos(t2, &s2, &c2);
    sincos(t2 + t3, &s23, &c23);
    feet[3 * leg] = l2 * s2 + l3 * s23 + P->hip_offset[3 * leg];
    feet[3 * leg + 1] = l1 * c1 - l2 * s1 * c2 - l3 * s1 * c23 + P->hip_offset[3 * leg + 1];
    feet[3 * leg + 2] = l1 * s1 + l2 * c1 * c2 + l3 * c1 * c23 + P->hip_offset[3 * leg + 2];
  }
  // feet = doubles 36..47, q = doubles 48..59 of the state record; qdot = doubles 24..35 of the swing record
  double2* qv = reinterpret_cast<double2*>(&swing[i]) + 12;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    sv[18 + k] = make_double2(feet[2 * k], feet[2 * k + 1]);
    sv[24 + k] = make_double2(q[2 * k], q[2 * k + 1]);
    qv[k] = make_double2(qd[2 * k], qd[2 * k + 1]);
  }
}

__global__ void torque_cmd_kernel(const qpb_params* __restrict__ P, const qpb_state_rec* __restrict__ states,
                                  const qpb_out_rec* __restrict__ out, qpb_torque_cmd* __restrict__ cmd, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t contact = *reinterpret_cast<const uint32_t*>(states[i].contact);
  const qpb_out_rec& o = out[i];
  const bool qp_ok = o.status == QPB_OK;
  qpb_torque_cmd c;
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < 12; k++) {
    c.torque[k] = 0.0;
    c.leg[k] = 0;
  }
  const int order[4] = { 1, 3, 0, 2 };  // std::map<std::string, vec3> iterates FL, FR, RL, RR (commander_node.cpp:519)
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int leg = order[k];
    const bool stance = (contact >> (8 * leg)) & 0xffu;
    // torque_map = J^T f for the stance legs the QP returned (none when it failed) + the swing-leg torques (:510-515)
    if (stance && !qp_ok) continue;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      c.torque[cnt] = fmin(fmax(o.tau[3 * leg + j], P->tau_min), P->tau_max);  // arma::clamp, :526
      c.leg[cnt] = (uint8_t)leg;
      cnt++;
    }
  }
  c.count = cnt;
  cmd[i] = c;
}

}  // namespace qpb
