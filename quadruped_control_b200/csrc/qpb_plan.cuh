// qpb_plan.cuh -- the caller code either side of the hot path, batched:
//   plan_kernel        FootPlanner::singleFoot (foot_planner.cpp:76-104) on stance->swing transitions and the sextic
//                      swing-foot reference of FootTrajectory / FootTrajectoryManager (trajectory.cpp:220-254, 300-307,
//                      323-324, 366-388), as the tick uses them at commander_node.cpp:429-461, 482-488
//   adapt_kernel       stateCallback + jointCallback + forwardKinematics (commander_node.cpp:127-187, 383-384)
//   torque_cmd_kernel  JointTorqueCmd assembly (commander_node.cpp:517-533)
// All three are HBM-bound record transforms: one thread per robot, every record touched once; on 32-byte-aligned arrays
// the adapter and the command kernel move their records with 256-bit accesses (one request per 32-byte sector).
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
  // swing legs flagged for re-planning (one bit per byte lane); the pose is only needed for those
  const uint32_t sw_b = ~(contact | (contact >> 1) | (contact >> 2) | (contact >> 3) | (contact >> 4) | (contact >> 5) |
                          (contact >> 6) | (contact >> 7)) & 0x01010101u;
  const uint32_t rp_b = (replan | (replan >> 1) | (replan >> 2) | (replan >> 3) | (replan >> 4) | (replan >> 5) |
                         (replan >> 6) | (replan >> 7)) & 0x01010101u;
  double R[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 }, x[3] = { 0, 0, 0 };
  if (sw_b & rp_b) {
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = s.Rwb[k];
#pragma unroll
    for (int k = 0; k < 3; k++) x[k] = s.x[k];
  }
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

// ---- 256-bit global accesses (sm_100a: LDG.E.256 / STG.E.256) ----------------------------------------------------
// These record transforms are bound by the number of 32-byte sectors their requests touch, not by DRAM bytes: a
// thread that walks its own record with 8- or 16-byte accesses touches every sector two to four times.  With
// 32-byte accesses on 32-byte-aligned fields each sector is touched once.
struct d4 { double a, b, c, d; };
__device__ __forceinline__ d4 ldg256(const void* p) {
  d4 v;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg256(void* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

constexpr int kXformThreads = 128;  // CTA size of adapt_kernel / torque_cmd_kernel (4 warps, one 32-robot tile each)

// One thread per robot.  V256 (all four array bases 32-byte aligned; qpb_adapt_inputs_batch checks): the 104-byte
// CoM messages of a warp's 32 robots are fetched as one contiguous 3 328-byte run of 32-byte chunks and handed to
// their owners through shared memory (stride 13 doubles: conflict-free), the joint message is six 32-byte loads,
// and the record fields are written with 32-byte stores wherever a field group covers a whole sector: 10 loads and
// 16 stores per robot instead of 25 and 28.  !V256: records only 16-byte aligned (the ABI minimum), double2 accesses.
template <bool V256>
__global__ void __launch_bounds__(kXformThreads)
adapt_kernel(const qpb_params* __restrict__ P, const qpb_com_msg* __restrict__ com,
             const qpb_joint_msg* __restrict__ joints, qpb_state_rec* __restrict__ states,
             qpb_swing_rec* __restrict__ swing, int64_t n) {
  __shared__ __align__(16) double com_s[V256 ? (kXformThreads / 32) * 32 * 13 : 2];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double cm[13];  // position[3] orientation[4] linear[3] angular[3]
  bool staged = false;
  if (V256) {
    const int lane = threadIdx.x & 31;
    const int64_t tile0 = i - lane;
    if (tile0 + 32 <= n) {  // full tile: cooperative fetch
      double* cs = com_s + (threadIdx.x >> 5) * (32 * 13);
      const char* src = reinterpret_cast<const char*>(com + tile0);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int c = lane + 32 * k;
        if (c < 104) {
          const d4 v = ldg256(src + 32 * c);
          reinterpret_cast<double2*>(cs)[2 * c] = make_double2(v.a, v.b);
          reinterpret_cast<double2*>(cs)[2 * c + 1] = make_double2(v.c, v.d);
        }
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 13; j++) cm[j] = cs[13 * lane + j];
      staged = true;
    }
  }
  if (!staged) {
    if (i >= n) return;
    const double* c = reinterpret_cast<const double*>(com + i);  // 104-byte stride: 8-byte loads
#pragma unroll
    for (int j = 0; j < 13; j++) cm[j] = __ldg(c + j);
  }
  double* sp = reinterpret_cast<double*>(&states[i]);
  double2* sv = reinterpret_cast<double2*>(sp);
  // stateCallback (commander_node.cpp:167-187): Quaternion(w, x, y, z).rotation().matrix(), i.e. Drake's
  // RotationMatrix(Eigen::Quaterniond): the 2/|q|^2 form, no normalisation of q
  const double x = cm[3], y = cm[4], z = cm[5], w = cm[6];
  const double two = 2.0 / (w * w + x * x + y * y + z * z);
  const double sx = two * x, sy = two * y, sz = two * z;
  const double swx = sx * w, swy = sy * w, swz = sz * w, sxx = sx * x, sxy = sy * x, sxz = sz * x, syy = sy * y, syz = sz * y,
               szz = sz * z;
  // Rwb = doubles 0..8 of the record; x, xdot, w = doubles 18..26
  if (V256) {
    stg256(sp, 1.0 - syy - szz, sxy - swz, sxz + swy, sxy + swz);
    stg256(sp + 4, 1.0 - sxx - szz, syz - swx, sxz - swy, syz + swx);
    sp[8] = 1.0 - sxx - syy;
    sv[9] = make_double2(cm[0], cm[1]);
    stg256(sp + 20, cm[2], cm[7], cm[8], cm[9]);
    sv[12] = make_double2(cm[10], cm[11]);
    sp[26] = cm[12];
  } else {
    sv[0] = make_double2(1.0 - syy - szz, sxy - swz);
    sv[1] = make_double2(sxz + swy, sxy + swz);
    sv[2] = make_double2(1.0 - sxx - szz, syz - swx);
    sv[3] = make_double2(sxz - swy, syz + swx);
    sp[8] = 1.0 - sxx - syy;
    sv[9] = make_double2(cm[0], cm[1]);
    sv[10] = make_double2(cm[2], cm[7]);
    sv[11] = make_double2(cm[8], cm[9]);
    sv[12] = make_double2(cm[10], cm[11]);
    sp[26] = cm[12];
  }
  // jointCallback (commander_node.cpp:127-165): message index 4 * joint + leg; forwardKinematics (kinematics.cpp:81-103)
  double pos[12], vel[12];
  if (V256) {
    const double* jp = reinterpret_cast<const double*>(&joints[i]);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const d4 a = ldg256(jp + 4 * k), b = ldg256(jp + 12 + 4 * k);
      pos[4 * k] = a.a; pos[4 * k + 1] = a.b; pos[4 * k + 2] = a.c; pos[4 * k + 3] = a.d;
      vel[4 * k] = b.a; vel[4 * k + 1] = b.b; vel[4 * k + 2] = b.c; vel[4 * k + 3] = b.d;
    }
  } else {
    const double2* jv = reinterpret_cast<const double2*>(&joints[i]);
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const double2 a = __ldg(jv + k), b = __ldg(jv + 6 + k);
      pos[2 * k] = a.x;
      pos[2 * k + 1] = a.y;
      vel[2 * k] = b.x;
      vel[2 * k + 1] = b.y;
    }
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
  if (V256) {
    double* wp = reinterpret_cast<double*>(&swing[i]) + 24;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      stg256(sp + 36 + 4 * k, feet[4 * k], feet[4 * k + 1], feet[4 * k + 2], feet[4 * k + 3]);
      stg256(sp + 48 + 4 * k, q[4 * k], q[4 * k + 1], q[4 * k + 2], q[4 * k + 3]);
      stg256(wp + 4 * k, qd[4 * k], qd[4 * k + 1], qd[4 * k + 2], qd[4 * k + 3]);
    }
  } else {
    double2* qv = reinterpret_cast<double2*>(&swing[i]) + 12;
#pragma unroll
    for (int k = 0; k < 6; k++) {
      sv[18 + k] = make_double2(feet[2 * k], feet[2 * k + 1]);
      sv[24 + k] = make_double2(q[2 * k], q[2 * k + 1]);
      qv[k] = make_double2(qd[2 * k], qd[2 * k + 1]);
    }
  }
}

// One thread per robot; the 112-byte commands of a warp's 32 robots are assembled in shared memory (compaction of
// the present legs is a dynamic shared-memory index, no local memory) and, for a full tile with 32-byte-aligned
// arrays (V256), leave as one contiguous 3 584-byte run of 32-byte stores; the torques arrive as three 32-byte loads.
template <bool V256>
__global__ void __launch_bounds__(kXformThreads)
torque_cmd_kernel(const qpb_params* __restrict__ P, const qpb_state_rec* __restrict__ states,
                  const qpb_out_rec* __restrict__ out, qpb_torque_cmd* __restrict__ cmd, int64_t n) {
  static_assert(sizeof(qpb_torque_cmd) == 112, "qpb_torque_cmd layout");
  __shared__ __align__(16) double cmd_s[(kXformThreads / 32) * 32 * 14];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int64_t tile0 = i - lane;
  if (tile0 >= n) return;  // the whole warp leaves together
  double* cs = cmd_s + (threadIdx.x >> 5) * (32 * 14);
  if (i < n) {
    const uint32_t contact = *reinterpret_cast<const uint32_t*>(states[i].contact);
    const bool qp_ok = out[i].status == QPB_OK;
    double tau[12];
    if (V256) {
      const double* tp = out[i].tau;  // byte 96 of a 256-byte record: 32-byte aligned
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const d4 v = ldg256(tp + 4 * k);
        tau[4 * k] = v.a; tau[4 * k + 1] = v.b; tau[4 * k + 2] = v.c; tau[4 * k + 3] = v.d;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 12; k++) tau[k] = out[i].tau[k];
    }
    const double lo = P->tau_min, hi = P->tau_max;
    double* mine = cs + 14 * lane;
    int cnt = 0;
    uint32_t sel = 0;  // leg of entry group s in byte s
    const int order[4] = { 1, 3, 0, 2 };  // std::map<std::string, vec3> iterates FL, FR, RL, RR (commander_node.cpp:519)
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int leg = order[k];
      const bool stance = (contact >> (8 * leg)) & 0xffu;
      // torque_map = J^T f for the stance legs the QP returned (none when it failed) + the swing-leg torques (:510-515)
      if (stance && !qp_ok) continue;
#pragma unroll
      for (int j = 0; j < 3; j++) mine[3 * cnt + j] = fmin(fmax(tau[3 * leg + j], lo), hi);  // arma::clamp, :526
      sel |= (uint32_t)leg << (8 * cnt);
      cnt++;
    }
    for (int e = 3 * cnt; e < 12; e++) mine[e] = 0.0;
    // leg[12] (entries 3s..3s+2 carry the leg of group s) and count, as the last two 8-byte words of the struct
    const uint32_t g0 = sel & 0xffu, g1 = (sel >> 8) & 0xffu, g2 = (sel >> 16) & 0xffu, g3 = sel >> 24;
    const unsigned long long w0 = (unsigned long long)(g0 * 0x010101u) | ((unsigned long long)(g1 * 0x010101u) << 24) |
                                  ((unsigned long long)(g2 * 0x0101u) << 48);
    const unsigned long long w1 = (unsigned long long)(g2 | ((g3 * 0x010101u) << 8)) | ((unsigned long long)(uint32_t)(3 * cnt) << 32);  // count = entries
    mine[12] = __longlong_as_double((long long)w0);
    mine[13] = __longlong_as_double((long long)w1);
  }
  __syncwarp();
  if (V256 && tile0 + 32 <= n) {
    char* dst = reinterpret_cast<char*>(cmd + tile0);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int c = lane + 32 * k;
      if (c < 112) {
        const double2 a = reinterpret_cast<const double2*>(cs)[2 * c], b = reinterpret_cast<const double2*>(cs)[2 * c + 1];
        stg256(dst + 32 * c, a.x, a.y, b.x, b.y);
      }
    }
  } else if (i < n) {
    double2* dst = reinterpret_cast<double2*>(cmd + i);  // 112-byte stride: 16-byte aligned
    const double* mine = cs + 14 * lane;
#pragma unroll
    for (int k = 0; k < 7; k++) dst[k] = make_double2(mine[2 * k], mine[2 * k + 1]);
  }
}

}  // namespace qpb
