// qpb_swing.cuh -- swing-leg half of the control tick, one thread per robot.
#pragma once

#include "qpb_kernel.cuh"

// ------------------------------------------------------------------------------------------------
// Swing-leg half of the control tick.
//   commander_node.cpp:482-505  reference foot state -> body frame -> IK -> J^-1 v
//   kinematics.cpp:117-160 (legInverseKinematics), :190-204 (legJacobianInverse: inv -> pinv -> J^T)
//   joint_controller.cpp:21-39 (joint PD), commander_node.cpp:515 (merge), :526 (clamp)
// ------------------------------------------------------------------------------------------------
namespace qpb {

__device__ __forceinline__ double wrap_2pi(double a) {  // math/numerics.cpp:23-35
  const double PI = 3.14159265358979323846;
  const double qf = floor(a / (2.0 * PI));
  a -= qf * 2.0 * PI;
  if (a < 0.0) a += 2.0 * PI;
  return a;
}
__device__ __forceinline__ double wrap_pi(double r) {  // math/numerics.cpp:37-50
  const double PI = 3.14159265358979323846;
  const double qf = floor((r + PI) / (2.0 * PI));
  r = (r + PI) - qf * 2.0 * PI;
  if (r < 0) r += 2.0 * PI;
  return r - PI;
}

// Moore-Penrose pseudo-inverse of a 3x3 via a cyclic Jacobi eigen-decomposition of J^T J
// (arma::pinv: SVD with tolerance max(m,n) * sigma_max * eps).  Only reached when a pivot is exactly zero.
__device__ void pinv3(const double (&J)[9], double (&out)[9]) {
  double B[9], V[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      B[3 * i + j] = J[i] * J[j] + J[3 + i] * J[3 + j] + J[6 + i] * J[6 + j];
      V[3 * i + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 30; sweep++) {
    if (fabs(B[1]) + fabs(B[2]) + fabs(B[5]) == 0.0) break;
#pragma unroll
    for (int pi = 0; pi < 2; pi++)
#pragma unroll
      for (int qi = pi + 1; qi < 3; qi++) {
        const double apq = B[3 * pi + qi];
        if (apq == 0.0) continue;
        const double theta = (B[4 * qi] - B[4 * pi]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const double bkp = B[3 * k + pi], bkq = B[3 * k + qi];
          B[3 * k + pi] = c * bkp - sn * bkq;
          B[3 * k + qi] = sn * bkp + c * bkq;
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const double bpk = B[3 * pi + k], bqk = B[3 * qi + k];
          B[3 * pi + k] = c * bpk - sn * bqk;
          B[3 * qi + k] = sn * bpk + c * bqk;
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const double vkp = V[3 * k + pi], vkq = V[3 * k + qi];
          V[3 * k + pi] = c * vkp - sn * vkq;
          V[3 * k + qi] = sn * vkp + c * vkq;
        }
      }
  }
  double w[3] = { fmax(B[0], 0.0), fmax(B[4], 0.0), fmax(B[8], 0.0) };
  const double smax = sqrt(fmax(w[0], fmax(w[1], w[2])));
  const double tol = 3.0 * smax * 2.220446049250313e-16;
  double M[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 3; k++)
        if (sqrt(w[k]) > tol) acc += V[3 * i + k] * V[3 * j + k] / w[k];
      M[3 * i + j] = acc;
    }
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      out[3 * i + j] = M[3 * i] * J[3 * j] + M[3 * i + 1] * J[3 * j + 1] + M[3 * i + 2] * J[3 * j + 2];
}

// inverse by Gauss-Jordan with partial pivoting; an exactly zero pivot falls through to the pseudo-inverse
__device__ void inv3_or_pinv(const double (&J)[9], double (&out)[9]) {
  double a[9], b[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
#pragma unroll
  for (int i = 0; i < 9; i++) a[i] = J[i];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    int piv = c;
#pragma unroll
    for (int r = c + 1; r < 3; r++)
      if (fabs(a[3 * r + c]) > fabs(a[3 * piv + c])) piv = r;
    const double pv = a[3 * piv + c];
    if (pv == 0.0 || !isfinite(pv)) {
      pinv3(J, out);
      return;
    }
#pragma unroll
    for (int r = c + 1; r < 3; r++)
      if (r == piv) {
#pragma unroll
        for (int j = 0; j < 3; j++) {
          double t = a[3 * c + j]; a[3 * c + j] = a[3 * r + j]; a[3 * r + j] = t;
          t = b[3 * c + j]; b[3 * c + j] = b[3 * r + j]; b[3 * r + j] = t;
        }
      }
    const double dd = a[3 * c + c];
#pragma unroll
    for (int j = 0; j < 3; j++) { a[3 * c + j] /= dd; b[3 * c + j] /= dd; }
#pragma unroll
    for (int r = 0; r < 3; r++)
      if (r != c) {
        const double f = a[3 * r + c];
#pragma unroll
        for (int j = 0; j < 3; j++) { a[3 * r + j] -= f * a[3 * c + j]; b[3 * r + j] -= f * b[3 * c + j]; }
      }
  }
#pragma unroll
  for (int i = 0; i < 9; i++) out[i] = b[i];
}

// One thread per ROBOT: the thread lists its swing legs and walks the list, so in a warp the j-th trip is taken
// by every robot with more than j swing legs (91 % / 55 % of the lanes on the mixed-contact workload) instead of
// 36 % of the lanes when each leg had its own thread.
__global__ void swing_kernel(const qpb_params* __restrict__ P, const qpb_joint_gains* __restrict__ G,
                             const qpb_state_rec* __restrict__ states, const qpb_swing_rec* __restrict__ swing,
                             qpb_out_rec* __restrict__ out, int64_t nrobots) {
  const int64_t rob = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (rob >= nrobots) return;
  const qpb_state_rec& s = states[rob];
  const uint32_t cbytes = *reinterpret_cast<const uint32_t*>(s.contact);
  uint32_t todo = ((cbytes & 0xffu) ? 0u : 1u) | ((cbytes & 0xff00u) ? 0u : 2u) | ((cbytes & 0xff0000u) ? 0u : 4u) |
                  ((cbytes & 0xff000000u) ? 0u : 8u);   // legs in swing; stance legs keep the balance controller's torque
  if (todo == 0u) return;
  if (out[rob].status == QPB_BAD_INPUT) return;  // nothing is commanded for a broken state
  for (; todo != 0u; todo &= todo - 1u) {
    const int leg = __ffs(todo) - 1;
    const qpb_swing_rec& sw = swing[rob];
    double pb[3], vb[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {  // commander_node.cpp:491-492: Rwb' p - x (sic), Rwb' v
      pb[a] = s.Rwb[a] * sw.foot_ref_pos[3 * leg] + s.Rwb[3 + a] * sw.foot_ref_pos[3 * leg + 1] +
              s.Rwb[6 + a] * sw.foot_ref_pos[3 * leg + 2] - s.x[a];
      vb[a] = s.Rwb[a] * sw.foot_ref_vel[3 * leg] + s.Rwb[3 + a] * sw.foot_ref_vel[3 * leg + 1] +
              s.Rwb[6 + a] * sw.foot_ref_vel[3 * leg + 2];
    }
    // legInverseKinematics, kinematics.cpp:117-160
    const double x = pb[0] - P->hip_offset[3 * leg], y = pb[1] - P->hip_offset[3 * leg + 1], z = pb[2] - P->hip_offset[3 * leg + 2];
    const double sl1 = P->link[3 * leg], sl2 = P->link[3 * leg + 1], sl3 = P->link[3 * leg + 2];
    const double l1 = fabs(sl1), l2 = fabs(sl2), l3 = fabs(sl3);
    double d = (x * x + y * y + z * z - l1 * l1 - l2 * l2 - l3 * l3) / (2.0 * l2 * l3);
    if (d > 1.0) d = 1.0;
    double sc = y * y + z * z - l1 * l1;
    if (sc < 0.0) sc = 0.0;
    const double rsc = sqrt(sc);
    double q0;
    if (sl1 < 0.0) q0 = atan2(z, y) + atan2(rsc, -l1);  // right legs
    else q0 = -(atan2(z, -y) + atan2(rsc, -l1));
    const double q2 = atan2(-sqrt(1.0 - d * d), d);
    double s3, c3;
    sincos(q2, &s3, &c3);
    const double q1 = -atan2(x, rsc) - atan2(l3 * s3, l2 + l3 * c3);
    // legJacobian at the reference pose, kinematics.cpp:162-188
    double s1, c1, s2, c2, s23, c23;
    sincos(q0, &s1, &c1);
    sincos(q1, &s2, &c2);
    sincos(q1 + q2, &s23, &c23);
    const double h = sl2 * s2 + sl3 * s23;
    const double J[9] = { 0.0, sl2 * c2 + sl3 * c23, sl3 * c23,
                          -sl1 * s1 - sl2 * c1 * c2 - sl3 * c1 * c23, h * s1, sl3 * s1 * s23,
                          sl1 * c1 - sl2 * s1 * c2 - sl3 * s1 * c23, -h * c1, -sl3 * s23 * c1 };
    double Ji[9];
    inv3_or_pinv(J, Ji);  // legJacobianInverse, kinematics.cpp:190-204
    const double qr[3] = { q0, q1, q2 };
#pragma unroll
    for (int a = 0; a < 3; a++) {  // JointController::control, joint_controller.cpp:27-35
      const double qdr = Ji[3 * a] * vb[0] + Ji[3 * a + 1] * vb[1] + Ji[3 * a + 2] * vb[2];
      const double qe = wrap_pi(wrap_2pi(qr[a]) - wrap_2pi(s.q[3 * leg + a]));
      double tau = G->kp[a] * qe + G->kd[a] * (qdr - sw.qdot[3 * leg + a]) + G->kff[a];
      if (P->clamp_tau) tau = fmin(fmax(tau, P->tau_min), P->tau_max);  // commander_node.cpp:526
      out[rob].tau[3 * leg + a] = tau;
    }
  }
}

}  // namespace qpb
