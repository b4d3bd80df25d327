// qpb_stages.h -- the per-QP stages every balance kernel shares, written once (host + device):
// the PD target with the reference's quirks, the dynamics right-hand side, the leg Jacobian.
// Each reference quirk (balance_controller.cpp:129, 139, 269) lives here and nowhere else.
#pragma once

#include "../../include/qpb200.h"
#include "qpb_math.h"

namespace qpb {

// slots of the 64-double state record (qpb_state_rec)
enum : int { kR = 0, kRd = 9, kX = 18, kXdot = 21, kW = 24, kXd = 27, kXdotd = 30, kWd = 33, kFeet = 36, kQ = 48, kContact = 60 };

// Right-hand side of the dynamics b = [m (a + g); Iw alpha + w_d x (Iw w_d)]:
// PD target balance_controller.cpp:126-139, Iw = R Ib R^T :251, b :265-270.  rec = the 60 doubles of a state record.
template <class Params>
QPB_HD void pd_rhs(const Params& P, const double* rec, double (&b6)[6]) {
  const double* R = rec + kR;
  double acc[3];
#pragma unroll
  for (int i = 0; i < 3; i++)
    acc[i] = P.kp_p[i] * (rec[kXd + i] - rec[kX + i]) + P.kd_p[i] * (rec[kXdotd + i] - rec[kXdot + i]);
  acc[0] += P.kff[0] * rec[kXdotd];
  acc[1] += P.kff[1] * rec[kXdotd + 1];
  acc[2] += P.kff[2] * P.mass * 9.81;  // quirk :129: gravity feed-forward, not the desired z velocity
  const double g[3] = { 0.0, 0.0, -9.81 };
#pragma unroll
  for (int i = 0; i < 3; i++) b6[i] = P.mass * (acc[i] + g[i]);  // :265 (sign of g as the reference has it)
  double Re[9], aa[3], wd[3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)  // R_d * R^T, :133
      Re[3 * i + j] = rec[kRd + 3 * i] * R[3 * j] + rec[kRd + 3 * i + 1] * R[3 * j + 1] + rec[kRd + 3 * i + 2] * R[3 * j + 2];
  angle_axis_total(Re, aa);
#pragma unroll
  for (int i = 0; i < 3; i++) wd[i] = P.kp_w[i] * aa[i] + P.kd_w[i] * (rec[kWd + i] - rec[kW + i]);
  wd[0] += P.kff[3] * rec[kWd];
  wd[1] += P.kff[4] * rec[kWd + 1];
  wd[1] += P.kff[5] * rec[kWd + 2];  // index 1 twice: reference quirk, :139
  double RI[9], Iw[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) RI[3 * i + j] = R[3 * i] * P.Ib[j] + R[3 * i + 1] * P.Ib[3 + j] + R[3 * i + 2] * P.Ib[6 + j];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) Iw[3 * i + j] = RI[3 * i] * R[3 * j] + RI[3 * i + 1] * R[3 * j + 1] + RI[3 * i + 2] * R[3 * j + 2];
  const double w0 = rec[kWd], w1 = rec[kWd + 1], w2 = rec[kWd + 2];  // desired omega in the gyroscopic term, quirk :269
  double Iwd[3], Iww[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    Iwd[i] = Iw[3 * i] * wd[0] + Iw[3 * i + 1] * wd[1] + Iw[3 * i + 2] * wd[2];
    Iww[i] = Iw[3 * i] * w0 + Iw[3 * i + 1] * w1 + Iw[3 * i + 2] * w2;
  }
  b6[3] = Iwd[0] + (w1 * Iww[2] - w2 * Iww[1]);
  b6[4] = Iwd[1] + (w2 * Iww[0] - w0 * Iww[2]);
  b6[5] = Iwd[2] + (w0 * Iww[1] - w1 * Iww[0]);
}

// Column ax of the leg Jacobian (kinematics.cpp:175-185) from the sines/cosines of t1, t2, t2+t3.
QPB_HD void leg_jacobian_col(int ax, double l1, double l2, double l3, double s1, double c1, double s2, double c2, double s23,
                             double c23, double& Jx, double& Jy, double& Jz) {
  if (ax == 0) {
    Jx = 0.0;
    Jy = -l1 * s1 - l2 * c1 * c2 - l3 * c1 * c23;
    Jz = l1 * c1 - l2 * s1 * c2 - l3 * s1 * c23;
  } else if (ax == 1) {
    const double h = l2 * s2 + l3 * s23;
    Jx = l2 * c2 + l3 * c23;
    Jy = h * s1;
    Jz = -h * c1;
  } else {
    Jx = l3 * c23;
    Jy = l3 * s1 * s23;
    Jz = -l3 * s23 * c1;
  }
}

// tau = J(q)^T f for one leg (kinematics.cpp:162-188, 218-231)
QPB_HD void leg_jt(double l1, double l2, double l3, double s1, double c1, double s2, double c2, double s23, double c23,
                   double fx, double fy, double fz, double (&tau)[3]) {
  const double h = l2 * s2 + l3 * s23;
  tau[0] = (-l1 * s1 - l2 * c1 * c2 - l3 * c1 * c23) * fy + (l1 * c1 - l2 * s1 * c2 - l3 * s1 * c23) * fz;
  tau[1] = (l2 * c2 + l3 * c23) * fx + h * s1 * fy - h * c1 * fz;
  tau[2] = l3 * c23 * fx + l3 * s1 * s23 * fy - l3 * s23 * c1 * fz;
}

}  // namespace qpb
