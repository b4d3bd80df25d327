/* plan_oracle.c -- see plan_oracle.h.  TEST INFRASTRUCTURE ONLY. */
#include "plan_oracle.h"

#include <math.h>
#include <string.h>

void orc_default_plan_params(orc_plan_params* p) {
  memset(p, 0, sizeof(*p));
  p->k_raibert = 0.01;
  p->g = 9.81;
  const double xbt = 0.196, ybt = 0.127, zbt = 0.0;
  const double sx[4] = { -1.0, 1.0, -1.0, 1.0 }, sy[4] = { 1.0, 1.0, -1.0, -1.0 }; /* RL FL RR FR */
  for (int leg = 0; leg < 4; leg++) {
    p->thigh_offset[3 * leg] = sx[leg] * xbt;
    p->thigh_offset[3 * leg + 1] = sy[leg] * ybt;
    p->thigh_offset[3 * leg + 2] = zbt;
  }
  p->height = 0.08;
  p->t_swing = 0.18;
  p->t_stance = 0.8;
}

static void matvec3(const double R[9], const double v[3], double out[3]) {
  for (int i = 0; i < 3; i++) out[i] = R[3 * i] * v[0] + R[3 * i + 1] * v[1] + R[3 * i + 2] * v[2];
}

/* foot_planner.cpp:76-104 */
void orc_single_foot(const orc_plan_params* p, int leg, const orc_state* s, double foothold[3]) {
  double p_thigh[3], pcom_foot[3];
  matvec3(s->Rwb, &p->thigh_offset[3 * leg], p_thigh);
  for (int i = 0; i < 3; i++) p_thigh[i] += s->x[i];
  matvec3(s->Rwb, &s->feet[3 * leg], pcom_foot);
  const double tang[3] = { s->w[1] * pcom_foot[2] - s->w[2] * pcom_foot[1], s->w[2] * pcom_foot[0] - s->w[0] * pcom_foot[2],
                           s->w[0] * pcom_foot[1] - s->w[1] * pcom_foot[0] };
  const double half = p->t_stance / 2.0, lip = 0.5 * sqrt(s->x[2] / p->g);
  for (int i = 0; i < 3; i++) {
    const double p_linear = half * s->xdot[i] + p->k_raibert * (s->xdot[i] - s->xdot_d[i]);
    const double p_tangent = half * tang[i];
    const double p_lip = lip * s->xdot[i];
    foothold[i] = ((p_thigh[i] + p_linear) + p_tangent) + p_lip;
  }
  foothold[2] = 0.0;
}

/* trajectory.cpp:220-225 with initSystem (:256-277) and constantTerms (:279-296); arma::solve(fast) = LU, partial pivoting */
int orc_foot_trajectory(const double p_start[3], const double p_center[3], const double p_final[3], double coef[21]) {
  double A[7][7] = { { 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0 },
                     { 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0 },
                     { 1.0, 0.5, pow(0.5, 2), pow(0.5, 3), pow(0.5, 4), pow(0.5, 5), pow(0.5, 6) },
                     { 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0 },
                     { 0.0, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0 },
                     { 0.0, 0.0, 2.0, 0.0, 0.0, 0.0, 0.0 },
                     { 0.0, 0.0, 2.0, 6.0, 12.0, 20.0, 30.0 } };
  double B[7][3];
  memset(B, 0, sizeof(B));
  for (int j = 0; j < 3; j++) { B[0][j] = p_start[j]; B[1][j] = p_final[j]; B[2][j] = p_center[j]; }
  for (int c = 0; c < 7; c++) {
    int piv = c;
    for (int r = c + 1; r < 7; r++)
      if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
    if (A[piv][c] == 0.0) return 1;
    if (piv != c) {
      for (int j = 0; j < 7; j++) { const double t = A[c][j]; A[c][j] = A[piv][j]; A[piv][j] = t; }
      for (int j = 0; j < 3; j++) { const double t = B[c][j]; B[c][j] = B[piv][j]; B[piv][j] = t; }
    }
    for (int r = c + 1; r < 7; r++) {
      const double f = A[r][c] / A[c][c];
      A[r][c] = f;
      for (int j = c + 1; j < 7; j++) A[r][j] -= f * A[c][j];
    }
  }
  for (int col = 0; col < 3; col++) {
    for (int r = 1; r < 7; r++)
      for (int j = 0; j < r; j++) B[r][col] -= A[r][j] * B[j][col];
    for (int r = 6; r >= 0; r--) {
      for (int j = r + 1; j < 7; j++) B[r][col] -= A[r][j] * B[j][col];
      B[r][col] /= A[r][r];
    }
  }
  for (int k = 0; k < 7; k++)
    for (int j = 0; j < 3; j++) coef[3 * k + j] = B[k][j];
  return 0;
}

/* trajectory.cpp:227-254 */
void orc_track_trajectory(const double coef[21], double t, double pos[3], double vel[3]) {
  const double pf[7] = { 1.0, t, pow(t, 2), pow(t, 3), pow(t, 4), pow(t, 5), pow(t, 6) };
  const double vf[7] = { 0.0, 1.0, 2.0 * t, 3.0 * pow(t, 2), 4.0 * pow(t, 3), 5.0 * pow(t, 4), 6.0 * pow(t, 5) };
  for (int j = 0; j < 3; j++) {
    double a = 0.0, b = 0.0;
    for (int k = 0; k < 7; k++) { a += pf[k] * coef[3 * k + j]; b += vf[k] * coef[3 * k + j]; }
    pos[j] = a;
    vel[j] = b;
  }
}

void orc_plan(const orc_plan_params* p, const orc_state* s, orc_plan_rec* plan, orc_swing* sw) {
  /* trajectory.cpp:300-307 */
  const double stance_phase = p->t_stance / (p->t_swing + p->t_stance);
  const double slope = 1.0 / (1.0 - stance_phase), y_intercept = 1.0 - slope;
  for (int leg = 0; leg < 4; leg++) {
    if (s->contact[leg]) continue; /* stance: nothing planned, nothing referenced (trajectory.cpp:353-361) */
    if (plan->replan[leg]) {
      /* commander_node.cpp:436-461: new foothold, start point = current foot in the world frame */
      orc_single_foot(p, leg, s, &plan->p_final[3 * leg]);
      double ps[3];
      matvec3(s->Rwb, &s->feet[3 * leg], ps);
      for (int i = 0; i < 3; i++) plan->p_start[3 * leg + i] = ps[i] + s->x[i];
      plan->replan[leg] = 0;
    }
    const double* p0 = &plan->p_start[3 * leg];
    const double* pf = &plan->p_final[3 * leg];
    double pc[3] = { (p0[0] + pf[0]) / 2.0, (p0[1] + pf[1]) / 2.0, p->height }; /* trajectory.cpp:323-324 */
    double coef[21];
    if (orc_foot_trajectory(p0, pc, pf, coef) != 0) memset(coef, 0, sizeof(coef));
    double t = slope * plan->phase[leg] + y_intercept; /* trajectory.cpp:373 */
    t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
    orc_track_trajectory(coef, t, &sw->foot_ref_pos[3 * leg], &sw->foot_ref_vel[3 * leg]);
  }
}

void orc_plan_batch(const orc_plan_params* p, const orc_state* s, orc_plan_rec* plan, orc_swing* sw, int64_t n) {
  for (int64_t i = 0; i < n; i++) orc_plan(p, &s[i], &plan[i], &sw[i]);
}

/* ---- message adapters ------------------------------------------------------------------------------------- */
void orc_adapt_inputs(const orc_params* p, const orc_com_msg* com, const orc_joint_msg* js, orc_state* s, orc_swing* sw) {
  /* stateCallback, commander_node.cpp:167-187: Quaternion(w, x, y, z).rotation().matrix() -> Drake
   * RotationMatrix(Eigen::Quaterniond): R = I + (2/|q|^2)(...) without normalising q first */
  const double x = com->orientation[0], y = com->orientation[1], z = com->orientation[2], w = com->orientation[3];
  const double two_over_norm_squared = 2.0 / (w * w + x * x + y * y + z * z);
  const double sx = two_over_norm_squared * x, sy = two_over_norm_squared * y, sz = two_over_norm_squared * z;
  const double swx = sx * w, swy = sy * w, swz = sz * w;
  const double sxx = sx * x, sxy = sy * x, sxz = sz * x;
  const double syy = sy * y, syz = sz * y, szz = sz * z;
  s->Rwb[0] = 1.0 - syy - szz; s->Rwb[1] = sxy - swz;       s->Rwb[2] = sxz + swy;
  s->Rwb[3] = sxy + swz;       s->Rwb[4] = 1.0 - sxx - szz; s->Rwb[5] = syz - swx;
  s->Rwb[6] = sxz - swy;       s->Rwb[7] = syz + swx;       s->Rwb[8] = 1.0 - sxx - syy;
  for (int i = 0; i < 3; i++) { s->x[i] = com->position[i]; s->xdot[i] = com->linear[i]; s->w[i] = com->angular[i]; }
  /* jointCallback, commander_node.cpp:127-165: message index = 4 * joint + leg (RL FL RR FR) */
  for (int leg = 0; leg < 4; leg++)
    for (int j = 0; j < 3; j++) {
      s->q[3 * leg + j] = js->position[4 * j + leg];
      sw->qdot[3 * leg + j] = js->velocity[4 * j + leg];
    }
  /* foot_actual_map = forwardKinematics(joint_states_map), commander_node.cpp:383-384 */
  for (int leg = 0; leg < 4; leg++) orc_forward_kinematics(p, leg, &s->q[3 * leg], &s->feet[3 * leg]);
}

int orc_torque_cmd(const orc_params* p, const double tau[12], const int present[4], double torque[12], int leg_of_entry[12]) {
  static const int map_order[4] = { 1, 3, 0, 2 }; /* std::map<std::string, ...>: "FL" < "FR" < "RL" < "RR" */
  int n = 0;
  for (int k = 0; k < 4; k++) {
    const int leg = map_order[k];
    if (!present[leg]) continue;
    for (int j = 0; j < 3; j++) {
      double t = tau[3 * leg + j];
      t = t < p->tau_min ? p->tau_min : (t > p->tau_max ? p->tau_max : t); /* arma::clamp, :526 */
      torque[n] = t;
      leg_of_entry[n] = leg;
      n++;
    }
  }
  return n;
}
