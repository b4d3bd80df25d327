/*
 * qpb_oracle.h -- CPU oracle for the balance-controller hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (quadruped_control_b200/,
 * include/, libqpb200.so) may include, link or call this.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and there only as the checker / the timed CPU baseline.
 *
 * It is a plain-C FP64 restatement of the reference's algorithm
 * (citations are relative to /root/reference/quadruped_controller/):
 *   src/quadruped_controller/balance_controller.cpp:98-330   (PD, dynamics, QP form, bounds, epilogue)
 *   src/quadruped_controller/math/rigid3d.cpp:61-74, 198-203 (skew, SO(3) log via Drake/Eigen)
 *   src/quadruped_controller/kinematics.cpp:19-47, 81-103, 162-188, 218-231 (FK, Jacobian, J^T f)
 *
 * PARITY UNPINNED for the QP solve: the reference calls qpOASES (master@326a651), which
 * is not vendored under /root/reference and not installed here, and the reference ships
 * no test or golden GRF.  The QP is strictly convex (W > 0) so its minimiser is unique;
 * this oracle solves it with an independent dense Goldfarb-Idnani dual active-set method
 * (QR-updated factors) and every answer can be certified by the solver-free KKT checker
 * in oracle/kkt.py.  The kinematics ARE pinned: against the notebook outputs stored in
 * scripts/kinematics/quadruped_kinematics.ipynb (cells 5-7), see tests/golden/.
 */
#ifndef QPB_ORACLE_H
#define QPB_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Constructor arguments of BalanceController (balance_controller.hpp:85-88) plus the
 * kinematic constants of QuadrupedKinematics (kinematics.cpp:23-47) and the caller's
 * torque clamp (commander_node.cpp:324-325, 526).  Row-major matrices. Leg order
 * RL, FL, RR, FR (commander_node.cpp:61). */
typedef struct orc_params {
  double mu, mass, fzmin, fzmax;
  double Ib[9];
  double S[36];
  double W[144];
  double kff[6], kp_p[3], kd_p[3], kp_w[3], kd_w[3];
  double hip_offset[12]; /* base->hip per leg (x,y,z) */
  double link[12];       /* signed (l1,l2,l3) per leg: +l1 left legs, -l1 right legs, -l2, -l3 */
  double tau_min, tau_max;
  int32_t clamp_tau;
  int32_t max_iter; /* nWSR_, balance_controller.cpp:85 */
} orc_params;

/* One robot state: the arguments of BalanceController::control (balance_controller.hpp:104-107)
 * and the joint angles consumed by jacobianTransposeControl (kinematics.hpp:106-107). 512 bytes. */
typedef struct orc_state {
  double Rwb[9], Rwb_d[9];
  double x[3], xdot[3], w[3], x_d[3], xdot_d[3], w_d[3];
  double feet[12]; /* body-frame foot positions, leg-major */
  double q[12];    /* joint angles, leg-major (hip, thigh, calf) */
  uint8_t contact[4]; /* 1 = stance, 0 = swing (types.hpp:91-95) */
  uint8_t pad[28];
} orc_state;

typedef struct orc_out {
  double grf_body[12]; /* -Rwb^T f_w per stance leg; swing legs: 0 (absent in the reference map) */
  double tau[12];      /* J^T f_b per stance leg; swing legs: 0 */
  int32_t status;      /* 0 ok, 1 iteration limit, 2 bad input / infeasible */
  int32_t iters;
  uint8_t pad[56];
} orc_out;

void orc_default_params(orc_params* p);

/* kinematics.cpp:81-103 */
void orc_forward_kinematics(const orc_params* p, int leg, const double q[3], double foot[3]);
/* kinematics.cpp:162-188, row-major 3x3 */
void orc_leg_jacobian(const orc_params* p, int leg, const double q[3], double J[9]);
/* Drake RotationMatrix(mat).ToAngleAxis() -> axis*angle  (rigid3d.cpp:198-203) */
void orc_angle_axis_total(const double R[9], double out[3]);

/* The QP exactly as the reference hands it to qpOASES (balance_controller.cpp:119-161):
 * Q 12x12 row-major, c 12, C 20x12 row-major, lbC/ubC 20.  Also returns A (6x12), b (6). */
void orc_assemble(const orc_params* p, const orc_state* s, double Q[144], double c[12], double C[240],
                  double lbC[20], double ubC[20], double A[72], double b[6]);

/* Dense strictly-convex QP: min 1/2 x'Qx + c'x  s.t. lb <= Cx <= ub (rows with lb==ub are
 * equalities).  Returns 0 ok, 1 iteration limit, 2 infeasible / not positive definite.
 * lam (length m, may be NULL): multiplier per row, >0 at the lower side, <0 at the upper. */
int orc_qp_solve(int n, int m, const double* Q, const double* c, const double* C, const double* lb,
                 const double* ub, int max_iter, double* x, double* lam, int* iters);

/* Whole path for one state: control() then jacobianTransposeControl(). Also returns the
 * world-frame QP solution fw (12) when non-NULL. */
int orc_control(const orc_params* p, const orc_state* s, orc_out* out, double* fw);

/* Batch driver used for the CPU baseline: nthreads pthreads over contiguous slices. */
void orc_control_batch(const orc_params* p, const orc_state* s, int64_t n, orc_out* out, int nthreads);

/* ---- swing-leg half of the control tick -----------------------------------------
 * src/commander_node.cpp:482-505 (reference foot state -> body frame -> IK -> J^-1 v), :503-504 and
 * joint_controller.cpp:21-39 (joint PD), :515 (merge with the stance torques), :526 (clamp);
 * kinematics.cpp:117-160 (legInverseKinematics), :190-204 (legJacobianInverse: inv -> pinv -> J^T). */
typedef struct orc_swing {
  double foot_ref_pos[12]; /* world-frame reference foot positions (FootTrajectoryManager::referenceState) */
  double foot_ref_vel[12]; /* world-frame reference foot velocities */
  double qdot[12];         /* measured joint velocities (JointStatesMap.qdot) */
} orc_swing;
typedef struct orc_joint_gains {
  double kff[3], kp[3], kd[3]; /* joint_control/{kff,kp,kd}: mit_cheetah_config.yaml:50-53 */
} orc_joint_gains;

void orc_default_joint_gains(orc_joint_gains* g);
void orc_leg_inverse_kinematics(const orc_params* p, int leg, const double foothold[3], double q[3]);
/* returns 0 = plain inverse, 1 = pseudo-inverse (a pivot was exactly zero), row-major 3x3 */
int orc_leg_jacobian_inverse(const orc_params* p, int leg, const double q[3], double Jinv[9]);
/* torques of the legs in swing (others untouched); present[leg] = 1 where written */
void orc_swing_torques(const orc_params* p, const orc_joint_gains* g, const orc_state* s, const orc_swing* sw,
                       double tau[12], int present[4]);
/* whole tick: control() + jacobianTransposeControl() + swing torques merged, clamp if p->clamp_tau */
int orc_tick(const orc_params* p, const orc_joint_gains* g, const orc_state* s, const orc_swing* sw, orc_out* out);
void orc_tick_batch(const orc_params* p, const orc_joint_gains* g, const orc_state* s, const orc_swing* sw, int64_t n,
                    orc_out* out, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
