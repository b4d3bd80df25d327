/*
 * plan_oracle.h -- CPU oracle of the foothold planner and the swing-foot trajectory and of the
 * message adapters (rank 3).  TEST INFRASTRUCTURE ONLY (see qpb_oracle.h).
 *
 * Restates, in the reference's operation order (paths relative to /root/reference/quadruped_controller/):
 *   src/quadruped_controller/foot_planner.cpp:22-42, 76-104      FootPlanner::singleFoot (Raibert heuristic + LIP term)
 *   src/quadruped_controller/trajectory.cpp:220-225, 256-296     FootTrajectory::generateTrajetory (7x7 solve)
 *   src/quadruped_controller/trajectory.cpp:227-254              FootTrajectory::trackTrajectory
 *   src/quadruped_controller/trajectory.cpp:300-307, 323-324, 366-388  FootTrajectoryManager (phase -> t, centre point)
 *   src/commander_node.cpp:436-461                               the caller: p_start = Rwb * foot + x, re-plan on stance->swing
 *   src/commander_node.cpp:127-187                               jointCallback / stateCallback (message -> controller inputs)
 *   src/commander_node.cpp:517-533                               JointTorqueCmd assembly (std::map order, clamp)
 * Pinned against the reference's own foot_planner.cpp and trajectory.cpp compiled in oracle/_ref
 * (tests/test_plan.py).  Third-party arithmetic restated, not pinned: arma::solve (LAPACK dgesv: LU with partial
 * pivoting) and Drake's RotationMatrix(Eigen::Quaterniond) used by stateCallback.
 */
#ifndef PLAN_ORACLE_H
#define PLAN_ORACLE_H

#include <stdint.h>

#include "qpb_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_plan_params {
  double k_raibert;        /* FootPlanner::k_, foot_planner.cpp:26 */
  double g;                /* foot_planner.cpp:22 */
  double thigh_offset[12]; /* base -> thigh per leg RL FL RR FR, foot_planner.cpp:28-42 */
  double height, t_swing, t_stance; /* mit_cheetah_config.yaml:17-19 */
} orc_plan_params;

typedef struct orc_plan_rec { /* 240 bytes */
  double p_start[12], p_final[12]; /* FootTrajBounds per leg, world frame (types.hpp) */
  double phase[4];                 /* GaitMap[leg].second */
  uint8_t replan[4];               /* the leg switched stance -> swing this tick (FootPlanner::updateStates) */
  uint8_t pad[12];
} orc_plan_rec;

void orc_default_plan_params(orc_plan_params* p);
void orc_single_foot(const orc_plan_params* p, int leg, const orc_state* s, double foothold[3]);
/* 7x3 coefficients, row-major: a[k][axis]; returns 0 on success */
int orc_foot_trajectory(const double p_start[3], const double p_center[3], const double p_final[3], double coef[21]);
void orc_track_trajectory(const double coef[21], double t, double pos[3], double vel[3]);
/* one robot: re-plan the flagged swing legs (plan record updated, flags cleared) and write the reference foot states of
 * every swing leg into sw->foot_ref_pos / foot_ref_vel; stance legs are left untouched */
void orc_plan(const orc_plan_params* p, const orc_state* s, orc_plan_rec* plan, orc_swing* sw);
void orc_plan_batch(const orc_plan_params* p, const orc_state* s, orc_plan_rec* plan, orc_swing* sw, int64_t n);

/* ---- message adapters (rank 3) ---------------------------------------------------------------------------- */
typedef struct orc_com_msg { /* quadruped_msgs/CoMState flattened in message field order */
  double position[3];
  double orientation[4]; /* geometry_msgs/Quaternion: x y z w */
  double linear[3], angular[3];
} orc_com_msg;
typedef struct orc_joint_msg { /* sensor_msgs/JointState position/velocity in joint_names order (config yaml:35-37) */
  double position[12], velocity[12];
} orc_joint_msg;
/* stateCallback + jointCallback + forwardKinematics (commander_node.cpp:383-384): fills Rwb, x, xdot, w, q, feet of the
 * state record and qdot of the swing record; the other fields are left as they are */
void orc_adapt_inputs(const orc_params* p, const orc_com_msg* com, const orc_joint_msg* js, orc_state* s, orc_swing* sw);
/* JointTorqueCmd.torque in the order the reference emits it: legs in std::map order (FL FR RL RR), 3 joints each,
 * clamped to [tau_min, tau_max] (commander_node.cpp:517-533); present[leg]==0 legs are skipped.  Returns the count. */
int orc_torque_cmd(const orc_params* p, const double tau[12], const int present[4], double torque[12], int leg_of_entry[12]);

#ifdef __cplusplus
}
#endif
#endif
