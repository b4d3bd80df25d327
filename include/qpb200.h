/*
 * qpb200.h -- C ABI of libqpb200.so: the batched, B200-native replacement of the reference's
 * balance-controller hot path.
 *
 * The reference (bostoncleek/quadruped_control) has no FFI layer; the boundary a maintainer binds is
 * the C++ class API below (paths relative to quadruped_controller/):
 *   BalanceController::BalanceController(...)               include/quadruped_controller/balance_controller.hpp:85-88
 *   ForceMap BalanceController::control(...) const          include/quadruped_controller/balance_controller.hpp:104-107
 *   TorqueMap QuadrupedKinematics::jacobianTransposeControl include/quadruped_controller/kinematics.hpp:106-107
 *   FootholdMap QuadrupedKinematics::forwardKinematics      include/quadruped_controller/kinematics.hpp:60
 * called once per control tick at src/commander_node.cpp:337-338 (ctor), 383-384 (FK) and 507-512.
 * quadruped_control_b200/cpp/balance_controller.hpp keeps those C++ signatures on top of this ABI.
 *
 * Conventions: plain pointers and sizes, no exceptions, every call returns 0 or a negative
 * qpb_error; all arithmetic FP64; matrices row-major; legs in the reference's order RL, FL, RR, FR
 * (commander_node.cpp:61); contact 1 = stance, 0 = swing (types.hpp:91-95).  There is NO CPU
 * fallback: without a CUDA device every compute entry point fails with QPB_ERR_CUDA.
 */
#ifndef QPB200_H
#define QPB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QPB_VERSION 200

typedef enum qpb_error {
  QPB_SUCCESS = 0,
  QPB_ERR_INVALID_ARG = -1, /* null pointer, negative n, misaligned record pointer */
  QPB_ERR_BAD_PARAMS = -2,  /* parameters outside what the solver proves correct (see qpb_create) */
  QPB_ERR_CUDA = -3,        /* CUDA runtime error; text in qpb_last_error() */
  QPB_ERR_NO_MEMORY = -4
} qpb_error;

/* per-QP status written next to each result */
enum { QPB_OK = 0, QPB_MAX_ITER = 1, QPB_BAD_INPUT = 2 };

/* Constructor arguments of BalanceController (balance_controller.hpp:85-88; stored at
 * balance_controller.cpp:70-96), the kinematic constants QuadrupedKinematics hard-codes
 * (kinematics.cpp:23-47) and the caller's torque clamp (commander_node.cpp:324-325, 526). */
typedef struct qpb_params {
  double mu, mass, fzmin, fzmax;
  double Ib[9];   /* body inertia, 3x3 */
  double S[36];   /* least-squares weight, 6x6 symmetric positive definite */
  double W[144];  /* force regulariser, 12x12 symmetric positive definite */
  double kff[6], kp_p[3], kd_p[3], kp_w[3], kd_w[3];
  double hip_offset[12]; /* base->hip translation per leg */
  double link[12];       /* signed (l1, l2, l3) per leg: +l1 left legs, -l1 right legs, -l2, -l3 */
  double tau_min, tau_max;
  int32_t clamp_tau; /* 0 (default): return J^T f unclamped, as jacobianTransposeControl does */
  int32_t max_iter;  /* working-set changes allowed per QP; nWSR_ = 200, balance_controller.cpp:85 */
} qpb_params;

/* One robot state = the argument list of control() + the joint angles of
 * jacobianTransposeControl().  512 bytes, 16-byte aligned: one coalesced load per warp. */
typedef struct qpb_state_rec {
  double Rwb[9], Rwb_d[9];
  double x[3], xdot[3], w[3], x_d[3], xdot_d[3], w_d[3];
  double feet[12]; /* body-frame foot positions (FootholdMap), leg-major */
  double q[12];    /* joint angles (JointStatesMap.q), leg-major: hip, thigh, calf */
  uint8_t contact[4];
  /* pad[0..3]: optional warm start, the counterpart of SQProblem::hotstart between ticks (balance_controller.cpp:177-202).
   * A little-endian uint32: bit 31 set = bits 0..23 hold the working set qpb_out_rec.pad[0..3] reported for this robot on
   * the previous tick (copy the four bytes over); 0 = cold start.  The hint only changes the number of working-set
   * changes, never the result (the optimum is unique); a stale or malformed hint falls back to the cold start.
   * Honoured when W = w I and fzmin >= 0 (the range-space kernels); the general-W kernels ignore it.
   * pad[4..27]: ignored. */
  uint8_t pad[28];
} qpb_state_rec;

/* ForceMap + TorqueMap flattened.  Swing legs (absent from the reference's maps,
 * balance_controller.cpp:223-228) are zero.  When status != QPB_OK every force and torque is
 * zero, which is the reference's "empty map" failure return (balance_controller.cpp:182-216). */
typedef struct qpb_out_rec {
  double grf_body[12];
  double tau[12];
  int32_t status;
  int32_t iters; /* working-set changes used */
  /* pad[0..3]: the working set at the optimum as a little-endian uint32 with bit 31 set (2 bits per leg and row group:
   * 0 none, 1 / 2 = which side of |fx| <= mu fz, |fy| <= mu fz, fzmin <= fz <= fzmax is active), to be fed back as the
   * next tick's warm start (every kernel reports it).  pad[4..55]: zero. */
  uint8_t pad[56];
} qpb_out_rec;

typedef struct qpb_handle qpb_handle;

int qpb_version(void);
/* Text of the last error raised on the calling thread ("" if none). */
const char* qpb_last_error(void);

/* mit_cheetah_config.yaml:66-99 with W = 1e-5*I (commander_node.cpp:289, 305). */
int qpb_default_params(qpb_params* out);

/* Replaces the BalanceController + QuadrupedKinematics constructors (commander_node.cpp:337-338, 358).
 * Picks the kernels once: W = w*I with fzmin >= 0 (the reference's configuration, commander_node.cpp:289, 305) takes the
 * range-space path (three launches: set-up, active-set loop, polish + epilogue -- for device-resident calls the second and
 * third are programmatic dependent launches, so the passes follow each other without an idle gap; cold batches below
 * 12 288 records and general W take the one-launch half-warp kernel, warm-started batches and single robots the one-launch
 * range-space kernel).  QPB_QPS_PER_WARP=1|2|32, QPB_TPQ_LPQ=1|2|4, QPB_TPQ_MIN_N and QPB_TPQ_PDL=0|1 in the environment
 * override the choices (experiments, tests).
 * Rejects (QPB_ERR_BAD_PARAMS): non-finite values, mu <= 0, fzmin > fzmax, fzmax < 0,
 * S or W not symmetric positive definite, max_iter < 1, and 2*mu*fzmax > 1e6 (the reference's
 * finite "far" bounds of +-1e6, balance_controller.cpp:296-297, are provably inactive below that
 * and the solver does not carry them). */
int qpb_create(const qpb_params* params, int device, qpb_handle** out);
int qpb_destroy(qpb_handle* h);

/* control() + jacobianTransposeControl() for n robots; packed records resident on the device.
 * stream is a cudaStream_t (NULL = default stream).  Asynchronous and capturable in a CUDA graph: the work counters a
 * launch draws from are re-armed by the launch itself, and its scratch comes from the stream-ordered allocator
 * (cudaMallocAsync).  At most 4096 launches of one handle may be in flight at a time. */
int qpb_control_batch_packed(qpb_handle* h, int64_t n, const qpb_state_rec* d_states, qpb_out_rec* d_out,
                             void* stream);

/* Same, argument-per-array (device pointers), mirroring the control() parameter list
 * (balance_controller.hpp:104-107): Rwb, Rwb_d [n*9]; x, xdot, w, x_d, xdot_d, w_d [n*3];
 * feet_body [n*12]; contact [n*4]; q [n*12] -> grf_body [n*12], tau [n*12], status [n].
 * tau and status may be NULL. */
int qpb_control_batch(qpb_handle* h, int64_t n, const double* Rwb, const double* Rwb_d, const double* x,
                      const double* xdot, const double* w, const double* x_d, const double* xdot_d,
                      const double* w_d, const double* feet_body, const uint8_t* contact, const double* q,
                      double* grf_body, double* tau, int32_t* status, void* stream);

/* Host-buffer entry point; returns when h_out is complete.  The records are copied to the device in stages, solved
 * and copied back, overlapping the three on internal streams (pinned buffers from qpb_host_alloc make the copies
 * asynchronous); with the one-launch kernels, pinned buffers are instead read and written by the kernel itself over
 * PCIe.  The kernels are chosen by the size of the whole batch, so a record's result does not depend on its stage. */
int qpb_control_batch_host(qpb_handle* h, int64_t n, const qpb_state_rec* h_states, qpb_out_rec* h_out);

/* Asynchronous form for callers that step many batches (a simulator bridge): queues the uploads, kernels and downloads
 * of this batch on the handle's internal streams and returns; h_out is complete after qpb_host_sync().  Several calls
 * may be in flight, so the upload of batch k+1 overlaps the download of batch k (PCIe is full duplex).  The buffers must
 * stay valid and untouched until the sync; pass pinned memory (qpb_host_alloc), pageable buffers make the copies
 * synchronous.  Not re-entrant: one calling thread per handle, as for BalanceController::control (mutable solver state,
 * balance_controller.hpp:161, 171-176). */
int qpb_control_batch_host_async(qpb_handle* h, int64_t n, const qpb_state_rec* h_states, qpb_out_rec* h_out);
int qpb_host_sync(qpb_handle* h);

/* Wire format of the host-buffer calls: the same fields without the padding that rounds the device records up to 512 and
 * 256 bytes -- 488 B up and 200 B down per robot instead of 768 B in all.  A host-buffer call is bound by the PCIe link
 * (bench.py `e2e.pcie_bound_qps`), so the bytes on the wire are its speed; the records are widened to qpb_state_rec and
 * narrowed from qpb_out_rec on the device, one small kernel either side of the solve. */
typedef struct qpb_wire_state {
  double Rwb[9], Rwb_d[9];
  double x[3], xdot[3], w[3], x_d[3], xdot_d[3], w_d[3];
  double feet[12];
  double q[12];
  uint8_t contact[4];
  uint32_t warm; /* = qpb_state_rec.pad[0..3]: last tick's qpb_wire_out.wset, or 0 */
} qpb_wire_state;  /* 488 bytes */

typedef struct qpb_wire_out {
  double grf_body[12];
  double tau[12];
  int16_t status;
  int16_t iters; /* saturates at 32767 */
  uint32_t wset; /* = qpb_out_rec.pad[0..3] */
} qpb_wire_out;    /* 200 bytes */

/* qpb_control_batch_host / _async on wire records; same results, field for field (tests/test_gpu_parity.py). */
int qpb_control_batch_wire_host(qpb_handle* h, int64_t n, const qpb_wire_state* h_states, qpb_wire_out* h_out);
int qpb_control_batch_wire_host_async(qpb_handle* h, int64_t n, const qpb_wire_state* h_states, qpb_wire_out* h_out);

/* Tell the handle that the records of its DEVICE-resident calls (qpb_control_batch_packed, qpb_tick_batch_packed) carry
 * warm-start words in pad[0..3] -- a controller in its loop, every tick handing the previous tick's qpb_out_rec.pad[0..3]
 * back, which is how the reference uses SQProblem::hotstart (balance_controller.cpp:177-202).  Such batches take a
 * one-launch kernel at every size (a warm QP is almost always optimal after its first 6x6 solve, so the three-pass
 * path's scratch traffic buys nothing).  Host-buffer calls need no flag: they look at the first record.  Records
 * without a valid word are still solved correctly, from a cold start, only slower than on the default path. */
int qpb_set_warm_batches(qpb_handle* h, int on);

/* jacobianTransposeControl() alone (kinematics.cpp:218-231): tau = J(q)^T f for stance legs,
 * 0 for swing legs.  Device pointers: q [n*12], grf_body [n*12], contact [n*4] (NULL = all stance). */
int qpb_jt_batch(qpb_handle* h, int64_t n, const double* q, const double* grf_body, const uint8_t* contact,
                 double* tau, void* stream);

/* forwardKinematics() (kinematics.cpp:81-103): q [n*12] -> feet_body [n*12]; device pointers. */
int qpb_fk_batch(qpb_handle* h, int64_t n, const double* q, double* feet_body, void* stream);

/* Host-buffer forms of the two kinematics calls (what a per-tick caller such as the C++ shim
 * uses): copy in, launch, copy out, synchronous. */
int qpb_jt_batch_host(qpb_handle* h, int64_t n, const double* q, const double* grf_body, const uint8_t* contact,
                      double* tau);
int qpb_fk_batch_host(qpb_handle* h, int64_t n, const double* q, double* feet_body);

/* ---- Swing-leg half of the control tick (the caller code around the hot path) ------------------------------
 * src/commander_node.cpp:482-505: reference foot state (world) -> body frame -> legInverseKinematics
 * (kinematics.cpp:117-160) -> legJacobianInverse (kinematics.cpp:190-204: inv, then pinv, then J^T) * velocity;
 * :503-504 + joint_controller.cpp:21-39: joint PD; :515: merged with the stance torques; :526: clamp. */
typedef struct qpb_joint_gains {
  double kff[3], kp[3], kd[3]; /* JointController(kff, kp, kd), joint_control/* in mit_cheetah_config.yaml:50-53 */
} qpb_joint_gains;

typedef struct qpb_swing_rec {
  double foot_ref_pos[12]; /* FootTrajectoryManager::referenceState(...).position per leg, WORLD frame */
  double foot_ref_vel[12]; /* ... .velocity per leg, world frame */
  double qdot[12];         /* measured joint velocities (JointStatesMap.qdot) */
} qpb_swing_rec;           /* 288 bytes; only the entries of legs in swing are read */

/* Defaults: kff = 0, kp = (40, 40, 50), kd = 1 (mit_cheetah_config.yaml:50-53). */
int qpb_set_joint_gains(qpb_handle* h, const qpb_joint_gains* gains);

/* One whole control tick for n robots (commander_node.cpp:482-533): control() + jacobianTransposeControl()
 * for the stance legs and the joint-space PD for the swing legs, merged into qpb_out_rec.tau (all 12 joints),
 * clamped when qpb_params.clamp_tau is set.  If the balance QP fails (status 1) the swing torques are still
 * returned, as the reference publishes them alone; a non-finite state (status 2) returns zeros.  Device pointers. */
int qpb_tick_batch_packed(qpb_handle* h, int64_t n, const qpb_state_rec* d_states, const qpb_swing_rec* d_swing,
                          qpb_out_rec* d_out, void* stream);
/* Same with host buffers (pipelined like qpb_control_batch_host). */
int qpb_tick_batch_host(qpb_handle* h, int64_t n, const qpb_state_rec* h_states, const qpb_swing_rec* h_swing,
                        qpb_out_rec* h_out);

/* Pinned host memory for qpb_control_batch_host callers (portable: every device can read and write it). */
int qpb_host_alloc(void** ptr, size_t bytes);
int qpb_host_free(void* ptr);

/* Number of CUDA kernels this handle has launched so far (for the benchmark's gpu_launches). */
int64_t qpb_launch_count(const qpb_handle* h);

/* ---- Single-process multi-GPU form of the host-buffer calls ---------------------------------------
 * Every QP is independent: the batch is cut into contiguous shards, shard r = records [lo_r, hi_r) with remainders
 * going to the lowest shards (the same ranges the one-process-per-GPU path uses), each solved on its own device by a
 * persistent host thread through that device's qpb_handle.  No data-path collective, no launcher.  The reference runs
 * one BalanceController per process (commander_node.cpp:337); this is what a bridge stepping a fleet on an 8-GPU box
 * binds instead of eight of them. */
typedef struct qpb_multi_handle qpb_multi_handle;
int qpb_device_count(int* count);
/* devices == NULL: one shard per visible device; otherwise devices[0..num_devices) (a device may be listed twice). */
int qpb_multi_create(const qpb_params* params, const int* devices, int num_devices, qpb_multi_handle** out);
int qpb_multi_destroy(qpb_multi_handle* m);
int qpb_multi_num_shards(const qpb_multi_handle* m);
int qpb_multi_shard_range(int64_t n, int shard, int num_shards, int64_t* lo, int64_t* hi);
int qpb_multi_set_joint_gains(qpb_multi_handle* m, const qpb_joint_gains* gains);
/* qpb_control_batch_host / qpb_tick_batch_host over all shards; return when h_out is complete.  On failure the
 * first failing shard's code is returned and qpb_last_error() names the shard and device. */
int qpb_multi_control_batch_host(qpb_multi_handle* m, int64_t n, const qpb_state_rec* h_states, qpb_out_rec* h_out);
int qpb_multi_tick_batch_host(qpb_multi_handle* m, int64_t n, const qpb_state_rec* h_states, const qpb_swing_rec* h_swing,
                              qpb_out_rec* h_out);
int64_t qpb_multi_launch_count(const qpb_multi_handle* m);

/* ---- The caller code either side of the tick -----------------------------------------
 * Foothold planner + swing-foot trajectory: FootPlanner::singleFoot (foot_planner.cpp:76-104) when a leg switches
 * stance -> swing, FootTrajectory/FootTrajectoryManager (trajectory.cpp:220-254, 300-307, 323-324, 366-388) every tick,
 * as the reference calls them at commander_node.cpp:429-461 and 482-488.  Stateless per call: what the reference keeps
 * in FootPlanner::state_map_ and FootTrajectoryManager::traj_map_ lives in the caller's qpb_plan_rec. */
typedef struct qpb_plan_params {
  double k_raibert;        /* FootPlanner::k_ = 0.01, foot_planner.cpp:26 */
  double g;                /* 9.81, foot_planner.cpp:22 */
  double thigh_offset[12]; /* base -> thigh per leg (+-0.196, +-0.127, 0), foot_planner.cpp:28-42 */
  double height, t_swing, t_stance; /* gait/height, t_swing, t_stance: mit_cheetah_config.yaml:17-19 */
} qpb_plan_params;

typedef struct qpb_plan_rec { /* 240 bytes */
  double p_start[12], p_final[12]; /* FootTrajBounds per leg, world frame (types.hpp:52-65) */
  double phase[4];                 /* GaitMap[leg].second */
  uint8_t replan[4];               /* in: 1 = the leg switched stance -> swing this tick; cleared once planned */
  uint8_t pad[12];
} qpb_plan_rec;

int qpb_default_plan_params(qpb_plan_params* out);
int qpb_set_plan_params(qpb_handle* h, const qpb_plan_params* p);
/* For every swing leg (qpb_state_rec.contact == 0): re-plan p_start / p_final if flagged, then write the reference foot
 * position and velocity (world frame) into d_swing[i].foot_ref_pos / foot_ref_vel, which qpb_tick_batch_packed consumes.
 * Stance legs are left untouched.  Device pointers; asynchronous on stream. */
int qpb_plan_batch(qpb_handle* h, int64_t n, const qpb_state_rec* d_states, qpb_plan_rec* d_plan, qpb_swing_rec* d_swing,
                   void* stream);

/* Message adapters.  quadruped_msgs/CoMState and sensor_msgs/JointState flattened in message field order; the
 * adapter does what stateCallback, jointCallback (commander_node.cpp:127-187) and forwardKinematics (:383-384) do:
 * fills Rwb, x, xdot, w, q, feet of the state record and qdot of the swing record (other fields are left as they are). */
typedef struct qpb_com_msg {
  double position[3];
  double orientation[4]; /* geometry_msgs/Quaternion: x y z w */
  double linear[3], angular[3];
} qpb_com_msg;
typedef struct qpb_joint_msg {
  double position[12], velocity[12]; /* joint_names order: 4 hips, 4 thighs, 4 calves, each RL FL RR FR (yaml:35-37) */
} qpb_joint_msg;
int qpb_adapt_inputs_batch(qpb_handle* h, int64_t n, const qpb_com_msg* d_com, const qpb_joint_msg* d_joints,
                           qpb_state_rec* d_states, qpb_swing_rec* d_swing, void* stream);

/* quadruped_msgs/JointTorqueCmd.torque as commander_node.cpp:517-533 fills it: legs in std::map order (FL FR RL RR),
 * stance legs only when the QP returned forces, three joints each, clamped to [tau_min, tau_max]; leg[k] names the leg
 * of entry k (stands in for actuator_name), count = number of entries. */
typedef struct qpb_torque_cmd {
  double torque[12];
  uint8_t leg[12];
  int32_t count;
} qpb_torque_cmd;
int qpb_torque_cmd_batch(qpb_handle* h, int64_t n, const qpb_state_rec* d_states, const qpb_out_rec* d_out,
                         qpb_torque_cmd* d_cmd, void* stream);

/* ---- 10-step convex-MPC ground-reaction-force QP (BASELINE.json config 4) --------------
 * The reference has NO code for this path (README.md:22-26 describes only the instantaneous QP of
 * balance_controller.cpp), so there is no reference interface to cite: the formulation is the condensed
 * single-rigid-body MPC of Di Carlo et al. (IROS 2018) with the reference's friction-pyramid rows
 * (balance_controller.cpp:278-289) and bounds (:296-301 stance, :312-316 swing) on every foot and step.
 *   state x = [roll pitch yaw | p | omega | v | g] (13), input u_k = 4 world-frame foot forces (RL FL RR FR)
 *   psi_k = xref[k][2];  T_k = Rz(psi_k)^T;  I_k = Rz(psi_k) Ib Rz(psi_k)^T
 *   Theta' = Theta + dt T_k omega;  p' = p + dt v;  omega' = omega + dt sum_i I_k^-1 (r_ki x f_i);
 *   v' = v + dt (sum_i f_i / mass + e_z g);  g' = g
 *   min sum_k (x_{k+1} - xref[k])' diag(Lw) (x_{k+1} - xref[k]) + alpha |u_k|^2
 *   s.t. per foot and step: |fx| <= mu fz, |fy| <= mu fz, fzmin <= fz <= fzmax (stance) or f = 0 (swing). */
typedef struct qpb_mpc_params {
  double mu, mass, fzmin, fzmax;
  double Ib[9];
  double dt;
  double Lw[13]; /* state weights (>= 0); Lw[12] (gravity state) is ignored */
  double alpha;  /* force weight (> 0: keeps the QP strictly convex) */
  int32_t max_iter; /* working-set changes allowed per QP */
  int32_t pad;
} qpb_mpc_params;

typedef struct qpb_mpc_rec { /* 2176 bytes, 16-byte aligned */
  double x0[13];
  double xref[10][13]; /* reference for x_1 .. x_10 */
  double r[10][4][3];  /* foot position minus CoM, world frame, per step and foot */
  uint8_t contact[10][4]; /* 1 = stance, 0 = swing */
  uint8_t pad[32];
} qpb_mpc_rec;

typedef struct qpb_mpc_out_rec { /* 1024 bytes */
  double U[120];  /* world-frame forces, step-major, then foot, then xyz; swing feet 0; all 0 when status != QPB_OK */
  int32_t status; /* QPB_OK / QPB_MAX_ITER (iteration or working-set capacity limit) / QPB_BAD_INPUT */
  int32_t iters;
  uint8_t pad[56];
} qpb_mpc_out_rec;

typedef struct qpb_mpc_handle qpb_mpc_handle;

/* Robot constants of mit_cheetah_config.yaml:95-99, mu = 0.6, dt = 0.03, Lw and alpha of Di Carlo et al. */
int qpb_mpc_default_params(qpb_mpc_params* out);
/* Rejects non-finite values, mu <= 0, fzmin > fzmax, fzmax < 0, 2*mu*fzmax > 1e6, mass <= 0, dt <= 0, alpha <= 0,
 * negative weights, singular Ib, max_iter < 1. */
int qpb_mpc_create(const qpb_mpc_params* params, int device, qpb_mpc_handle** out);
int qpb_mpc_destroy(qpb_mpc_handle* h);
/* n QPs, records resident on the device; asynchronous on stream. */
int qpb_mpc_batch_packed(qpb_mpc_handle* h, int64_t n, const qpb_mpc_rec* d_recs, qpb_mpc_out_rec* d_out, void* stream);
/* Host buffers (pinned or pageable); returns when h_out is complete. */
int qpb_mpc_batch_host(qpb_mpc_handle* h, int64_t n, const qpb_mpc_rec* h_recs, qpb_mpc_out_rec* h_out);
int64_t qpb_mpc_launch_count(const qpb_mpc_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* QPB200_H */
