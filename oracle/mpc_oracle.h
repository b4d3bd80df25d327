/*
 * mpc_oracle.h -- CPU oracle of the 10-step convex-MPC QP (BASELINE config 4).
 * TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (the reference has no code for this path) -- see mpc_oracle.c.
 */
#ifndef MPC_ORACLE_H
#define MPC_ORACLE_H

#include <stdint.h>

#include "qpb_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_MPC_H = 10, ORC_MPC_NV = 120, ORC_MPC_NC = 200 };

typedef struct orc_mpc_params {
  double mu, mass, fzmin, fzmax;
  double Ib[9];
  double dt;
  double Lw[13]; /* state weights: roll pitch yaw | p | omega | v | g */
  double alpha;  /* force weight */
  int32_t max_iter;
  int32_t pad;
} orc_mpc_params;

typedef struct orc_mpc_rec { /* 2176 bytes */
  double x0[13];
  double xref[10][13];   /* reference for x_1 .. x_10 */
  double r[10][4][3];    /* foot position minus CoM, world frame, per step and foot (RL FL RR FR) */
  uint8_t contact[10][4];
  uint8_t pad[32];
} orc_mpc_rec;

typedef struct orc_mpc_out { /* 1024 bytes */
  double U[120]; /* world-frame foot forces, step-major, then foot, then xyz */
  int32_t status, iters;
  uint8_t pad[56];
} orc_mpc_out;

void orc_mpc_default_params(orc_mpc_params* p);
/* dense condensed QP: H 120x120, gvec 120, C 200x120, lb/ub 200 (row-major) */
void orc_mpc_assemble(const orc_mpc_params* p, const orc_mpc_rec* s, double* H, double* gvec, double* C, double* lb,
                      double* ub);
int orc_mpc_solve(const orc_mpc_params* p, const orc_mpc_rec* s, orc_mpc_out* out);
void orc_mpc_batch(const orc_mpc_params* p, const orc_mpc_rec* s, int64_t n, orc_mpc_out* out, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
