/*
 * mpc_oracle.c -- CPU oracle of the 10-step convex-MPC ground-reaction-force QP (BASELINE config 4).
 *
 * TEST INFRASTRUCTURE ONLY (see qpb_oracle.h): used by tests/, __graft_entry__.smoke() and bench.py's
 * CPU-baseline legs as the checker, never by the product path.
 *
 * PARITY UNPINNED: the reference has NO code for this path.  README.md:22-26 describes only the
 * instantaneous balance QP (balance_controller.cpp); SURVEY.md 8f row 2 adopts the condensed
 * single-rigid-body MPC of Di Carlo et al., "Dynamic Locomotion in the MIT Cheetah 3 Through Convex
 * Model-Predictive Control" (IROS 2018), with this repo's friction pyramid rows (the 5-row Cf of
 * balance_controller.cpp:278-289) and bounds (balance_controller.cpp:296-301, 312-321) per foot and step.
 *
 * This file builds the condensed QP the literal way -- simulating the linear time-varying model column by
 * column into dense A_qp (130x13) and B_qp (130x120) -- and solves it with the dense Goldfarb-Idnani
 * solver of qpb_oracle.c in the reference's two-sided row form (200 rows, swing feet as zero equalities).
 * The CUDA path derives the same Hessian in closed form and solves in a whitened operator form, so the two
 * share neither the assembly nor the factorisation.
 *
 * Model (state x = [roll pitch yaw | p | omega | v | g] in R^13, input u_k = 4 world-frame foot forces):
 *   psi_k   = xref[k][2]                          yaw the step is linearised at
 *   T_k     = Rz(psi_k)^T                         Euler-rate map for small roll/pitch
 *   I_k     = Rz(psi_k) Ib Rz(psi_k)^T            world inertia
 *   Theta'  = Theta + dt T_k omega
 *   p'      = p + dt v
 *   omega'  = omega + dt sum_i I_k^-1 (r_ki x f_i)
 *   v'      = v + dt (sum_i f_i / m + e_z g)
 *   g'      = g                                   (x[12] = -9.81 in the record)
 *   cost    = sum_k (x_{k+1} - xref[k])' diag(Lw) (x_{k+1} - xref[k]) + alpha |u_k|^2
 *   QP      = min 1/2 U'HU + gvec'U,  H = 2 (B'LB + alpha I),  gvec = 2 B'L (A x0 - Xref)
 */
#include "mpc_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

void orc_mpc_default_params(orc_mpc_params* p) {
  memset(p, 0, sizeof(*p));
  p->mu = 0.6;
  p->mass = 11.0;   /* mit_cheetah_config.yaml:96 */
  p->fzmin = 10.0;  /* :98 */
  p->fzmax = 120.0; /* :99 */
  p->Ib[0] = 0.011253; p->Ib[4] = 0.036203; p->Ib[8] = 0.042673; /* :95 */
  p->dt = 0.03;
  const double Lw[13] = { 0.25, 0.25, 10.0, 2.0, 2.0, 50.0, 0.0, 0.0, 0.3, 0.2, 0.2, 0.1, 0.0 };
  memcpy(p->Lw, Lw, sizeof(Lw));
  p->alpha = 4e-5;
  p->max_iter = 1000;
}

static void inv3(const double M[9], double out[9]) {
  const double a = M[0], b = M[1], c = M[2], d = M[3], e = M[4], f = M[5], g = M[6], h = M[7], i = M[8];
  const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
  out[0] = (e * i - f * h) / det; out[1] = (c * h - b * i) / det; out[2] = (b * f - c * e) / det;
  out[3] = (f * g - d * i) / det; out[4] = (a * i - c * g) / det; out[5] = (c * d - a * f) / det;
  out[6] = (d * h - e * g) / det; out[7] = (b * g - a * h) / det; out[8] = (a * e - b * d) / det;
}

static void mm3(const double A[9], const double B[9], double C[9]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double t = 0.0;
      for (int k = 0; k < 3; k++) t += A[3 * i + k] * B[3 * k + j];
      C[3 * i + j] = t;
    }
}

/* one step of the homogeneous model: x <- A_k x */
static void step_free(const orc_mpc_params* p, double psi, double x[13]) {
  const double c = cos(psi), s = sin(psi), dt = p->dt;
  const double w0 = x[6], w1 = x[7], w2 = x[8];
  x[0] += dt * (c * w0 + s * w1);
  x[1] += dt * (-s * w0 + c * w1);
  x[2] += dt * w2;
  for (int i = 0; i < 3; i++) x[3 + i] += dt * x[9 + i];
  x[11] += dt * x[12];
}

/* B_k e_a for input a = 3*foot + comp: the state increment a unit force produces in one step */
static void input_column(const orc_mpc_params* p, double psi, const double r[3], int comp, double col[13]) {
  const double c = cos(psi), s = sin(psi);
  const double Rz[9] = { c, -s, 0, s, c, 0, 0, 0, 1 }, RzT[9] = { c, s, 0, -s, c, 0, 0, 0, 1 };
  double t[9], Iw[9], Iinv[9];
  mm3(Rz, p->Ib, t);
  mm3(t, RzT, Iw);
  inv3(Iw, Iinv);
  /* r x e_comp */
  double e[3] = { 0, 0, 0 }, rx[3];
  e[comp] = 1.0;
  rx[0] = r[1] * e[2] - r[2] * e[1];
  rx[1] = r[2] * e[0] - r[0] * e[2];
  rx[2] = r[0] * e[1] - r[1] * e[0];
  memset(col, 0, 13 * sizeof(double));
  for (int i = 0; i < 3; i++)
    col[6 + i] = p->dt * (Iinv[3 * i] * rx[0] + Iinv[3 * i + 1] * rx[1] + Iinv[3 * i + 2] * rx[2]);
  col[9 + comp] = p->dt / p->mass;
}

void orc_mpc_assemble(const orc_mpc_params* p, const orc_mpc_rec* s, double* H, double* gvec, double* C, double* lb,
                      double* ub) {
  enum { NH = ORC_MPC_H, NV = ORC_MPC_NV, NX = 13 };
  double* B = (double*)calloc((size_t)NH * NX * NV, sizeof(double)); /* B[(k*13+s)*120 + a] */
  double free_resp[NH * NX];
  double x[13];
  memcpy(x, s->x0, sizeof(x));
  for (int k = 0; k < NH; k++) {
    step_free(p, s->xref[k][2], x);
    memcpy(&free_resp[k * NX], x, sizeof(x));
  }
  for (int j = 0; j < NH; j++)
    for (int foot = 0; foot < 4; foot++)
      for (int comp = 0; comp < 3; comp++) {
        const int a = 12 * j + 3 * foot + comp;
        double col[13];
        input_column(p, s->xref[j][2], s->r[j][foot], comp, col);
        for (int i = 0; i < NX; i++) B[(j * NX + i) * NV + a] = col[i];
        for (int k = j + 1; k < NH; k++) {
          step_free(p, s->xref[k][2], col);
          for (int i = 0; i < NX; i++) B[(k * NX + i) * NV + a] = col[i];
        }
      }
  /* H = 2 (B'LB + alpha I), gvec = 2 B'L (A x0 - Xref) */
  for (int a = 0; a < NV; a++) {
    for (int b = 0; b <= a; b++) {
      double t = 0.0;
      for (int k = 0; k < NH; k++)
        for (int i = 0; i < NX; i++) t += B[(k * NX + i) * NV + a] * p->Lw[i] * B[(k * NX + i) * NV + b];
      if (a == b) t += p->alpha;
      H[a * NV + b] = H[b * NV + a] = 2.0 * t;
    }
    double t = 0.0;
    for (int k = 0; k < NH; k++)
      for (int i = 0; i < NX; i++) t += B[(k * NX + i) * NV + a] * p->Lw[i] * (free_resp[k * NX + i] - s->xref[k][i]);
    gvec[a] = 2.0 * t;
  }
  free(B);
  /* rows: balance_controller.cpp:278-289 per foot and step; bounds :296-301 (stance), :312-316 (swing) */
  memset(C, 0, sizeof(double) * ORC_MPC_NC * NV);
  for (int k = 0; k < NH; k++)
    for (int foot = 0; foot < 4; foot++) {
      const int v = 12 * k + 3 * foot, row = 5 * (4 * k + foot);
      C[(row + 0) * NV + v + 0] = 1.0; C[(row + 0) * NV + v + 2] = -p->mu;
      C[(row + 1) * NV + v + 1] = 1.0; C[(row + 1) * NV + v + 2] = -p->mu;
      C[(row + 2) * NV + v + 1] = 1.0; C[(row + 2) * NV + v + 2] = p->mu;
      C[(row + 3) * NV + v + 0] = 1.0; C[(row + 3) * NV + v + 2] = p->mu;
      C[(row + 4) * NV + v + 2] = 1.0;
      if (s->contact[k][foot]) {
        const double l[5] = { -1e6, -1e6, 0.0, 0.0, p->fzmin }, u[5] = { 0.0, 0.0, 1e6, 1e6, p->fzmax };
        memcpy(&lb[row], l, sizeof(l));
        memcpy(&ub[row], u, sizeof(u));
      } else {
        for (int i = 0; i < 5; i++) lb[row + i] = ub[row + i] = 0.0;
      }
    }
}

int orc_mpc_solve(const orc_mpc_params* p, const orc_mpc_rec* s, orc_mpc_out* out) {
  enum { NV = ORC_MPC_NV, NC = ORC_MPC_NC };
  memset(out, 0, sizeof(*out));
  const double* in = (const double*)s;
  for (int i = 0; i < 13 + 130 + 120; i++)
    if (!isfinite(in[i])) { out->status = 2; return 2; }
  double* ws = (double*)malloc(sizeof(double) * (NV * NV + NV + NC * NV + 2 * NC + NV));
  double *H = ws, *g = H + NV * NV, *C = g + NV, *lb = C + NC * NV, *ub = lb + NC, *U = ub + NC;
  orc_mpc_assemble(p, s, H, g, C, lb, ub);
  int iters = 0;
  const int st = orc_qp_solve(NV, NC, H, g, C, lb, ub, 100000, U, NULL, &iters);
  out->status = st;
  out->iters = iters;
  if (st == 0) memcpy(out->U, U, sizeof(double) * NV);
  free(ws);
  return st;
}

typedef struct {
  const orc_mpc_params* p;
  const orc_mpc_rec* s;
  orc_mpc_out* out;
  int64_t lo, hi;
} mpc_job;

static void* mpc_worker(void* arg) {
  mpc_job* j = (mpc_job*)arg;
  for (int64_t i = j->lo; i < j->hi; i++) orc_mpc_solve(j->p, &j->s[i], &j->out[i]);
  return NULL;
}

void orc_mpc_batch(const orc_mpc_params* p, const orc_mpc_rec* s, int64_t n, orc_mpc_out* out, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  mpc_job jobs[256];
  for (int t = 0; t < nthreads; t++) {
    jobs[t].p = p; jobs[t].s = s; jobs[t].out = out;
    jobs[t].lo = n * t / nthreads;
    jobs[t].hi = n * (t + 1) / nthreads;
  }
  for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, mpc_worker, &jobs[t]);
  mpc_worker(&jobs[0]);
  for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
}
