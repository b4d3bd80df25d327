/*
 * qpb_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see qpb_oracle.h).
 *
 * Plain C99, FP64, no dependencies.  Dense arithmetic in the operation order of the
 * reference; the QP is solved in the reference's own form (12 variables, 20 two-sided
 * rows, swing legs as zero-equality rows) by a textbook Goldfarb-Idnani dual active-set
 * method with QR-updated factors (Goldfarb & Idnani, Math. Programming 27 (1983)).
 * Citations: /root/reference/quadruped_controller/src/quadruped_controller/...
 *
 * Pinning (details in qpb_oracle.h and DESIGN.md section 5): every line the reference itself wrote on this path is pinned
 * to 3e-11 against the reference's own sources compiled in oracle/_ref; the kinematics against the notebook vectors;
 * the QP SOLVE IS PARITY UNPINNED (qpOASES is not installable here; uniqueness of the minimiser + a KKT certificate stand in).
 */
#include "qpb_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * parameters: mit_cheetah_config.yaml:66-99, commander_node.cpp:289-305, kinematics.cpp:23-47
 * ---------------------------------------------------------------------------------------- */
void orc_default_params(orc_params* p) {
  memset(p, 0, sizeof(*p));
  p->mu = 0.8;
  p->mass = 11.0;
  p->fzmin = 10.0;
  p->fzmax = 120.0;
  p->Ib[0] = 0.011253;
  p->Ib[4] = 0.036203;
  p->Ib[8] = 0.042673;
  const double sdiag[6] = { 1.0, 1.0, 1.0, 10.0, 10.0, 5.0 };
  for (int i = 0; i < 6; i++) p->S[i * 6 + i] = sdiag[i];
  for (int i = 0; i < 12; i++) p->W[i * 12 + i] = 1e-5;
  p->kff[2] = 0.15;
  for (int i = 0; i < 3; i++) {
    p->kp_p[i] = 100.0;
    p->kd_p[i] = 50.0;
    p->kp_w[i] = 5000.0;
    p->kd_w[i] = 500.0;
  }
  const double xbh = 0.196, ybh = 0.050, zbh = 0.0;
  const double l1 = 0.077, l2 = 0.211, l3 = 0.230;
  const double sx[4] = { -1, 1, -1, 1 };  /* RL FL RR FR */
  const double sy[4] = { 1, 1, -1, -1 };
  for (int leg = 0; leg < 4; leg++) {
    p->hip_offset[3 * leg + 0] = sx[leg] * xbh;
    p->hip_offset[3 * leg + 1] = sy[leg] * ybh;
    p->hip_offset[3 * leg + 2] = zbh;
    p->link[3 * leg + 0] = sy[leg] * l1; /* left legs +l1, right legs -l1 (kinematics.cpp:41-42) */
    p->link[3 * leg + 1] = -l2;
    p->link[3 * leg + 2] = -l3;
  }
  p->tau_min = -20.0;
  p->tau_max = 20.0;
  p->clamp_tau = 0;
  p->max_iter = 200;
}

/* ------------------------------------------------------------------------------------------
 * kinematics.cpp:81-103
 * ---------------------------------------------------------------------------------------- */
void orc_forward_kinematics(const orc_params* p, int leg, const double q[3], double foot[3]) {
  const double l1 = p->link[3 * leg], l2 = p->link[3 * leg + 1], l3 = p->link[3 * leg + 2];
  const double t1 = q[0], t2 = q[1], t3 = q[2];
  foot[0] = l2 * sin(t2) + l3 * sin(t2 + t3) + p->hip_offset[3 * leg];
  foot[1] = l1 * cos(t1) - l2 * sin(t1) * cos(t2) - l3 * sin(t1) * cos(t2 + t3) + p->hip_offset[3 * leg + 1];
  foot[2] = l1 * sin(t1) + l2 * cos(t1) * cos(t2) + l3 * cos(t1) * cos(t2 + t3) + p->hip_offset[3 * leg + 2];
}

/* kinematics.cpp:162-188 */
void orc_leg_jacobian(const orc_params* p, int leg, const double q[3], double J[9]) {
  const double l1 = p->link[3 * leg], l2 = p->link[3 * leg + 1], l3 = p->link[3 * leg + 2];
  const double t1 = q[0], t2 = q[1], t3 = q[2];
  J[0] = 0.0;
  J[1] = l2 * cos(t2) + l3 * cos(t2 + t3);
  J[2] = l3 * cos(t2 + t3);
  J[3] = -l1 * sin(t1) - l2 * cos(t1) * cos(t2) - l3 * cos(t1) * cos(t2 + t3);
  J[4] = (l2 * sin(t2) + l3 * sin(t2 + t3)) * sin(t1);
  J[5] = l3 * sin(t1) * sin(t2 + t3);
  J[6] = l1 * cos(t1) - l2 * sin(t1) * cos(t2) - l3 * sin(t1) * cos(t2 + t3);
  J[7] = -(l2 * sin(t2) + l3 * sin(t2 + t3)) * cos(t1);
  J[8] = -l3 * sin(t2 + t3) * cos(t1);
}

/* ------------------------------------------------------------------------------------------
 * rigid3d.cpp:177-179, 198-203: drake::math::RotationMatrix(R).ToAngleAxis() is
 * Eigen::AngleAxisd(Matrix3d): matrix -> quaternion (Shepperd's branches) -> angle-axis with
 * angle in [0, pi].  Drake/Eigen are not under /root/reference; restated from their published
 * algorithm (Eigen 3.3 Quaternion.h quaternionbase_assign_impl<Other,3,3>, AngleAxis.h operator=).
 * ---------------------------------------------------------------------------------------- */
void orc_angle_axis_total(const double R[9], double out[3]) {
  double qw, qv[3];
  double t = R[0] + R[4] + R[8];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    qw = 0.5 * t;
    t = 0.5 / t;
    qv[0] = (R[7] - R[5]) * t;
    qv[1] = (R[2] - R[6]) * t;
    qv[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    qv[i] = 0.5 * t;
    t = 0.5 / t;
    qw = (R[3 * k + j] - R[3 * j + k]) * t;
    qv[j] = (R[3 * j + i] + R[3 * i + j]) * t;
    qv[k] = (R[3 * k + i] + R[3 * i + k]) * t;
  }
  double n = sqrt(qv[0] * qv[0] + qv[1] * qv[1] + qv[2] * qv[2]);
  if (n != 0.0) {
    const double angle = 2.0 * atan2(n, fabs(qw));
    if (qw < 0.0) n = -n;
    out[0] = qv[0] / n * angle;
    out[1] = qv[1] / n * angle;
    out[2] = qv[2] / n * angle;
  } else {
    out[0] = out[1] = out[2] = 0.0; /* angle 0, axis (1,0,0) */
  }
}

/* small dense helpers, row-major */
static void matmul(int m, int k, int n, const double* A, const double* B, double* C) {
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) {
      double s = 0.0;
      for (int l = 0; l < k; l++) s += A[i * k + l] * B[l * n + j];
      C[i * n + j] = s;
    }
}
static void transpose(int m, int n, const double* A, double* At) {
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) At[j * m + i] = A[i * n + j];
}

/* ------------------------------------------------------------------------------------------
 * balance_controller.cpp:107-161 and 237-330: everything up to the qpOASES call
 * ---------------------------------------------------------------------------------------- */
void orc_assemble(const orc_params* p, const orc_state* s, double Q[144], double c[12], double C[240],
                  double lbC[20], double ubC[20], double A[72], double b[6]) {
  /* frictionConeBounds, :294-330 */
  const double upper = 1000000.0, lower = -1000000.0;
  const double lbf[5] = { lower, lower, 0.0, 0.0, p->fzmin };
  const double ubf[5] = { 0.0, 0.0, upper, upper, p->fzmax };
  for (int leg = 0; leg < 4; leg++)
    for (int r = 0; r < 5; r++) {
      const int swing = (s->contact[leg] == 0);
      lbC[5 * leg + r] = swing ? 0.0 : lbf[r];
      ubC[5 * leg + r] = swing ? 0.0 : ubf[r];
    }
  /* frictionConeConstraint, :274-292 */
  const double mu = p->mu;
  const double Cf[15] = { 1.0, 0.0, -mu, 0.0, 1.0, -mu, 0.0, 1.0, mu, 1.0, 0.0, mu, 0.0, 0.0, 1.0 };
  memset(C, 0, 240 * sizeof(double));
  for (int leg = 0; leg < 4; leg++)
    for (int r = 0; r < 5; r++)
      for (int k = 0; k < 3; k++) C[(5 * leg + r) * 12 + 3 * leg + k] = Cf[3 * r + k];

  /* PD, :126-139 (note :139 adds to index 1, reproduced) */
  double xddot_d[3], wdot_d[3];
  for (int i = 0; i < 3; i++)
    xddot_d[i] = p->kp_p[i] * (s->x_d[i] - s->x[i]) + p->kd_p[i] * (s->xdot_d[i] - s->xdot[i]);
  xddot_d[0] += p->kff[0] * s->xdot_d[0];
  xddot_d[1] += p->kff[1] * s->xdot_d[1];
  xddot_d[2] += p->kff[2] * p->mass * 9.81;

  double Rt[9], Rerr[9], aa[3];
  transpose(3, 3, s->Rwb, Rt);
  matmul(3, 3, 3, s->Rwb_d, Rt, Rerr); /* :133 */
  orc_angle_axis_total(Rerr, aa);
  for (int i = 0; i < 3; i++) wdot_d[i] = p->kp_w[i] * aa[i] + p->kd_w[i] * (s->w_d[i] - s->w[i]);
  wdot_d[0] += p->kff[3] * s->w_d[0];
  wdot_d[1] += p->kff[4] * s->w_d[1];
  wdot_d[1] += p->kff[5] * s->w_d[2];

  /* dynamics, :237-272 */
  double r[12];
  for (int leg = 0; leg < 4; leg++)
    for (int i = 0; i < 3; i++) {
      double acc = 0.0;
      for (int k = 0; k < 3; k++) acc += s->Rwb[3 * i + k] * s->feet[3 * leg + k];
      r[3 * leg + i] = acc;
    }
  double RI[9], Iw[9];
  matmul(3, 3, 3, s->Rwb, p->Ib, RI);
  matmul(3, 3, 3, RI, Rt, Iw); /* :251 */
  memset(A, 0, 72 * sizeof(double));
  for (int leg = 0; leg < 4; leg++) {
    for (int i = 0; i < 3; i++) A[i * 12 + 3 * leg + i] = 1.0;
    const double* v = &r[3 * leg]; /* skew_symmetric, rigid3d.cpp:61-74 */
    double* blk = &A[3 * 12 + 3 * leg];
    blk[0 * 12 + 1] = -v[2];
    blk[0 * 12 + 2] = v[1];
    blk[1 * 12 + 2] = -v[0];
    blk[1 * 12 + 0] = v[2];
    blk[2 * 12 + 0] = -v[1];
    blk[2 * 12 + 1] = v[0];
  }
  const double g[3] = { 0.0, 0.0, -9.81 }; /* :78 */
  for (int i = 0; i < 3; i++) b[i] = p->mass * (xddot_d[i] + g[i]);
  double Iwd[3], Iww[3];
  matmul(3, 3, 1, Iw, wdot_d, Iwd);
  matmul(3, 3, 1, Iw, s->w_d, Iww);
  b[3] = Iwd[0] + (s->w_d[1] * Iww[2] - s->w_d[2] * Iww[1]);
  b[4] = Iwd[1] + (s->w_d[2] * Iww[0] - s->w_d[0] * Iww[2]);
  b[5] = Iwd[2] + (s->w_d[0] * Iww[1] - s->w_d[1] * Iww[0]);

  /* Q = 2(A'SA + W), c = -2 A'S b, :152-153 */
  double At[72], AtS[72], AtSA[144], Sb[6];
  transpose(6, 12, A, At);
  matmul(12, 6, 6, At, p->S, AtS);
  matmul(12, 6, 12, AtS, A, AtSA);
  for (int i = 0; i < 144; i++) Q[i] = 2.0 * (AtSA[i] + p->W[i]);
  matmul(6, 6, 1, p->S, b, Sb);
  for (int i = 0; i < 12; i++) {
    double acc = 0.0;
    for (int k = 0; k < 6; k++) acc += At[i * 6 + k] * Sb[k];
    c[i] = -2.0 * acc;
  }
}

/* ------------------------------------------------------------------------------------------
 * Goldfarb-Idnani dual active-set, QR form.  Constraints n_j'x >= b_j (first me: equalities).
 * Invariant: J'N = [R;0] for the active normals N, J J' = Q^-1.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int n;
  double *J, *R; /* n x n row-major; R upper triangular, q x q used */
  int q;
} gi_fact;

static void gi_add(gi_fact* f, double* d) {
  const int n = f->n, q = f->q;
  for (int j = n - 1; j > q; j--) {
    /* rotate (d[j-1], d[j]) -> (h, 0) */
    const double a = d[j - 1], bb = d[j];
    if (bb == 0.0) continue;
    const double h = hypot(a, bb);
    const double cs = a / h, sn = bb / h;
    d[j - 1] = h;
    d[j] = 0.0;
    for (int i = 0; i < n; i++) {
      const double t1 = f->J[i * n + j - 1], t2 = f->J[i * n + j];
      f->J[i * n + j - 1] = cs * t1 + sn * t2;
      f->J[i * n + j] = -sn * t1 + cs * t2;
    }
  }
  for (int i = 0; i <= q; i++) f->R[i * n + q] = d[i];
  f->q = q + 1;
}

static void gi_drop(gi_fact* f, int k) {
  const int n = f->n, q = f->q;
  for (int j = k; j < q - 1; j++)
    for (int i = 0; i <= j + 1; i++) f->R[i * n + j] = f->R[i * n + j + 1];
  for (int j = k; j < q - 1; j++) {
    /* zero R[j+1][j] with a rotation of rows j, j+1 */
    const double a = f->R[j * n + j], bb = f->R[(j + 1) * n + j];
    if (bb == 0.0) continue;
    const double h = hypot(a, bb);
    const double cs = a / h, sn = bb / h;
    for (int l = j; l < q - 1; l++) {
      const double t1 = f->R[j * n + l], t2 = f->R[(j + 1) * n + l];
      f->R[j * n + l] = cs * t1 + sn * t2;
      f->R[(j + 1) * n + l] = -sn * t1 + cs * t2;
    }
    for (int i = 0; i < n; i++) {
      const double t1 = f->J[i * n + j], t2 = f->J[i * n + j + 1];
      f->J[i * n + j] = cs * t1 + sn * t2;
      f->J[i * n + j + 1] = -sn * t1 + cs * t2;
    }
  }
  f->q = q - 1;
}

int orc_qp_solve(int n, int m, const double* Q, const double* c, const double* C, const double* lb,
                 const double* ub, int max_iter, double* x, double* lam, int* iters) {
  const double INF = 1e20;
  int status = 0, it = 0;
  /* one-sided list */
  int mc = 0;
  /* stack workspace for the balance QP sizes; heap only for larger problems */
  enum { NS = 16, MS = 32 };
  int crow_s[2 * MS], ceq_s[2 * MS], cact_s[2 * MS], A_s[NS + 1];
  double csgn_s[2 * MS], cb_s[2 * MS], ws_s[3 * NS * NS + 6 * NS + 8];
  const int small = (n <= NS && m <= MS);
  int* crow = small ? crow_s : (int*)malloc(sizeof(int) * 2 * m);
  double* csgn = small ? csgn_s : (double*)malloc(sizeof(double) * 2 * m);
  double* cb = small ? cb_s : (double*)malloc(sizeof(double) * 2 * m);
  int* ceq = small ? ceq_s : (int*)malloc(sizeof(int) * 2 * m);
  int* cact = small ? cact_s : (int*)malloc(sizeof(int) * 2 * m);
  memset(cact, 0, sizeof(int) * 2 * m);
  for (int pass = 0; pass < 2; pass++) /* equalities first */
    for (int i = 0; i < m; i++) {
      const int eq = (lb[i] == ub[i]);
      if (pass == 0 && eq) {
        crow[mc] = i; csgn[mc] = 1.0; cb[mc] = lb[i]; ceq[mc] = 1; mc++;
      } else if (pass == 1 && !eq) {
        if (lb[i] > -INF) { crow[mc] = i; csgn[mc] = 1.0; cb[mc] = lb[i]; ceq[mc] = 0; mc++; }
        if (ub[i] < INF) { crow[mc] = i; csgn[mc] = -1.0; cb[mc] = -ub[i]; ceq[mc] = 0; mc++; }
      }
    }
  const size_t wsn = (size_t)(3 * n * n + 6 * n + 8);
  double* ws = small ? ws_s : (double*)malloc(wsn * sizeof(double));
  memset(ws, 0, wsn * sizeof(double));
  double *L = ws, *J = L + n * n, *R = J + n * n;
  double *d = R + n * n, *z = d + n, *r = z + n, *u = r + n, *nv = u + n + 1;
  int* A = small ? A_s : (int*)malloc(sizeof(int) * (n + 1));
  gi_fact f = { n, J, R, 0 };
  if (lam) memset(lam, 0, sizeof(double) * m);

  /* Cholesky Q = L L' */
  for (int j = 0; j < n; j++) {
    double sum = Q[j * n + j];
    for (int k = 0; k < j; k++) sum -= L[j * n + k] * L[j * n + k];
    if (!(sum > 0.0)) { status = 2; goto done; }
    L[j * n + j] = sqrt(sum);
    for (int i = j + 1; i < n; i++) {
      double t = Q[i * n + j];
      for (int k = 0; k < j; k++) t -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = t / L[j * n + j];
    }
  }
  /* J = L^-T: column j solves L' J(:,j) = e_j */
  for (int j = 0; j < n; j++)
    for (int i = n - 1; i >= 0; i--) {
      double t = (i == j) ? 1.0 : 0.0;
      for (int k = i + 1; k < n; k++) t -= L[k * n + i] * J[k * n + j];
      J[i * n + j] = t / L[i * n + i];
    }
  /* x = -J J' c */
  for (int j = 0; j < n; j++) {
    double t = 0.0;
    for (int i = 0; i < n; i++) t += J[i * n + j] * c[i];
    d[j] = t;
  }
  for (int i = 0; i < n; i++) {
    double t = 0.0;
    for (int j = 0; j < n; j++) t += J[i * n + j] * d[j];
    x[i] = -t;
  }

  int p = -1;      /* constraint being added */
  int next_eq = 0; /* equalities are added first, in order */
  for (;;) {
    if (p < 0) {
      double xinf = 0.0;
      for (int i = 0; i < n; i++) xinf = fmax(xinf, fabs(x[i]));
      if (next_eq < mc && ceq[next_eq]) {
        p = next_eq++;
      } else {
        double worst = 0.0;
        for (int j = 0; j < mc; j++) {
          if (cact[j] || ceq[j]) continue;
          const double* row = &C[crow[j] * n];
          double sj = -cb[j], n1 = 0.0;
          for (int i = 0; i < n; i++) { sj += csgn[j] * row[i] * x[i]; n1 += fabs(row[i]); }
          const double tol = 1e-10 * (1.0 + fabs(cb[j]) + n1 * xinf);
          if (sj < -tol && sj < worst) { worst = sj; p = j; }
        }
        if (p < 0) break; /* optimal */
      }
      u[f.q] = 0.0;
    }
    if (it >= max_iter) { status = 1; break; }
    it++;
    for (int i = 0; i < n; i++) nv[i] = csgn[p] * C[crow[p] * n + i];
    double sp = -cb[p];
    for (int i = 0; i < n; i++) sp += nv[i] * x[i];
    const int q = f.q;
    /* d = J' n; z = J2 d2; r = R^-1 d1 */
    double dn = 0.0, d2n = 0.0;
    for (int j = 0; j < n; j++) {
      double t = 0.0;
      for (int i = 0; i < n; i++) t += J[i * n + j] * nv[i];
      d[j] = t;
      dn += t * t;
      if (j >= q) d2n += t * t;
    }
    const int dependent = !(d2n > 1e-22 * dn);
    for (int i = 0; i < n; i++) {
      double t = 0.0;
      for (int j = q; j < n; j++) t += J[i * n + j] * d[j];
      z[i] = t;
    }
    for (int i = q - 1; i >= 0; i--) {
      double t = d[i];
      for (int j = i + 1; j < q; j++) t -= R[i * n + j] * r[j];
      r[i] = t / R[i * n + i];
    }
    if (ceq[p] && dependent) {
      /* linearly dependent equality (swing legs: 5 rows of rank 3): must already hold */
      double xinf = 0.0;
      for (int i = 0; i < n; i++) xinf = fmax(xinf, fabs(x[i]));
      if (fabs(sp) > 1e-9 * (1.0 + fabs(cb[p]) + xinf)) { status = 2; break; }
      p = -1;
      continue;
    }
    /* step lengths */
    double t1 = INFINITY, t2 = INFINITY;
    int k = -1;
    for (int j = 0; j < q; j++) {
      if (ceq[A[j]]) continue;
      if (r[j] > 0.0) {
        const double t = u[j] / r[j];
        if (t < t1) { t1 = t; k = j; }
      }
    }
    if (!dependent) {
      double zn = 0.0;
      for (int i = 0; i < n; i++) zn += z[i] * nv[i];
      t2 = -sp / zn;
      if (ceq[p]) { /* equality: sign-free full step */
        t1 = INFINITY;
      } else if (t2 < 0.0) t2 = 0.0;
    }
    const double t = (t1 < t2) ? t1 : t2;
    if (t == INFINITY) { status = 2; break; } /* infeasible */
    if (t2 == INFINITY) {
      /* dual step only, then drop k */
      for (int j = 0; j < q; j++) u[j] -= t * r[j];
      u[q] += t;
    } else {
      for (int i = 0; i < n; i++) x[i] += t * z[i];
      for (int j = 0; j < q; j++) u[j] -= t * r[j];
      u[q] += t;
    }
    if (t2 != INFINITY && t2 <= t1) {
      /* full step: add p */
      gi_add(&f, d);
      A[q] = p;
      cact[p] = 1;
      p = -1;
    } else {
      cact[A[k]] = 0;
      gi_drop(&f, k);
      for (int j = k; j < q - 1; j++) A[j] = A[j + 1];
      for (int j = k; j < q; j++) u[j] = u[j + 1]; /* u[q-1] is now the pending u+ of p */
    }
  }
  if (lam)
    for (int j = 0; j < f.q; j++) lam[crow[A[j]]] += csgn[A[j]] * u[j];
done:
  if (iters) *iters = it;
  if (!small) { free(A); free(ws); free(cact); free(ceq); free(cb); free(csgn); free(crow); }
  return status;
}

/* ------------------------------------------------------------------------------------------
 * control() + jacobianTransposeControl(): balance_controller.cpp:98-235, kinematics.cpp:218-231
 * ---------------------------------------------------------------------------------------- */
int orc_control(const orc_params* p, const orc_state* s, orc_out* out, double* fw_out) {
  memset(out, 0, sizeof(*out));
  const double* in = (const double*)s;
  for (int i = 0; i < 60; i++)
    if (!isfinite(in[i])) { out->status = 2; return 2; }
  double Q[144], c[12], C[240], lbC[20], ubC[20], A[72], b[6], fw[12];
  orc_assemble(p, s, Q, c, C, lbC, ubC, A, b);
  int iters = 0;
  const int st = orc_qp_solve(12, 20, Q, c, C, lbC, ubC, p->max_iter, fw, NULL, &iters);
  out->status = st;
  out->iters = iters;
  if (fw_out) memcpy(fw_out, fw, sizeof(fw));
  if (st != 0) return st; /* reference: empty ForceMap, :182-216 */
  for (int leg = 0; leg < 4; leg++) {
    if (s->contact[leg] == 0) continue; /* swing legs get no entry, :223-228 */
    double fb[3];
    for (int i = 0; i < 3; i++) {
      double acc = 0.0; /* Rbw = Rwb^T, :218, :226 */
      for (int k = 0; k < 3; k++) acc += s->Rwb[3 * k + i] * fw[3 * leg + k];
      fb[i] = -1.0 * acc;
    }
    double J[9];
    orc_leg_jacobian(p, leg, &s->q[3 * leg], J);
    for (int i = 0; i < 3; i++) {
      out->grf_body[3 * leg + i] = fb[i];
      double tau = J[0 * 3 + i] * fb[0] + J[1 * 3 + i] * fb[1] + J[2 * 3 + i] * fb[2]; /* kinematics.cpp:226 */
      if (p->clamp_tau) tau = fmin(fmax(tau, p->tau_min), p->tau_max); /* commander_node.cpp:526 */
      out->tau[3 * leg + i] = tau;
    }
  }
  return 0;
}

typedef struct {
  const orc_params* p;
  const orc_state* s;
  orc_out* out;
  int64_t lo, hi;
} batch_job;

static void* batch_worker(void* arg) {
  batch_job* j = (batch_job*)arg;
  for (int64_t i = j->lo; i < j->hi; i++) orc_control(j->p, &j->s[i], &j->out[i], NULL);
  return NULL;
}

void orc_control_batch(const orc_params* p, const orc_state* s, int64_t n, orc_out* out, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  batch_job jobs[256];
  for (int t = 0; t < nthreads; t++) {
    jobs[t].p = p; jobs[t].s = s; jobs[t].out = out;
    jobs[t].lo = n * t / nthreads;
    jobs[t].hi = n * (t + 1) / nthreads;
  }
  for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
  batch_worker(&jobs[0]);
  for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
}

/* ==========================================================================================
 * Swing-leg half of the control tick.
 * ======================================================================================== */
void orc_default_joint_gains(orc_joint_gains* g) { /* mit_cheetah_config.yaml:50-53 */
  const double kp[3] = { 40.0, 40.0, 50.0 };
  for (int i = 0; i < 3; i++) { g->kff[i] = 0.0; g->kp[i] = kp[i]; g->kd[i] = 1.0; }
}

/* kinematics.cpp:117-160.  links_ holds the unsigned lengths (kinematics.cpp:28-31). */
void orc_leg_inverse_kinematics(const orc_params* p, int leg, const double foothold[3], double q[3]) {
  const double x = foothold[0] - p->hip_offset[3 * leg];
  const double y = foothold[1] - p->hip_offset[3 * leg + 1];
  const double z = foothold[2] - p->hip_offset[3 * leg + 2];
  const double l1 = fabs(p->link[3 * leg]), l2 = fabs(p->link[3 * leg + 1]), l3 = fabs(p->link[3 * leg + 2]);
  double d = (x * x + y * y + z * z - l1 * l1 - l2 * l2 - l3 * l3) / (2.0 * l2 * l3);
  if (d > 1.0) d = 1.0;
  double sc = y * y + z * z - l1 * l1;
  if (sc < 0.0) sc = 0.0;
  const int right = p->link[3 * leg] < 0.0; /* "FR" or "RR", :147 */
  if (right) q[0] = atan2(z, y) + atan2(sqrt(sc), -l1);
  else q[0] = -(atan2(z, -y) + atan2(sqrt(sc), -l1));
  q[2] = atan2(-sqrt(1.0 - d * d), d);
  q[1] = -atan2(x, sqrt(sc)) - atan2(l3 * sin(q[2]), l2 + l3 * cos(q[2]));
}

/* cyclic Jacobi eigen-decomposition of a symmetric 3x3 (B = V diag(w) V') */
static void jacobi3(double B[9], double V[9], double w[3]) {
  for (int i = 0; i < 9; i++) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; sweep++) {
    const double off = fabs(B[1]) + fabs(B[2]) + fabs(B[5]);
    if (off == 0.0) break;
    for (int pi = 0; pi < 2; pi++)
      for (int qi = pi + 1; qi < 3; qi++) {
        const double apq = B[3 * pi + qi];
        if (apq == 0.0) continue;
        const double theta = (B[4 * qi] - B[4 * pi]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < 3; k++) { /* B <- B G */
          const double bkp = B[3 * k + pi], bkq = B[3 * k + qi];
          B[3 * k + pi] = c * bkp - sn * bkq;
          B[3 * k + qi] = sn * bkp + c * bkq;
        }
        for (int k = 0; k < 3; k++) { /* B <- G' B */
          const double bpk = B[3 * pi + k], bqk = B[3 * qi + k];
          B[3 * pi + k] = c * bpk - sn * bqk;
          B[3 * qi + k] = sn * bpk + c * bqk;
        }
        for (int k = 0; k < 3; k++) {
          const double vkp = V[3 * k + pi], vkq = V[3 * k + qi];
          V[3 * k + pi] = c * vkp - sn * vkq;
          V[3 * k + qi] = sn * vkp + c * vkq;
        }
      }
  }
  w[0] = B[0]; w[1] = B[4]; w[2] = B[8];
}

/* Moore-Penrose pseudo-inverse of a 3x3 (arma::pinv: SVD, tolerance max(m,n) * sigma_max * eps) */
static void pinv3(const double J[9], double out[9]) {
  double B[9], V[9], w[3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) B[3 * i + j] = J[i] * J[j] + J[3 + i] * J[3 + j] + J[6 + i] * J[6 + j]; /* J'J */
  jacobi3(B, V, w);
  double smax = 0.0;
  for (int i = 0; i < 3; i++) { if (w[i] < 0.0) w[i] = 0.0; if (sqrt(w[i]) > smax) smax = sqrt(w[i]); }
  const double tol = 3.0 * smax * 2.220446049250313e-16;
  double M[9]; /* V diag(1/sigma^2) V' */
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double acc = 0.0;
      for (int k = 0; k < 3; k++)
        if (sqrt(w[k]) > tol) acc += V[3 * i + k] * V[3 * j + k] / w[k];
      M[3 * i + j] = acc;
    }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) out[3 * i + j] = M[3 * i] * J[3 * j] + M[3 * i + 1] * J[3 * j + 1] + M[3 * i + 2] * J[3 * j + 2]; /* M J' */
}

/* kinematics.cpp:190-204: arma::inv, then arma::pinv, then the transpose.  Armadillo's inv ends in LAPACK
 * getrf/getri, which fails only on an exactly zero pivot; pinv (SVD) then always succeeds. */
int orc_leg_jacobian_inverse(const orc_params* p, int leg, const double q[3], double Jinv[9]) {
  double J[9], a[9], b[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
  orc_leg_jacobian(p, leg, q, J);
  memcpy(a, J, sizeof(a));
  for (int c = 0; c < 3; c++) {
    int piv = c;
    for (int r = c + 1; r < 3; r++)
      if (fabs(a[3 * r + c]) > fabs(a[3 * piv + c])) piv = r;
    if (a[3 * piv + c] == 0.0 || !isfinite(a[3 * piv + c])) { pinv3(J, Jinv); return 1; }
    for (int j = 0; j < 3; j++) {
      double t = a[3 * c + j]; a[3 * c + j] = a[3 * piv + j]; a[3 * piv + j] = t;
      t = b[3 * c + j]; b[3 * c + j] = b[3 * piv + j]; b[3 * piv + j] = t;
    }
    const double dd = a[3 * c + c];
    for (int j = 0; j < 3; j++) { a[3 * c + j] /= dd; b[3 * c + j] /= dd; }
    for (int r = 0; r < 3; r++)
      if (r != c) {
        const double f = a[3 * r + c];
        for (int j = 0; j < 3; j++) { a[3 * r + j] -= f * a[3 * c + j]; b[3 * r + j] -= f * b[3 * c + j]; }
      }
  }
  memcpy(Jinv, b, sizeof(b));
  return 0;
}

/* math/numerics.cpp:23-50 */
static double wrap_2pi(double a) {
  const double PI = 3.14159265358979323846;
  const double qf = floor(a / (2.0 * PI));
  a -= qf * 2.0 * PI;
  if (a < 0.0) a += 2.0 * PI;
  return a;
}
static double wrap_pi(double r) {
  const double PI = 3.14159265358979323846;
  const double qf = floor((r + PI) / (2.0 * PI));
  r = (r + PI) - qf * 2.0 * PI;
  if (r < 0) r += 2.0 * PI;
  return r - PI;
}

void orc_swing_torques(const orc_params* p, const orc_joint_gains* g, const orc_state* s, const orc_swing* sw,
                       double tau[12], int present[4]) {
  for (int leg = 0; leg < 4; leg++) {
    present[leg] = 0;
    if (s->contact[leg] != 0) continue; /* commander_node.cpp:485 */
    double pb[3], vb[3];
    for (int i = 0; i < 3; i++) { /* :491-492: Rwb' * position - x (sic), Rwb' * velocity */
      double ap = 0.0, av = 0.0;
      for (int k = 0; k < 3; k++) {
        ap += s->Rwb[3 * k + i] * sw->foot_ref_pos[3 * leg + k];
        av += s->Rwb[3 * k + i] * sw->foot_ref_vel[3 * leg + k];
      }
      pb[i] = ap - s->x[i];
      vb[i] = av;
    }
    double qr[3], Jinv[9], qdr[3];
    orc_leg_inverse_kinematics(p, leg, pb, qr);      /* :494 */
    orc_leg_jacobian_inverse(p, leg, qr, Jinv);      /* :495-496 */
    for (int i = 0; i < 3; i++) qdr[i] = Jinv[3 * i] * vb[0] + Jinv[3 * i + 1] * vb[1] + Jinv[3 * i + 2] * vb[2];
    for (int i = 0; i < 3; i++) { /* joint_controller.cpp:27-35 */
      const double qe = wrap_pi(wrap_2pi(qr[i]) - wrap_2pi(s->q[3 * leg + i]));
      const double qde = qdr[i] - sw->qdot[3 * leg + i];
      tau[3 * leg + i] = g->kp[i] * qe + g->kd[i] * qde + g->kff[i];
    }
    present[leg] = 1;
  }
}

int orc_tick(const orc_params* p, const orc_joint_gains* g, const orc_state* s, const orc_swing* sw, orc_out* out) {
  orc_params q = *p;
  q.clamp_tau = 0; /* clamp once, after the merge (:515, :526) */
  const int st = orc_control(&q, s, out, NULL);
  const double* in = (const double*)s;
  for (int i = 0; i < 60; i++)
    if (!isfinite(in[i])) return st; /* nothing is published for a broken state */
  /* the swing references are used as they come, like the reference does (NaN in -> NaN torque for that leg) */
  double tau[12];
  int present[4];
  orc_swing_torques(p, g, s, sw, tau, present);
  for (int leg = 0; leg < 4; leg++)
    if (present[leg])
      for (int i = 0; i < 3; i++) out->tau[3 * leg + i] = tau[3 * leg + i]; /* torque_map.insert(swing...), :515 */
  if (p->clamp_tau)
    for (int i = 0; i < 12; i++) out->tau[i] = fmin(fmax(out->tau[i], p->tau_min), p->tau_max); /* :526 */
  return st;
}

typedef struct {
  const orc_params* p;
  const orc_joint_gains* g;
  const orc_state* s;
  const orc_swing* sw;
  orc_out* out;
  int64_t lo, hi;
} tick_job;

static void* tick_worker(void* arg) {
  tick_job* j = (tick_job*)arg;
  for (int64_t i = j->lo; i < j->hi; i++) orc_tick(j->p, j->g, &j->s[i], &j->sw[i], &j->out[i]);
  return NULL;
}

void orc_tick_batch(const orc_params* p, const orc_joint_gains* g, const orc_state* s, const orc_swing* sw, int64_t n,
                    orc_out* out, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  tick_job jobs[256];
  for (int t = 0; t < nthreads; t++) {
    jobs[t].p = p; jobs[t].g = g; jobs[t].s = s; jobs[t].sw = sw; jobs[t].out = out;
    jobs[t].lo = n * t / nthreads;
    jobs[t].hi = n * (t + 1) / nthreads;
  }
  for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, tick_worker, &jobs[t]);
  tick_worker(&jobs[0]);
  for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
}
