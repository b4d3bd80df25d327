// qpb_math.h -- scalar FP64 helpers shared by every balance kernel (and by the host build of the
// thread-per-QP solver that tests/ compile with g++ to check the device algorithm on the CPU).
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define QPB_HD __host__ __device__ __forceinline__
#else
#define QPB_HD inline
#endif

// warp-wide barrier + memory ordering where lanes of a warp hand data to each other through shared memory; a no-op in
// the host build (there the lanes of a QP are stepped one after the other)
#if defined(__CUDA_ARCH__)
#define QPB_SYNCWARP() __syncwarp()
#else
#define QPB_SYNCWARP() ((void)0)
#endif

namespace qpb {

// MUFU seeds carry >= 20 good bits (e <= 2^-20); one third-order step leaves e^3 <= 2^-60.
QPB_HD double rcp_fast(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  return fma(y, fma(e, e, e), y);  // y (1 + e + e^2)
#else
  return 1.0 / x;
#endif
}

QPB_HD double rsqrt_fast(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-(x * y), y, 1.0);    // 1 - x y^2
  return fma(y * fma(0.375, e, 0.5), e, y);  // y (1 + e/2 + 3e^2/8)
#else
  return 1.0 / sqrt(x);
#endif
}
// the same with one second-order step: e^2 <= 2^-40 -- for the active-set loop, which only decides with it
QPB_HD double rsqrt_loop(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-(x * y), y, 1.0);  // 1 - x y^2
  return fma(0.5 * y, e, y);               // y (1 + e/2)
#else
  return 1.0 / sqrt(x);
#endif
}
QPB_HD double sqrt_fast(double x) { return x > 0.0 ? x * rsqrt_fast(x) : 0.0; }

// atan on [0, inf) x [0, inf): first-quadrant atan2(n, w) (n, w >= 0, not both 0).  Octant reduction
// to |x| <= tan(pi/8), then x + x s P(s) with a degree-11 fit (max relative error 2.2e-16).
QPB_HD double atan2_q1(double n, double w) {
  const bool swap = n > w;
  const double num = swap ? w : n, den = swap ? n : w;
  double r = num * rcp_fast(den);  // in [0, 1]
  const bool hi = r > 0.41421356237309503;
  if (hi) r = (r - 1.0) * rcp_fast(r + 1.0);  // atan(r) = pi/4 + atan((r-1)/(r+1))
  const double s = r * r;
  double q = 1.08884100212413085e-02;
  q = fma(q, s, -2.97502137203377037e-02);
  q = fma(q, s, 4.36598234771156321e-02);
  q = fma(q, s, -5.19008551015063060e-02);
  q = fma(q, s, 5.87346776764138684e-02);
  q = fma(q, s, -6.66594734155613600e-02);
  q = fma(q, s, 7.69226916464344074e-02);
  q = fma(q, s, -9.09090775828966802e-02);
  q = fma(q, s, 1.11111110828050572e-01);
  q = fma(q, s, -1.42857142853820868e-01);
  q = fma(q, s, 1.99999999999983052e-01);
  q = fma(q, s, -3.33333333333333370e-01);
  double a = fma(r * s, q, r);
  if (hi) a += 0.78539816339744831;
  return swap ? 1.5707963267948966 - a : a;
}

// SO(3) log map exactly as Eigen::AngleAxisd(Matrix3d) does it (reference rigid3d.cpp:198-203 ->
// drake RotationMatrix::ToAngleAxis -> Eigen quaternion-from-matrix + angle-axis-from-quaternion).
QPB_HD void angle_axis_total(const double (&R)[9], double (&out)[3]) {
  double qw, qv[3];
  double t = R[0] + R[4] + R[8];
  if (t > 0.0) {
    t = sqrt_fast(t + 1.0);
    qw = 0.5 * t;
    t = 0.5 * rcp_fast(t);
    qv[0] = (R[7] - R[5]) * t;
    qv[1] = (R[2] - R[6]) * t;
    qv[2] = (R[3] - R[1]) * t;
  } else if (R[0] >= R[4] && R[0] >= R[8]) {  // i = 0 (Eigen picks the first largest diagonal)
    t = sqrt_fast(R[0] - R[4] - R[8] + 1.0);
    qv[0] = 0.5 * t;
    t = 0.5 * rcp_fast(t);
    qw = (R[7] - R[5]) * t;
    qv[1] = (R[3] + R[1]) * t;
    qv[2] = (R[6] + R[2]) * t;
  } else if (R[4] > R[0] && R[4] >= R[8]) {  // i = 1
    t = sqrt_fast(R[4] - R[8] - R[0] + 1.0);
    qv[1] = 0.5 * t;
    t = 0.5 * rcp_fast(t);
    qw = (R[2] - R[6]) * t;
    qv[2] = (R[7] + R[5]) * t;
    qv[0] = (R[1] + R[3]) * t;
  } else {  // i = 2
    t = sqrt_fast(R[8] - R[0] - R[4] + 1.0);
    qv[2] = 0.5 * t;
    t = 0.5 * rcp_fast(t);
    qw = (R[3] - R[1]) * t;
    qv[0] = (R[2] + R[6]) * t;
    qv[1] = (R[5] + R[7]) * t;
  }
  double n = sqrt_fast(qv[0] * qv[0] + qv[1] * qv[1] + qv[2] * qv[2]);
  if (n > 1e-150) {
    const double angle = 2.0 * atan2_q1(n, fabs(qw));
    if (qw < 0.0) n = -n;
    const double s = angle * rcp_fast(n);
    out[0] = qv[0] * s;
    out[1] = qv[1] * s;
    out[2] = qv[2] * s;
  } else {
    out[0] = out[1] = out[2] = 0.0;
  }
}

}  // namespace qpb
