// fpfh_math.cuh -- per-point arithmetic of the normal / FPFH kernels (host+device inline).
//
// Restates the PCL 1.8.1 routines the reference reaches through src/teaser_utils/fpfh.cc:58-72:
//   pcl::computeRoots / computeRoots2 / eigen33 (smallest eigenpair of the 3x3 covariance),
//   solvePlaneParameters + flipNormalTowardsViewpoint (viewpoint 0,0,0),
//   pcl::computePairFeatures (Darboux frame angles) and the FPFH bin index.
// Every expression keeps a fixed left-to-right binary32 evaluation order (library built with
// -fmad=false); tests/test_math.py compiles this header for the host and compares against the
// CPU oracle bit for bit.
#pragma once
#include <float.h>

#include "qb_math.cuh"

QB_HD_FN void qb_swapf(float& a, float& b) {
  const float t = a;
  a = b;
  b = t;
}

QB_HD_FN void qb_compute_roots2(float b, float c, float roots[3]) {
  roots[0] = 0.0f;
  float d = (float)((double)(b * b) - 4.0 * (double)c);
  if (d < 0.0f) d = 0.0f;
  const float sd = sqrtf(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}

// eigenvalues (ascending) of a symmetric 3x3, row-major m[9]
QB_HD_FN void qb_compute_roots(const float m[9], float roots[3]) {
  const float m00 = m[0], m01 = m[1], m02 = m[2], m11 = m[4], m12 = m[5], m22 = m[8];
  const float c0 = m00 * m11 * m22 + 2.0f * m01 * m02 * m12 - m00 * m12 * m12 - m11 * m02 * m02 - m22 * m01 * m01;
  const float c1 = m00 * m11 - m01 * m01 + m00 * m22 - m02 * m02 + m11 * m22 - m12 * m12;
  const float c2 = m00 + m11 + m22;
  if (fabsf(c0) < FLT_EPSILON) {
    qb_compute_roots2(c2, c1, roots);
    return;
  }
  const float s_inv3 = (float)(1.0 / 3.0);
  const float s_sqrt3 = sqrtf(3.0f);
  const float c2_over_3 = c2 * s_inv3;
  float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.0f) a_over_3 = 0.0f;
  const float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
  float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.0f) q = 0.0f;
  const float rho = sqrtf(-a_over_3);
  const float theta = qb_atan2f(sqrtf(-q), half_b) * s_inv3;
  float sin_theta, cos_theta;
  qb_sincosf(theta, &sin_theta, &cos_theta);
  roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
  roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  if (roots[0] >= roots[1]) qb_swapf(roots[0], roots[1]);
  if (roots[1] >= roots[2]) {
    qb_swapf(roots[1], roots[2]);
    if (roots[0] >= roots[1]) qb_swapf(roots[0], roots[1]);
  }
  if (roots[0] <= 0.0f) qb_compute_roots2(c2, c1, roots);
}

QB_HD_FN void qb_cross3(const float a[3], const float b[3], float o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
QB_HD_FN float qb_dot3(const float a[3], const float b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

QB_HD_FN void qb_eigen33_smallest(const float cov[9], float* eigenvalue, float evec[3]) {
  float scale = 0.0f;
  for (int i = 0; i < 9; ++i) {
    const float a = fabsf(cov[i]);
    scale = (scale < a) ? a : scale;
  }
  if (!(scale > FLT_MIN)) scale = 1.0f;
  float s[9];
  for (int i = 0; i < 9; ++i) s[i] = cov[i] / scale;
  float roots[3];
  qb_compute_roots(s, roots);
  *eigenvalue = roots[0] * scale;
  s[0] -= roots[0];
  s[4] -= roots[0];
  s[8] -= roots[0];
  float v1[3], v2[3], v3[3];
  qb_cross3(&s[0], &s[3], v1);
  qb_cross3(&s[0], &s[6], v2);
  qb_cross3(&s[3], &s[6], v3);
  const float l1 = qb_dot3(v1, v1), l2 = qb_dot3(v2, v2), l3 = qb_dot3(v3, v3);
  float vx, vy, vz, l;
  if (l1 >= l2 && l1 >= l3) { vx = v1[0]; vy = v1[1]; vz = v1[2]; l = l1; }
  else if (l2 >= l1 && l2 >= l3) { vx = v2[0]; vy = v2[1]; vz = v2[2]; l = l2; }
  else { vx = v3[0]; vy = v3[1]; vz = v3[2]; l = l3; }
  const float sl = sqrtf(l);
  evec[0] = vx / sl;
  evec[1] = vy / sl;
  evec[2] = vz / sl;
}

// accu = {sum xx, xy, xz, yy, yz, zz, x, y, z} over cnt neighbours (single-pass float, PCL 1.8.1
// computeMeanAndCovarianceMatrix); (px,py,pz) = query point; out = {nx, ny, nz, curvature}
QB_HD_FN void qb_normal_from_accu(float accu[9], int cnt, float px, float py, float pz, float out[4]) {
  if (cnt < 3) {
    out[0] = out[1] = out[2] = out[3] = NAN;
    return;
  }
  const float fc = (float)cnt;
  for (int i = 0; i < 9; ++i) accu[i] /= fc;
  float cov[9];
  cov[0] = accu[0] - accu[6] * accu[6];
  cov[1] = accu[1] - accu[6] * accu[7];
  cov[2] = accu[2] - accu[6] * accu[8];
  cov[4] = accu[3] - accu[7] * accu[7];
  cov[5] = accu[4] - accu[7] * accu[8];
  cov[8] = accu[5] - accu[8] * accu[8];
  cov[3] = cov[1];
  cov[6] = cov[2];
  cov[7] = cov[5];
  float ev, e[3];
  qb_eigen33_smallest(cov, &ev, e);
  const float eig_sum = cov[0] + cov[4] + cov[8];
  const float curv = (eig_sum != 0.0f) ? fabsf(ev / eig_sum) : 0.0f;
  const float vx = 0.0f - px, vy = 0.0f - py, vz = 0.0f - pz;
  const float cos_theta = (vx * e[0] + vy * e[1]) + vz * e[2];
  if (cos_theta < 0.0f) {
    e[0] *= -1.0f;
    e[1] *= -1.0f;
    e[2] *= -1.0f;
  }
  out[0] = e[0];
  out[1] = e[1];
  out[2] = e[2];
  out[3] = curv;
}

// Darboux features of (p1,n1) -> (p2,n2).  Returns false for the pairs PCL skips.
QB_HD_FN bool qb_pair_features(float p1x, float p1y, float p1z, float n1x, float n1y, float n1z, float p2x, float p2y,
                               float p2z, float n2x, float n2y, float n2z, float* f1, float* f2, float* f3) {
  float dp[3] = {p2x - p1x, p2y - p1y, p2z - p1z};
  const float f4 = sqrtf(qb_dot3(dp, dp));
  if (f4 == 0.0f) return false;
  float n1c[3] = {n1x, n1y, n1z}, n2c[3] = {n2x, n2y, n2z};
  const float angle1 = qb_dot3(n1c, dp) / f4;
  const float angle2 = qb_dot3(n2c, dp) / f4;
  if (qb_acosf(fabsf(angle1)) > qb_acosf(fabsf(angle2))) {
    qb_swapf(n1c[0], n2c[0]);
    qb_swapf(n1c[1], n2c[1]);
    qb_swapf(n1c[2], n2c[2]);
    dp[0] *= -1.0f;
    dp[1] *= -1.0f;
    dp[2] *= -1.0f;
    *f3 = -angle2;
  } else {
    *f3 = angle1;
  }
  float v[3];
  qb_cross3(dp, n1c, v);
  const float v_norm = sqrtf(qb_dot3(v, v));
  if (v_norm == 0.0f) return false;
  v[0] /= v_norm;
  v[1] /= v_norm;
  v[2] /= v_norm;
  float w[3];
  qb_cross3(n1c, v, w);
  *f2 = qb_dot3(v, n2c);
  *f1 = qb_atan2f(qb_dot3(w, n2c), qb_dot3(n1c, n2c));
  return true;
}

// bin of an 11-bin third; scaled = 11 * normalised feature (double, as PCL evaluates it); NaN -> 0
QB_HD_FN int qb_bin_of(double scaled) {
  if (scaled != scaled) return 0;
  int h = (int)floor(scaled);
  if (h < 0) h = 0;
  if (h >= 11) h = 10;
  return h;
}

QB_HD_FN void qb_feature_bins(float f1, float f2, float f3, int* b1, int* b2, int* b3) {
  const float d_pi = 1.0f / (2.0f * (float)M_PI);
  *b1 = qb_bin_of(11 * (((double)f1 + M_PI) * (double)d_pi));
  *b2 = qb_bin_of(11 * (((double)f2 + 1.0) * 0.5));
  *b3 = qb_bin_of(11 * (((double)f3 + 1.0) * 0.5));
}
