// Device-side rotation math shared by the geometry kernels and the fused SMPL kernels.
// Each routine follows one reference function branch-for-branch; paths are relative to
// the reference repository.  No fast-math: these feed theta, which is fed back as input.
#pragma once
#include "common.cuh"

namespace tp {

// lib/utils/geometry.py:330-343  rot6d_to_rotmat.  x = 6 numbers read as a (3,2) block:
// a1 = (x0,x2,x4), a2 = (x1,x3,x5).  F.normalize(v, eps=1e-6) = v / max(||v||, 1e-6).
// R (row-major 3x3) has COLUMNS b1, b2, b3.
__device__ __forceinline__ void rot6d_to_rotmat(const float* __restrict__ x, float* R) {
  // Explicit round-to-nearest mul/add (no FMA contraction): torch evaluates these as separate
  // elementwise ops, and for near-parallel a1/a2 the result is decided by the last bit.
  float a1x = x[0], a1y = x[2], a1z = x[4];
  float a2x = x[1], a2y = x[3], a2z = x[5];
  float n1 = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(a1x, a1x), __fmul_rn(a1y, a1y)), __fmul_rn(a1z, a1z)));
  float i1 = fmaxf(n1, 1e-6f);
  float b1x = a1x / i1, b1y = a1y / i1, b1z = a1z / i1;
  float d = __fadd_rn(__fadd_rn(__fmul_rn(b1x, a2x), __fmul_rn(b1y, a2y)), __fmul_rn(b1z, a2z));
  float ux = __fsub_rn(a2x, __fmul_rn(d, b1x)), uy = __fsub_rn(a2y, __fmul_rn(d, b1y)), uz = __fsub_rn(a2z, __fmul_rn(d, b1z));
  float n2 = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ux, ux), __fmul_rn(uy, uy)), __fmul_rn(uz, uz)));
  float i2 = fmaxf(n2, 1e-6f);
  float b2x = ux / i2, b2y = uy / i2, b2z = uz / i2;
  float b3x = __fsub_rn(__fmul_rn(b1y, b2z), __fmul_rn(b1z, b2y));
  float b3y = __fsub_rn(__fmul_rn(b1z, b2x), __fmul_rn(b1x, b2z));
  float b3z = __fsub_rn(__fmul_rn(b1x, b2y), __fmul_rn(b1y, b2x));
  R[0] = b1x; R[1] = b2x; R[2] = b3x;
  R[3] = b1y; R[4] = b2y; R[5] = b3y;
  R[6] = b1z; R[7] = b2z; R[8] = b3z;
}

// smplx.lbs.batch_rodrigues (third-party, restated): angle = ||r + 1e-8||, k = r/angle,
// R = I + sin(angle) K + (1 - cos(angle)) K K.
__device__ __forceinline__ void rodrigues_smplx(const float* __restrict__ r, float* R) {
  float ex = r[0] + 1e-8f, ey = r[1] + 1e-8f, ez = r[2] + 1e-8f;
  float angle = sqrtf(ex * ex + ey * ey + ez * ez);
  float kx = r[0] / angle, ky = r[1] / angle, kz = r[2] / angle;
  float s = sinf(angle), c = 1.0f - cosf(angle);
  // K = [[0,-kz,ky],[kz,0,-kx],[-ky,kx,0]] ; K*K written out
  float kxx = kx * kx, kyy = ky * ky, kzz = kz * kz;
  float kxy = kx * ky, kxz = kx * kz, kyz = ky * kz;
  R[0] = 1.0f + c * (-kzz - kyy);
  R[1] = -s * kz + c * kxy;
  R[2] = s * ky + c * kxz;
  R[3] = s * kz + c * kxy;
  R[4] = 1.0f + c * (-kzz - kxx);
  R[5] = -s * kx + c * kyz;
  R[6] = -s * ky + c * kxz;
  R[7] = s * kx + c * kyz;
  R[8] = 1.0f + c * (-kyy - kxx);
}

// lib/utils/geometry.py:22-65  batch_rodrigues + quat2mat (quaternion form, used by the loss).
__device__ __forceinline__ void rodrigues_quat(const float* __restrict__ r, float* R) {
  float ex = r[0] + 1e-8f, ey = r[1] + 1e-8f, ez = r[2] + 1e-8f;
  float n = sqrtf(ex * ex + ey * ey + ez * ez);
  float ax = r[0] / n, ay = r[1] / n, az = r[2] / n;
  float half = n * 0.5f;
  float vc = cosf(half), vs = sinf(half);
  float w = vc, x = vs * ax, y = vs * ay, z = vs * az;
  float qn = sqrtf(w * w + x * x + y * y + z * z);
  w /= qn; x /= qn; y /= qn; z /= qn;
  float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
  float wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
  R[0] = w2 + x2 - y2 - z2; R[1] = 2 * xy - 2 * wz;     R[2] = 2 * wy + 2 * xz;
  R[3] = 2 * wz + 2 * xy;     R[4] = w2 - x2 + y2 - z2; R[5] = 2 * yz - 2 * wx;
  R[6] = 2 * xz - 2 * wy;     R[7] = 2 * wx + 2 * yz;     R[8] = w2 - x2 - y2 + z2;
}

// lib/utils/geometry.py:68-233: rotation_matrix_to_quaternion on M = R^T (four-way
// select), q = (q_sel / sqrt(t_sel)) * 0.5, then quaternion_to_angle_axis, NaN -> 0.
// R row-major; M[i][j] = R[j][i].
__device__ __forceinline__ void rotmat_to_angle_axis(const float* __restrict__ R, float* aa) {
  float m00 = R[0], m01 = R[3], m02 = R[6];
  float m10 = R[1], m11 = R[4], m12 = R[7];
  float m20 = R[2], m21 = R[5], m22 = R[8];
  float q0, q1, q2, q3, t;
  if (m22 < 1e-6f) {
    if (m00 > m11) {
      t = 1.0f + m00 - m11 - m22;
      q0 = m12 - m21; q1 = t; q2 = m01 + m10; q3 = m20 + m02;
    } else {
      t = 1.0f - m00 + m11 - m22;
      q0 = m20 - m02; q1 = m01 + m10; q2 = t; q3 = m12 + m21;
    }
  } else {
    if (m00 < -m11) {
      t = 1.0f - m00 - m11 + m22;
      q0 = m01 - m10; q1 = m20 + m02; q2 = m12 + m21; q3 = t;
    } else {
      t = 1.0f + m00 + m11 + m22;
      q0 = t; q1 = m12 - m21; q2 = m20 - m02; q3 = m01 - m10;
    }
  }
  float st = sqrtf(t);
  float w = (q0 / st) * 0.5f, x = (q1 / st) * 0.5f, y = (q2 / st) * 0.5f, z = (q3 / st) * 0.5f;
  float s2 = x * x + y * y + z * z;
  float s = sqrtf(s2);
  float two_theta = 2.0f * ((w < 0.0f) ? atan2f(-s, -w) : atan2f(s, w));
  float k = (s2 > 0.0f) ? (two_theta / s) : 2.0f;
  float ax = x * k, ay = y * k, az = z * k;
  aa[0] = isnan(ax) ? 0.0f : ax;
  aa[1] = isnan(ay) ? 0.0f : ay;
  aa[2] = isnan(az) ? 0.0f : az;
}

}  // namespace tp
