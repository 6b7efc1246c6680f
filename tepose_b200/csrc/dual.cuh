// Forward-mode dual numbers for the small geometric functions of the path (6 -> 9 and 9 -> 3 maps): the backward kernels
// evaluate the SAME branch-for-branch code as rotations.cuh on Dual<N> and contract the resulting Jacobian with the
// incoming gradient.  A handful of inputs per item makes forward mode as cheap as a hand-derived adjoint, and it cannot
// drift from the forward code.
#pragma once
#include "common.cuh"

namespace tp {

template <int N>
struct Dual {
  float v;
  float d[N];
  __device__ __forceinline__ Dual() {}
  __device__ __forceinline__ explicit Dual(float c) : v(c) {
#pragma unroll
    for (int i = 0; i < N; ++i) d[i] = 0.0f;
  }
  __device__ __forceinline__ static Dual var(float c, int k) {
    Dual r(c);
    r.d[k] = 1.0f;
    return r;
  }
};

#define TP_DUAL_BIN(op, val, da, db)                                                          \
  template <int N>                                                                            \
  __device__ __forceinline__ Dual<N> operator op(const Dual<N>& a, const Dual<N>& b) {        \
    Dual<N> r;                                                                                \
    r.v = val;                                                                                \
    const float ca = da, cb = db;                                                             \
    _Pragma("unroll") for (int i = 0; i < N; ++i) r.d[i] = ca * a.d[i] + cb * b.d[i];         \
    return r;                                                                                 \
  }
TP_DUAL_BIN(+, a.v + b.v, 1.0f, 1.0f)
TP_DUAL_BIN(-, a.v - b.v, 1.0f, -1.0f)
TP_DUAL_BIN(*, a.v * b.v, b.v, a.v)
TP_DUAL_BIN(/, a.v / b.v, 1.0f / b.v, -a.v / (b.v * b.v))
#undef TP_DUAL_BIN

template <int N>
__device__ __forceinline__ Dual<N> operator*(const Dual<N>& a, float c) {
  Dual<N> r;
  r.v = a.v * c;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * c;
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator+(const Dual<N>& a, float c) {
  Dual<N> r = a;
  r.v += c;
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator-(const Dual<N>& a) { return a * -1.0f; }
template <int N>
__device__ __forceinline__ Dual<N> dsqrt(const Dual<N>& a) {
  Dual<N> r;
  r.v = sqrtf(a.v);
  const float c = 0.5f / r.v;                 // inf at 0, like torch's sqrt backward
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = c * a.d[i];
  return r;
}
// torch.clamp / F.normalize's max(||v||, eps): the gradient passes only where the norm is the larger argument
template <int N>
__device__ __forceinline__ Dual<N> dmax_const(const Dual<N>& a, float c) { return a.v > c ? a : Dual<N>(c); }
template <int N>
__device__ __forceinline__ Dual<N> datan2(const Dual<N>& y, const Dual<N>& x) {
  Dual<N> r;
  r.v = atan2f(y.v, x.v);
  const float den = x.v * x.v + y.v * y.v;
  const float cy = x.v / den, cx = -y.v / den;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = cy * y.d[i] + cx * x.d[i];
  return r;
}

// lib/utils/geometry.py:330-343 on duals (same steps as rot6d_to_rotmat in rotations.cuh)
template <int N>
__device__ __forceinline__ void rot6d_to_rotmat_dual(const Dual<N>* x, Dual<N>* R) {
  const Dual<N> a1x = x[0], a1y = x[2], a1z = x[4], a2x = x[1], a2y = x[3], a2z = x[5];
  const Dual<N> i1 = dmax_const(dsqrt(a1x * a1x + a1y * a1y + a1z * a1z), 1e-6f);
  const Dual<N> b1x = a1x / i1, b1y = a1y / i1, b1z = a1z / i1;
  const Dual<N> d = b1x * a2x + b1y * a2y + b1z * a2z;
  const Dual<N> ux = a2x - d * b1x, uy = a2y - d * b1y, uz = a2z - d * b1z;
  const Dual<N> i2 = dmax_const(dsqrt(ux * ux + uy * uy + uz * uz), 1e-6f);
  const Dual<N> b2x = ux / i2, b2y = uy / i2, b2z = uz / i2;
  R[0] = b1x; R[1] = b2x; R[2] = b1y * b2z - b1z * b2y;
  R[3] = b1y; R[4] = b2y; R[5] = b1z * b2x - b1x * b2z;
  R[6] = b1z; R[7] = b2z; R[8] = b1x * b2y - b1y * b2x;
}

// lib/utils/geometry.py:68-233 on duals (same branches as rotmat_to_angle_axis in rotations.cuh).  Where the reference's
// result is replaced (NaN -> 0) or comes from the constant branch of a torch.where, the derivative is zero.
template <int N>
__device__ __forceinline__ void rotmat_to_angle_axis_dual(const Dual<N>* R, Dual<N>* aa) {
  const Dual<N> m00 = R[0], m01 = R[3], m02 = R[6], m10 = R[1], m11 = R[4], m12 = R[7], m20 = R[2], m21 = R[5], m22 = R[8];
  Dual<N> q0, q1, q2, q3, t;
  if (m22.v < 1e-6f) {
    if (m00.v > m11.v) {
      t = m00 - m11 - m22 + 1.0f;
      q0 = m12 - m21; q1 = t; q2 = m01 + m10; q3 = m20 + m02;
    } else {
      t = m11 - m00 - m22 + 1.0f;
      q0 = m20 - m02; q1 = m01 + m10; q2 = t; q3 = m12 + m21;
    }
  } else {
    if (m00.v < -m11.v) {
      t = m22 - m00 - m11 + 1.0f;
      q0 = m01 - m10; q1 = m20 + m02; q2 = m12 + m21; q3 = t;
    } else {
      t = m00 + m11 + m22 + 1.0f;
      q0 = t; q1 = m12 - m21; q2 = m20 - m02; q3 = m01 - m10;
    }
  }
  const Dual<N> st = dsqrt(t);
  const Dual<N> w = (q0 / st) * 0.5f, x = (q1 / st) * 0.5f, y = (q2 / st) * 0.5f, z = (q3 / st) * 0.5f;
  const Dual<N> s2 = x * x + y * y + z * z;
  Dual<N> k(2.0f);
  if (s2.v > 0.0f) {
    const Dual<N> s = dsqrt(s2);
    const Dual<N> two_theta = ((w.v < 0.0f) ? datan2(-s, -w) : datan2(s, w)) * 2.0f;
    k = two_theta / s;
  }
  const Dual<N> ax = x * k, ay = y * k, az = z * k;
  aa[0] = isnan(ax.v) ? Dual<N>(0.0f) : ax;
  aa[1] = isnan(ay.v) ? Dual<N>(0.0f) : ay;
  aa[2] = isnan(az.v) ? Dual<N>(0.0f) : az;
}

}  // namespace tp
