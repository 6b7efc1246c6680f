// Stand-alone rotation / projection kernels (one thread per item; these are tiny and
// exist for API parity with lib/utils/geometry.py -- the hot path uses the same device
// functions fused into the SMPL kernels).
#include "rotations.cuh"

namespace tp {

__global__ void k_rot6d(const float* __restrict__ x, float* __restrict__ R, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float xi[6], Ri[9];
#pragma unroll
  for (int k = 0; k < 6; ++k) xi[k] = x[i * 6 + k];
  rot6d_to_rotmat(xi, Ri);
#pragma unroll
  for (int k = 0; k < 9; ++k) R[i * 9 + k] = Ri[k];
}

__global__ void k_r2aa(const float* __restrict__ R, float* __restrict__ aa, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float Ri[9], a[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) Ri[k] = R[i * 9 + k];
  rotmat_to_angle_axis(Ri, a);
  aa[i * 3 + 0] = a[0]; aa[i * 3 + 1] = a[1]; aa[i * 3 + 2] = a[2];
}

__global__ void k_rodrigues(const float* __restrict__ aa, float* __restrict__ R, int64_t n, int form) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a[3] = {aa[i * 3], aa[i * 3 + 1], aa[i * 3 + 2]}, Ri[9];
  if (form == TP_RODRIGUES_QUAT) rodrigues_quat(a, Ri); else rodrigues_smplx(a, Ri);
#pragma unroll
  for (int k = 0; k < 9; ++k) R[i * 9 + k] = Ri[k];
}

// lib/models/spin.py:307-351: t = (cam1, cam2, 2*5000/(224*cam0 + 1e-9)); p = X + t;
// p /= p.z; kp = 5000*p.xy (+0 centre); kp /= 112.  Same operation order as the reference.
__device__ __forceinline__ void project_point(const float* X, const float* cam, float* kp) {
  float tz = 2.0f * 5000.0f / (224.0f * cam[0] + 1e-9f);
  float px = X[0] + cam[1], py = X[1] + cam[2], pz = X[2] + tz;
  px = px / pz; py = py / pz;
  kp[0] = (5000.0f * px) / 112.0f;
  kp[1] = (5000.0f * py) / 112.0f;
}

__global__ void k_projection(const float* __restrict__ joints, const float* __restrict__ cam,
                             float* __restrict__ kp2d, int n, int nj) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * nj) return;
  int b = i / nj;
  float X[3] = {joints[i * 3], joints[i * 3 + 1], joints[i * 3 + 2]};
  float c[3] = {cam[b * 3], cam[b * 3 + 1], cam[b * 3 + 2]}, kp[2];
  project_point(X, c, kp);
  kp2d[i * 2] = kp[0]; kp2d[i * 2 + 1] = kp[1];
}

}  // namespace tp

using namespace tp;

extern "C" int tp_rot6d_to_rotmat(const float* x, float* R, int64_t n, void* stream) {
  TP_CHECK_ARG(n >= 0 && (n == 0 || (x && R)), "tp_rot6d_to_rotmat: null pointer");
  if (n == 0) return TP_OK;
  k_rot6d<<<(unsigned)ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(x, R, n);
  TP_LAUNCH_CHECK();
  return TP_OK;
}

extern "C" int tp_rotmat_to_angle_axis(const float* R, float* aa, int64_t n, void* stream) {
  TP_CHECK_ARG(n >= 0 && (n == 0 || (aa && R)), "tp_rotmat_to_angle_axis: null pointer");
  if (n == 0) return TP_OK;
  k_r2aa<<<(unsigned)ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(R, aa, n);
  TP_LAUNCH_CHECK();
  return TP_OK;
}

extern "C" int tp_batch_rodrigues(const float* aa, float* R, int64_t n, int form, void* stream) {
  TP_CHECK_ARG(n >= 0 && (n == 0 || (aa && R)), "tp_batch_rodrigues: null pointer");
  TP_CHECK_ARG(form == TP_RODRIGUES_SMPLX || form == TP_RODRIGUES_QUAT, "tp_batch_rodrigues: bad form %d", form);
  if (n == 0) return TP_OK;
  k_rodrigues<<<(unsigned)ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(aa, R, n, form);
  TP_LAUNCH_CHECK();
  return TP_OK;
}

extern "C" int tp_projection(const float* joints, const float* cam, float* kp2d, int n, int nj, void* stream) {
  TP_CHECK_ARG(n >= 0 && nj >= 0, "tp_projection: negative size");
  if (n == 0 || nj == 0) return TP_OK;
  TP_CHECK_ARG(joints && cam && kp2d, "tp_projection: null pointer");
  k_projection<<<(unsigned)ceil_div((int64_t)n * nj, 128), 128, 0, (cudaStream_t)stream>>>(joints, cam, kp2d, n, nj);
  TP_LAUNCH_CHECK();
  return TP_OK;
}
