// Data terms of the generator loss (lib/core/loss.py:59-171, TePoseLoss.forward) -- value AND gradient in one pass.
//
//   loss_kp_2d  = e_loss_weight    * mean( conf * (pred_kp2d - gt_kp2d)^2 )                     loss.py:106, :179-192 (keypoint_loss)
//   loss_kp_3d  = e_3d_loss_weight * mean( ((P - pelvis_P) - (G - pelvis_G))^2 ), joints 25..38   loss.py:107-108, :194-217
//   loss_pose   = e_pose_weight    * mean( (R(pred_aa) - R(gt_aa))^2 ),  R = quaternion Rodrigues loss.py:122-126, :219-231 (smpl_losses)
//   loss_shape  = e_shape_weight   * mean( (betas_pred - betas_gt)^2 )                           loss.py:124, :229
//
// The reference builds these from ~40 elementwise torch ops and lets autograd differentiate them; the loss is a handful of
// scalars, so forward and backward are one kernel here: every thread produces its term of the sums and the gradient of the
// outputs it touched (the upstream gradient of a scalar loss is a constant factor the caller applies).  Partial sums are reduced
// in a fixed order (per block, then by one block over the blocks): bit-reproducible.
#include "rotations.cuh"

namespace tp {

constexpr int kLossThreads = 128;

struct LossArgs {
  const float* kp2d; const float* real2d; int n2;          // [n2,49,2], [n2,49,3] (x, y, confidence)
  const float* kp3d; const float* real3d; int n3;          // [n3,49,3] both (joints 25..38 are used)
  const float* theta; const float* real_theta; int ns;     // [ns,85] both (axis-angle 3..74, betas 75..84)
  float w2d, w3d, wpose, wshape, openpose_w, gt_w;
  float* g_kp2d; float* g_kp3d; float* g_theta;            // same shapes as the predictions
  float* partial;                                          // [blocks][4]
  int blocks2, blocks3, blockss;
};

// backward of rodrigues_quat (rotations.cuh; lib/utils/geometry.py:22-65): gR[9] -> ga[3]
__device__ __forceinline__ void rodrigues_quat_backward(const float* __restrict__ a, const float* __restrict__ gR, float* ga) {
  const float ex = a[0] + 1e-8f, ey = a[1] + 1e-8f, ez = a[2] + 1e-8f;
  const float n = sqrtf(ex * ex + ey * ey + ez * ez);
  const float ux = a[0] / n, uy = a[1] / n, uz = a[2] / n;
  const float half = n * 0.5f;
  const float c = cosf(half), s = sinf(half);
  const float q0 = c, q1 = s * ux, q2 = s * uy, q3 = s * uz;
  const float qn = sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
  const float w = q0 / qn, x = q1 / qn, y = q2 / qn, z = q3 / qn;
  const float gw = 2.f * (w * (gR[0] + gR[4] + gR[8]) - z * gR[1] + y * gR[2] + z * gR[3] - x * gR[5] - y * gR[6] + x * gR[7]);
  const float gx = 2.f * (x * (gR[0] - gR[4] - gR[8]) + y * gR[1] + z * gR[2] + y * gR[3] - w * gR[5] + z * gR[6] + w * gR[7]);
  const float gy = 2.f * (y * (-gR[0] + gR[4] - gR[8]) + x * gR[1] + w * gR[2] + x * gR[3] + z * gR[5] - w * gR[6] + z * gR[7]);
  const float gz = 2.f * (z * (-gR[0] - gR[4] + gR[8]) - w * gR[1] + x * gR[2] + w * gR[3] + y * gR[5] + x * gR[6] + y * gR[7]);
  const float dot = gw * w + gx * x + gy * y + gz * z;
  const float gq0 = (gw - dot * w) / qn, gq1 = (gx - dot * x) / qn, gq2 = (gy - dot * y) / qn, gq3 = (gz - dot * z) / qn;
  const float g_c = gq0, g_s = gq1 * ux + gq2 * uy + gq3 * uz;
  const float gux = s * gq1, guy = s * gq2, guz = s * gq3;
  float g_n = 0.5f * (-s * g_c + c * g_s);
  g_n -= (gux * a[0] + guy * a[1] + guz * a[2]) / (n * n);
  ga[0] = gux / n + g_n * ex / n;
  ga[1] = guy / n + g_n * ey / n;
  ga[2] = guz / n + g_n * ez / n;
}

__device__ __forceinline__ void block_sum_to(float v, float* dst) {
  __shared__ float red[kLossThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kLossThreads / 32; ++i) s += red[i];
    *dst = s;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kLossThreads) k_tepose_loss_terms(const LossArgs a) {
  const int blk = blockIdx.x, tid = threadIdx.x;
  float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
  if (blk < a.blocks2) {
    // ---- 2-D keypoints: one thread per (row, joint)
    const int64_t i = (int64_t)blk * kLossThreads + tid;
    if (i < (int64_t)a.n2 * 49) {
      const int j = (int)(i % 49);
      const float conf = a.real2d[i * 3 + 2] * (j < 25 ? a.openpose_w : a.gt_w);
      const float dx = a.kp2d[i * 2] - a.real2d[i * 3], dy = a.kp2d[i * 2 + 1] - a.real2d[i * 3 + 1];
      const float sc = a.w2d / ((float)a.n2 * 98.0f);
      t0 = conf * (dx * dx + dy * dy) * sc;
      a.g_kp2d[i * 2] = 2.f * conf * dx * sc;
      a.g_kp2d[i * 2 + 1] = 2.f * conf * dy * sc;
    }
  } else if (blk < a.blocks2 + a.blocks3) {
    // ---- 3-D keypoints: one thread per row, the 14 common joints 25..38, both sets centred on their own pelvis (joints 27, 28)
    const int64_t m = (int64_t)(blk - a.blocks2) * kLossThreads + tid;
    if (m < a.n3) {
      const float* P = a.kp3d + m * 147 + 75;
      const float* G = a.real3d + m * 147 + 75;
      float* gP = a.g_kp3d + m * 147;
      const float sc = a.w3d / ((float)a.n3 * 42.0f);
      float pp[3], pg[3], esum[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 3; ++c) { pp[c] = 0.5f * (P[2 * 3 + c] + P[3 * 3 + c]); pg[c] = 0.5f * (G[2 * 3 + c] + G[3 * 3 + c]); }
      for (int i = 0; i < 75; ++i) gP[i] = 0.f;
      for (int i = 75 + 42; i < 147; ++i) gP[i] = 0.f;
      float e[42];
#pragma unroll
      for (int j = 0; j < 14; ++j)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float d = (P[j * 3 + c] - pp[c]) - (G[j * 3 + c] - pg[c]);
          t1 += d * d * sc;
          e[j * 3 + c] = 2.f * d * sc;
          esum[c] += e[j * 3 + c];
        }
#pragma unroll
      for (int j = 0; j < 14; ++j)
#pragma unroll
        for (int c = 0; c < 3; ++c) gP[75 + j * 3 + c] = e[j * 3 + c] - ((j == 2 || j == 3) ? 0.5f * esum[c] : 0.f);
    }
  } else {
    // ---- SMPL parameters: one thread per (row, joint): rotation matrices of both axis-angle sets; joint 0 also takes the betas / cam
    const int64_t i = (int64_t)(blk - a.blocks2 - a.blocks3) * kLossThreads + tid;
    if (i < (int64_t)a.ns * 24) {
      const int64_t k = i / 24; const int j = (int)(i - k * 24);
      const float* tp_ = a.theta + k * 85 + 3 + j * 3;
      const float* tg = a.real_theta + k * 85 + 3 + j * 3;
      float Rp[9], Rg[9], gR[9], ga[3];
      rodrigues_quat(tp_, Rp);
      rodrigues_quat(tg, Rg);
      const float sc = a.wpose / ((float)a.ns * 216.0f);
#pragma unroll
      for (int q = 0; q < 9; ++q) { const float d = Rp[q] - Rg[q]; t2 += d * d * sc; gR[q] = 2.f * d * sc; }
      rodrigues_quat_backward(tp_, gR, ga);
      float* g = a.g_theta + k * 85;
      g[3 + j * 3] = ga[0]; g[3 + j * 3 + 1] = ga[1]; g[3 + j * 3 + 2] = ga[2];
      if (j == 0) {
        const float ss = a.wshape / ((float)a.ns * 10.0f);
        g[0] = g[1] = g[2] = 0.f;
#pragma unroll
        for (int l = 0; l < 10; ++l) {
          const float d = a.theta[k * 85 + 75 + l] - a.real_theta[k * 85 + 75 + l];
          t3 += d * d * ss;
          g[75 + l] = 2.f * d * ss;
        }
      }
    }
  }
  float* out = a.partial + (int64_t)blk * 4;
  block_sum_to(t0, out); block_sum_to(t1, out + 1); block_sum_to(t2, out + 2); block_sum_to(t3, out + 3);
}

__global__ void __launch_bounds__(kLossThreads) k_tepose_loss_reduce(const float* __restrict__ partial, int nblocks, float* __restrict__ out) {
  // fixed order: thread t adds blocks t, t + 128, ...; then the block tree
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  for (int b = threadIdx.x; b < nblocks; b += kLossThreads)
#pragma unroll
    for (int q = 0; q < 4; ++q) s[q] += partial[(int64_t)b * 4 + q];
#pragma unroll
  for (int q = 0; q < 4; ++q) block_sum_to(s[q], out + q);
}

}  // namespace tp

using namespace tp;

extern "C" size_t tp_tepose_loss_workspace_bytes(int n2, int n3, int ns) {
  const int64_t b = ceil_div((int64_t)n2 * 49, kLossThreads) + ceil_div(n3, kLossThreads) + ceil_div((int64_t)ns * 24, kLossThreads);
  return (size_t)(b > 0 ? b : 1) * 4 * sizeof(float);
}

extern "C" int tp_tepose_loss(const float* kp2d, const float* real2d, int n2, const float* kp3d, const float* real3d, int n3,
                              const float* theta, const float* real_theta, int ns, const float* weights6,
                              float* losses4, float* g_kp2d, float* g_kp3d, float* g_theta, void* workspace, size_t workspace_bytes,
                              void* stream) {
  TP_CHECK_ARG(n2 >= 0 && n3 >= 0 && ns >= 0 && weights6 && losses4, "tp_tepose_loss: bad sizes / null weights or output");
  TP_CHECK_ARG(n2 == 0 || (kp2d && real2d && g_kp2d), "tp_tepose_loss: null 2-D keypoint arrays");
  TP_CHECK_ARG(n3 == 0 || (kp3d && real3d && g_kp3d), "tp_tepose_loss: null 3-D keypoint arrays");
  TP_CHECK_ARG(ns == 0 || (theta && real_theta && g_theta), "tp_tepose_loss: null theta arrays");
  TP_CHECK_ARG(workspace && workspace_bytes >= tp_tepose_loss_workspace_bytes(n2, n3, ns), "tp_tepose_loss: workspace too small");
  LossArgs a;
  a.kp2d = kp2d; a.real2d = real2d; a.n2 = n2; a.kp3d = kp3d; a.real3d = real3d; a.n3 = n3;
  a.theta = theta; a.real_theta = real_theta; a.ns = ns;
  a.w2d = weights6[0]; a.w3d = weights6[1]; a.wpose = weights6[2]; a.wshape = weights6[3]; a.openpose_w = weights6[4]; a.gt_w = weights6[5];
  a.g_kp2d = g_kp2d; a.g_kp3d = g_kp3d; a.g_theta = g_theta;
  a.partial = reinterpret_cast<float*>(workspace);
  a.blocks2 = (int)ceil_div((int64_t)n2 * 49, kLossThreads);
  a.blocks3 = (int)ceil_div(n3, kLossThreads);
  a.blockss = (int)ceil_div((int64_t)ns * 24, kLossThreads);
  const int nblocks = a.blocks2 + a.blocks3 + a.blockss;
  cudaStream_t st = (cudaStream_t)stream;
  if (nblocks == 0) {
    TP_CUDA(cudaMemsetAsync(losses4, 0, 4 * sizeof(float), st));
    return TP_OK;
  }
  k_tepose_loss_terms<<<nblocks, kLossThreads, 0, st>>>(a);
  TP_LAUNCH_CHECK();
  k_tepose_loss_reduce<<<1, kLossThreads, 0, st>>>(a.partial, nblocks, losses4);
  TP_LAUNCH_CHECK();
  return TP_OK;
}
