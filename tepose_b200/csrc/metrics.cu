// Evaluation metrics on the device (lib/utils/eval_utils.py; evaluate.py:424-443, lib/core/tester.py:266-315):
// per-frame MPJPE after pelvis alignment, Procrustes-aligned MPJPE (orthogonal Procrustes, 3x3 SVD per frame),
// acceleration error, and the per-body mean vertex distance (MPVPE).  They consume the outputs of the hot path
// where it leaves them (HBM) instead of shipping 82 KB of vertices per frame to the host.
#include "common.cuh"

namespace tp {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// K = U diag(S) V^T for a 3x3 matrix (row-major), one-sided Jacobi in double.  Columns come out sorted by
// descending singular value; a (numerically) zero singular direction gets a unit vector completing U to a
// proper orthonormal basis, so U and V are always orthogonal matrices.
__device__ void svd3(const double* K, double* U, double* S, double* V) {
  double A[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) { A[i] = K[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      double al = 0, be = 0, ga = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i) { al += A[i * 3 + p] * A[i * 3 + p]; be += A[i * 3 + q] * A[i * 3 + q]; ga += A[i * 3 + p] * A[i * 3 + q]; }
      if (ga == 0.0 || fabs(ga) <= 1e-17 * sqrt(al * be)) continue;
      off = fmax(off, fabs(ga) / sqrt(al * be));
      const double zeta = (be - al) / (2.0 * ga);
      const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
      const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double ap = A[i * 3 + p], aq = A[i * 3 + q];
        A[i * 3 + p] = c * ap - s * aq; A[i * 3 + q] = s * ap + c * aq;
        const double vp = V[i * 3 + p], vq = V[i * 3 + q];
        V[i * 3 + p] = c * vp - s * vq; V[i * 3 + q] = s * vp + c * vq;
      }
    }
    if (off < 1e-15) break;
  }
  double n[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) n[j] = sqrt(A[j] * A[j] + A[3 + j] * A[3 + j] + A[6 + j] * A[6 + j]);
  // sort columns by descending norm (3-element network), applied to A and V together
  auto swp = [&](int a, int b) {
    if (n[a] < n[b]) {
      double tn = n[a]; n[a] = n[b]; n[b] = tn;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double ta = A[i * 3 + a]; A[i * 3 + a] = A[i * 3 + b]; A[i * 3 + b] = ta;
        double tv = V[i * 3 + a]; V[i * 3 + a] = V[i * 3 + b]; V[i * 3 + b] = tv;
      }
    }
  };
  swp(0, 1); swp(0, 2); swp(1, 2);
  const double tiny = 1e-300 + 1e-14 * n[0];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    S[j] = n[j];
    if (n[j] > tiny) {
#pragma unroll
      for (int i = 0; i < 3; ++i) U[i * 3 + j] = A[i * 3 + j] / n[j];
    }
  }
  // rank-deficient input: complete U (at most the trailing columns are affected, S is sorted)
  if (!(n[0] > tiny)) { for (int i = 0; i < 9; ++i) U[i] = (i % 4 == 0) ? 1.0 : 0.0; return; }
  if (!(n[1] > tiny)) {
    // any unit vector orthogonal to u0
    const double a0 = fabs(U[0]), a1 = fabs(U[3]), a2 = fabs(U[6]);
    double e[3] = {0, 0, 0};
    e[(a0 <= a1 && a0 <= a2) ? 0 : (a1 <= a2 ? 1 : 2)] = 1.0;
    const double d = e[0] * U[0] + e[1] * U[3] + e[2] * U[6];
    double v[3] = {e[0] - d * U[0], e[1] - d * U[3], e[2] - d * U[6]};
    const double vn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    U[1] = v[0] / vn; U[4] = v[1] / vn; U[7] = v[2] / vn;
  }
  if (!(n[2] > tiny)) {
    U[2] = U[3] * U[7] - U[6] * U[4];
    U[5] = U[6] * U[1] - U[0] * U[7];
    U[8] = U[0] * U[4] - U[3] * U[1];
  }
}

__device__ __forceinline__ double det3(const double* M) {
  return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

// pelvis of one frame: mean of joints p0, p1 (p1 < 0: joint p0 alone; p0 < 0: none)
__device__ __forceinline__ void pelvis_of(const float* x, int p0, int p1, float* c) {
  if (p0 < 0) { c[0] = c[1] = c[2] = 0.f; return; }
  if (p1 < 0) { c[0] = x[p0 * 3]; c[1] = x[p0 * 3 + 1]; c[2] = x[p0 * 3 + 2]; return; }
#pragma unroll
  for (int k = 0; k < 3; ++k) c[k] = (x[p0 * 3 + k] + x[p1 * 3 + k]) / 2.0f;       // evaluate.py:424-425
}

// One warp per frame.  S1 / S2 [n,J,3]; optional pelvis alignment first (evaluate.py:424-428); outputs (each
// optional): S1_hat [n,J,3] (lib/utils/eval_utils.py:287-337), mpjpe [n] (evaluate.py:433), pa_mpjpe [n] (:435-436).
__global__ void k_pose_metrics(const float* __restrict__ S1, const float* __restrict__ S2, int n, int J, int p0, int p1,
                               float* __restrict__ S1_hat, float* __restrict__ mpjpe, float* __restrict__ pa) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n) return;
  const float* a = S1 + (size_t)w * J * 3;
  const float* b = S2 + (size_t)w * J * 3;
  float ca[3], cb[3];
  pelvis_of(a, p0, p1, ca);
  pelvis_of(b, p0, p1, cb);
  // 1. means of the (pelvis-aligned) point sets, plain MPJPE on the way
  float m1[3] = {0, 0, 0}, m2[3] = {0, 0, 0}, e = 0.f;
  for (int j = lane; j < J; j += 32) {
    float d2 = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float x = a[j * 3 + k] - ca[k], y = b[j * 3 + k] - cb[k];
      m1[k] += x; m2[k] += y;
      d2 += (x - y) * (x - y);
    }
    e += sqrtf(d2);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) { m1[k] = warp_sum(m1[k]) / (float)J; m2[k] = warp_sum(m2[k]) / (float)J; }
  e = warp_sum(e);
  if (mpjpe && lane == 0) mpjpe[w] = e / (float)J;
  if (!S1_hat && !pa) return;
  // 2./3. variance of X1 and K = X1 X2^T (accumulated in double: 3x3 sums of products of centred fp32 values)
  double Kd[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, var1 = 0.0;
  for (int j = lane; j < J; j += 32) {
    float x[3], y[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { x[k] = (a[j * 3 + k] - ca[k]) - m1[k]; y[k] = (b[j * 3 + k] - cb[k]) - m2[k]; }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      var1 += (double)x[r] * x[r];
#pragma unroll
      for (int c = 0; c < 3; ++c) Kd[r * 3 + c] += (double)x[r] * y[c];
    }
  }
  var1 = warp_sum(var1);
#pragma unroll
  for (int i = 0; i < 9; ++i) Kd[i] = warp_sum(Kd[i]);
  // 4. R = V Z U^T with det(R) = +1; 5. scale = tr(R K) / var1; 6. t = mu2 - scale R mu1   (every lane computes it:
  //    the warp is otherwise idle and this avoids 13 shuffles)
  double U[9], S[3], V[9], R[9];
  svd3(Kd, U, S, V);
  double UVt[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) UVt[r * 3 + c] = U[r * 3] * V[c * 3] + U[r * 3 + 1] * V[c * 3 + 1] + U[r * 3 + 2] * V[c * 3 + 2];
  const double dz = det3(UVt);
  const double z = dz > 0 ? 1.0 : (dz < 0 ? -1.0 : 0.0);                // torch.sign
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) R[r * 3 + c] = V[r * 3] * U[c * 3] + V[r * 3 + 1] * U[c * 3 + 1] + z * V[r * 3 + 2] * U[c * 3 + 2];
  double tr = 0.0;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) tr += R[r * 3 + c] * Kd[c * 3 + r];
  const double scale = tr / var1;
  float Rf[9], tf[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    tf[r] = (float)((double)m2[r] - scale * (R[r * 3] * m1[0] + R[r * 3 + 1] * m1[1] + R[r * 3 + 2] * m1[2]));
#pragma unroll
    for (int c = 0; c < 3; ++c) Rf[r * 3 + c] = (float)(scale * R[r * 3 + c]);
  }
  // 7. S1_hat = scale R S1 + t, and the aligned error
  float epa = 0.f;
  for (int j = lane; j < J; j += 32) {
    const float x0 = a[j * 3] - ca[0], x1 = a[j * 3 + 1] - ca[1], x2 = a[j * 3 + 2] - ca[2];
    float d2 = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float h = Rf[r * 3] * x0 + Rf[r * 3 + 1] * x1 + Rf[r * 3 + 2] * x2 + tf[r];
      if (S1_hat) S1_hat[((size_t)w * J + j) * 3 + r] = h;
      const float d = h - (b[j * 3 + r] - cb[r]);
      d2 += d * d;
    }
    epa += sqrtf(d2);
  }
  epa = warp_sum(epa);
  if (pa && lane == 0) pa[w] = epa / (float)J;
}

// One warp per (sequence, interior frame): out[s, i] = mean_j || accel(pred)_i - accel(target)_i ||,
// accel(x)_i = x_i - 2 x_{i+1} + x_{i+2}  (lib/utils/eval_utils.py:79-138); target == NULL gives the norm of the
// predicted acceleration itself (compute_accel, :53-76).  Pelvis alignment as in k_pose_metrics.
__global__ void k_accel_error(const float* __restrict__ P, const float* __restrict__ G, int n_seq, int len, int J, int p0, int p1,
                              float* __restrict__ out) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int per = len - 2;
  if (per <= 0 || w >= n_seq * per) return;
  const int s = w / per, i = w - s * per;
  const float* p = P + ((size_t)s * len + i) * J * 3;
  const float* g = G ? G + ((size_t)s * len + i) * J * 3 : nullptr;
  const size_t fs = (size_t)J * 3;
  float cp[3][3], cg[3][3];
#pragma unroll
  for (int f = 0; f < 3; ++f) {
    pelvis_of(p + f * fs, p0, p1, cp[f]);
    if (g) pelvis_of(g + f * fs, p0, p1, cg[f]);
  }
  float e = 0.f;
  for (int j = lane; j < J; j += 32) {
    float d2 = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float ap = (p[j * 3 + k] - cp[0][k]) - 2.0f * (p[fs + j * 3 + k] - cp[1][k]) + (p[2 * fs + j * 3 + k] - cp[2][k]);
      float d = ap;
      if (g) d -= (g[j * 3 + k] - cg[0][k]) - 2.0f * (g[fs + j * 3 + k] - cg[1][k]) + (g[2 * fs + j * 3 + k] - cg[2][k]);
      d2 += d * d;
    }
    e += sqrtf(d2);
  }
  e = warp_sum(e);
  if (lane == 0) out[w] = e / (float)J;
}

// One CTA per body: out[b] = mean_v || A[b,v] - B[b,v] ||  (lib/utils/eval_utils.py:173-175).  HBM-bound: both
// meshes are read exactly once (2 x 12 V bytes per body), nothing is written but 4 bytes.
__global__ void __launch_bounds__(256) k_vertex_error(const float* __restrict__ A, const float* __restrict__ B, int V,
                                                      float* __restrict__ out) {
  const size_t base = (size_t)blockIdx.x * V * 3;
  const float* a = A + base;
  const float* b = B + base;
  float e = 0.f;
  for (int v = threadIdx.x; v < V; v += 256) {
    const float dx = a[v * 3] - b[v * 3], dy = a[v * 3 + 1] - b[v * 3 + 1], dz = a[v * 3 + 2] - b[v * 3 + 2];
    e += sqrtf(dx * dx + dy * dy + dz * dz);
  }
  __shared__ float part[8];
  e = warp_sum(e);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = e;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i];
    out[blockIdx.x] = t / (float)V;
  }
}

}  // namespace tp

extern "C" int tp_pose_metrics(const float* pred, const float* target, int n, int n_joints, int pelvis0, int pelvis1,
                               float* aligned, float* mpjpe, float* pa_mpjpe, void* stream) {
  using namespace tp;
  TP_CHECK_ARG(n >= 0 && n_joints >= 1, "tp_pose_metrics: bad sizes n=%d joints=%d", n, n_joints);
  TP_CHECK_ARG(pelvis0 < n_joints && pelvis1 < n_joints, "tp_pose_metrics: pelvis joint index out of range");
  if (n == 0) return TP_OK;
  TP_CHECK_ARG(pred && target, "tp_pose_metrics: null pointer");
  TP_CHECK_ARG(aligned || mpjpe || pa_mpjpe, "tp_pose_metrics: no output requested");
  const int wpb = 4;
  k_pose_metrics<<<(unsigned)ceil_div(n, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(pred, target, n, n_joints, pelvis0, pelvis1,
                                                                                    aligned, mpjpe, pa_mpjpe);
  TP_LAUNCH_CHECK();
  return TP_OK;
}

extern "C" int tp_accel_error(const float* pred, const float* target, int n_seq, int len, int n_joints, int pelvis0, int pelvis1,
                              float* out, void* stream) {
  using namespace tp;
  TP_CHECK_ARG(n_seq >= 0 && len >= 0 && n_joints >= 1, "tp_accel_error: bad sizes");
  TP_CHECK_ARG(pelvis0 < n_joints && pelvis1 < n_joints, "tp_accel_error: pelvis joint index out of range");
  if (n_seq == 0 || len < 3) return TP_OK;
  TP_CHECK_ARG(pred && out, "tp_accel_error: null pointer");
  const int64_t warps = (int64_t)n_seq * (len - 2);
  TP_CHECK_ARG(warps < (1ll << 31) / 32, "tp_accel_error: too many frames");
  const int wpb = 8;
  k_accel_error<<<(unsigned)ceil_div(warps, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(pred, target, n_seq, len, n_joints, pelvis0,
                                                                                      pelvis1, out);
  TP_LAUNCH_CHECK();
  return TP_OK;
}

extern "C" int tp_vertex_error(const float* verts_a, const float* verts_b, int n, int n_verts, float* out, void* stream) {
  using namespace tp;
  TP_CHECK_ARG(n >= 0 && n_verts >= 1, "tp_vertex_error: bad sizes");
  if (n == 0) return TP_OK;
  TP_CHECK_ARG(verts_a && verts_b && out, "tp_vertex_error: null pointer");
  k_vertex_error<<<(unsigned)n, 256, 0, (cudaStream_t)stream>>>(verts_a, verts_b, n_verts, out);
  TP_LAUNCH_CHECK();
  return TP_OK;
}
