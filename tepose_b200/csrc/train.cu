// Backward kernels of the training step (BASELINE.json configs[4]; the caller is lib/core/trainer.py:203,235-237:
// `generator(inp, is_train=True)` ... `loss.backward()`).  The reference gets these from torch.autograd over its torch
// ops; here every non-GEMM stage has an explicit adjoint kernel and the GEMM-shaped stages reuse the library's GEMMs on
// transposed operands.  The algebra of the SMPL part is stated (and checked against autograd on the CPU) in
// oracle/smpl_backward_proto.py; the 6 -> 9 and 9 -> 3 rotation maps are differentiated in forward mode (dual.cuh).
#include "common.cuh"
#include "dual.cuh"
#include "rotations.cuh"

namespace tp {

// ------------------------------------------------------------------------------------------ small utilities
// dst[c][r] = src[r][c]  (fp32 in, fp32 or bf16 out); 32 x 32 tiles through shared memory
template <typename OutT>
__global__ void k_transpose(const float* __restrict__ src, int64_t ld_src, int rows, int cols, OutT* __restrict__ dst, int64_t ld_dst,
                            int dst_rows_padded, int relu) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    float x = (r < rows && c < cols) ? src[(int64_t)r * ld_src + c] : 0.0f;
    if (relu) x = fmaxf(x, 0.0f);
    tile[i][threadIdx.x] = x;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;       // dst row = source column
    if (c < dst_rows_padded && r < ld_dst) {
      const float v = tile[threadIdx.x][i];
      if constexpr (sizeof(OutT) == 2) dst[(int64_t)c * ld_dst + r] = __float2bfloat16_rn(v);
      else dst[(int64_t)c * ld_dst + r] = v;
    }
  }
}

// out[c] = beta * out[c] + sum_r A[r][c]   (bias gradients).  Block = 32 columns x 8 row groups; every thread sums its rows in
// order, the 8 partial sums are added in a fixed order (deterministic).
__global__ void k_colsum(const float* __restrict__ A, int64_t lda, int rows, int cols, float* __restrict__ out, float beta) {
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.0f;
  if (c < cols)
    for (int r = threadIdx.y; r < rows; r += 8) s += A[(int64_t)r * lda + c];
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += part[q][threadIdx.x];
    out[c] = beta != 0.0f ? beta * out[c] + t : t;
  }
}

// a[r][c] *= mask[r][c] * scale   (dropout forward and backward: the same op)
__global__ void k_mask_scale(float* __restrict__ a, int64_t ld, const float* __restrict__ mask, int64_t ldm, int rows, int cols, float scale) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
  a[(int64_t)r * ld + c] *= mask[(int64_t)r * ldm + c] * scale;
}

// g[r][c] = h[r][c] > 0 ? g[r][c] : 0   (relu backward)
__global__ void k_relu_backward(float* __restrict__ g, int64_t ld, const float* __restrict__ h, int64_t ldh, int rows, int cols) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
  if (!(h[(int64_t)r * ldh + c] > 0.0f)) g[(int64_t)r * ld + c] = 0.0f;
}

// dst[r][c] = alpha * src[r][c] + beta * dst[r][c]
__global__ void k_axpby(const float* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd, int rows, int cols, float alpha, float beta) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
  float* d = dst + (int64_t)r * ldd + c;
  *d = alpha * src[(int64_t)r * lds + c] + (beta != 0.0f ? beta * *d : 0.0f);
}

// ------------------------------------------------------------------------------------------ GRU cell adjoint
// torch.nn.GRU cell:  r = s(gi_r + gh_r), z = s(gi_z + gh_z), n = tanh(gi_n + r * hn), hn = W_hn h + b_hn,
// h' = (1 - z) n + z h.  Given g = dL/dh', the saved (r, z, n, hn) [B,4H] and h (NULL = zeros):
//   d_gi = [g_r r (1-r), g_z z (1-z), g_a]   d_gh = [same, same, g_a r]   with g_a = g (1-z)(1-n^2), g_z = g (h - n),
//   g_r = g_a hn;   g <- g z  (the direct path to h; the W_hh^T d_gh term is added by the caller's GEMM).
__global__ void k_gru_cell_backward(float* __restrict__ g, int64_t ldg, const float* __restrict__ gates, int64_t ld_gates,
                                    const float* __restrict__ hprev, int64_t ldh, float* __restrict__ dgi, int64_t ld_dgi,
                                    float* __restrict__ dgh, int64_t ld_dgh, int B, int H) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * H) return;
  const int b = (int)(i / H), u = (int)(i - (int64_t)b * H);
  const float* gt = gates + (int64_t)b * ld_gates;
  const float r = gt[u], z = gt[H + u], n = gt[2 * H + u], hn = gt[3 * H + u];
  const float h = hprev ? hprev[(int64_t)b * ldh + u] : 0.0f;
  const float gh = g[(int64_t)b * ldg + u];
  const float ga = gh * (1.0f - z) * (1.0f - n * n);
  const float gz = gh * (h - n) * z * (1.0f - z);
  const float gr = ga * hn * r * (1.0f - r);
  float* di = dgi + (int64_t)b * ld_dgi;
  float* dh = dgh + (int64_t)b * ld_dgh;
  di[u] = gr; di[H + u] = gz; di[2 * H + u] = ga;
  dh[u] = gr; dh[H + u] = gz; dh[2 * H + u] = ga * r;
  g[(int64_t)b * ldg + u] = gh * z;
}

// ------------------------------------------------------------------------------------------ rotation adjoints
__global__ void k_rot6d_backward(const float* __restrict__ x6, const float* __restrict__ gR, float* __restrict__ gx, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Dual<6> x[6], R[9];
#pragma unroll
  for (int k = 0; k < 6; ++k) x[k] = Dual<6>::var(x6[i * 6 + k], k);
  rot6d_to_rotmat_dual(x, R);
  float acc[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int o = 0; o < 9; ++o) {
    const float go = gR[i * 9 + o];
#pragma unroll
    for (int k = 0; k < 6; ++k) acc[k] += go * R[o].d[k];
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) gx[i * 6 + k] = acc[k];
}

// gR (+)= J^T g_aa for the rotation-matrix -> axis-angle map of lib/utils/geometry.py:68-233
__global__ void k_r2aa_backward(const float* __restrict__ Rm, const float* __restrict__ gaa, int64_t ld_gaa, float* __restrict__ gR,
                                int64_t n, int per_row, int accumulate) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Dual<9> R[9], aa[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) R[k] = Dual<9>::var(Rm[i * 9 + k], k);
  rotmat_to_angle_axis_dual(R, aa);
  const float* ga = gaa + (i / per_row) * ld_gaa + (i % per_row) * 3;      // per_row rotations per gradient row (24 per body)
  float acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    float a = ga[0] * aa[0].d[k] + ga[1] * aa[1].d[k] + ga[2] * aa[2].d[k];
    if (isnan(a) || isinf(a)) a = 0.0f;     // degenerate branch points (t_sel -> 0): the reference would emit NaN gradients
    acc[k] = a;
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) gR[i * 9 + k] = accumulate ? gR[i * 9 + k] + acc[k] : acc[k];
}

// ------------------------------------------------------------------------------------------ SMPL adjoint
constexpr int BW_VT = 128;          // vertices per block
constexpr int BW_NB = 8;            // bodies per block (blend rows are loaded once per 8 bodies)
constexpr int BW_CHAIN = 24 * (3 + 9 + 3 + 3 + 12);   // floats of forward chain state kept per body: J, RW, t, rel, A
constexpr int BW_MAXJ = 64, BW_MAXREG = 32;

struct BwWs {
  float* chain;     // [n][BW_CHAIN]
  float* gposed;    // [n][24][3]
  float* gextra;    // [n][BW_MAXREG][3]
  float* gj;        // [n][BW_MAXJ][3]   gradient of every output joint (after the projection adjoint)
  int* pick_vid;    // [BW_MAXJ]         vertex id of output joint k if it is a vertex pick, else -1
  float* gvp;       // [n][3][vp]        gradient of v_posed, plane layout of the blend table
  float* partial;   // [n][tiles][288]   per-tile sums of the skinning-transform gradients
  float* gbl;       // [n][224]          [g_pose_feature (207) | g_betas (10)] from the blend GEMM
  void* gemm_ws; size_t gemm_ws_bytes;
};

// block per body: forward chain (thread 0), projection adjoint + joint scatter (all threads)
__global__ void k_bw_prepare(const tp_smpl_model m, int n, const float* __restrict__ R, const float* __restrict__ betas, int64_t ld_betas,
                             const float* __restrict__ cam, int64_t ld_cam, int nreg, const int* __restrict__ joint_src, int nj,
                             const float* __restrict__ joints, const float* __restrict__ g_joints, const float* __restrict__ g_kp2d,
                             float* __restrict__ g_cam, BwWs ws) {
  const int b = blockIdx.x, tid = threadIdx.x;
  __shared__ float s_gp[24 * 3], s_ge[BW_MAXREG * 3], s_gt[BW_MAXJ * 3];
  for (int i = tid; i < 24 * 3; i += blockDim.x) s_gp[i] = 0.0f;
  for (int i = tid; i < BW_MAXREG * 3; i += blockDim.x) s_ge[i] = 0.0f;
  __syncthreads();
  if (tid == 0) {
    float* ch = ws.chain + (size_t)b * BW_CHAIN;
    float* J = ch; float* RW = J + 72; float* t = RW + 216; float* rel = t + 72; float* A = rel + 72;
    const float* be = betas + (int64_t)b * ld_betas;
    for (int i = 0; i < 72; ++i) {
      float a = m.j_template[i];
      for (int l = 0; l < 10; ++l) a += m.j_shapedirs[i * 10 + l] * be[l];
      J[i] = a;
    }
    const float* Rb = R + (size_t)b * 216;
    for (int i = 0; i < 24; ++i) {
      const int p = m.parents[i];
      for (int c = 0; c < 3; ++c) rel[i * 3 + c] = J[i * 3 + c] - (p >= 0 ? J[p * 3 + c] : 0.0f);
      if (p < 0) {
        for (int e = 0; e < 9; ++e) RW[i * 9 + e] = Rb[i * 9 + e];
        for (int c = 0; c < 3; ++c) t[i * 3 + c] = rel[i * 3 + c];
      } else {
        for (int r = 0; r < 3; ++r) {
          for (int c = 0; c < 3; ++c)
            RW[i * 9 + r * 3 + c] = RW[p * 9 + r * 3] * Rb[i * 9 + c] + RW[p * 9 + r * 3 + 1] * Rb[i * 9 + 3 + c] + RW[p * 9 + r * 3 + 2] * Rb[i * 9 + 6 + c];
          t[i * 3 + r] = RW[p * 9 + r * 3] * rel[i * 3] + RW[p * 9 + r * 3 + 1] * rel[i * 3 + 1] + RW[p * 9 + r * 3 + 2] * rel[i * 3 + 2] + t[p * 3 + r];
        }
      }
    }
    for (int i = 0; i < 24; ++i)
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) A[i * 12 + r * 4 + c] = RW[i * 9 + r * 3 + c];
        A[i * 12 + r * 4 + 3] = t[i * 3 + r] - (RW[i * 9 + r * 3] * J[i * 3] + RW[i * 9 + r * 3 + 1] * J[i * 3 + 1] + RW[i * 9 + r * 3 + 2] * J[i * 3 + 2]);
      }
  }
  // projection adjoint (lib/models/spin.py:307-351) and scatter of the joint gradients onto their sources
  for (int k = tid; k < nj; k += blockDim.x) {
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (g_joints) { const float* g = g_joints + ((size_t)b * nj + k) * 3; gx = g[0]; gy = g[1]; gz = g[2]; }
    float tx = 0.f, ty = 0.f, tz = 0.f;
    if (g_kp2d && cam) {
      const float* c = cam + (int64_t)b * ld_cam;
      const float s = 5000.0f / 112.0f;
      const float den = 224.0f * c[0] + 1e-9f;
      const float* Xj = joints + ((size_t)b * nj + k) * 3;
      const float px = Xj[0] + c[1], py = Xj[1] + c[2], pz = Xj[2] + 2.0f * 5000.0f / den;
      const float gu = g_kp2d[((size_t)b * nj + k) * 2], gv = g_kp2d[((size_t)b * nj + k) * 2 + 1];
      tx = s * gu / pz; ty = s * gv / pz; tz = -s * (gu * px + gv * py) / (pz * pz);
      gx += tx; gy += ty; gz += tz;
    }
    s_gt[k * 3] = tx; s_gt[k * 3 + 1] = ty; s_gt[k * 3 + 2] = tz;
    float* gj = ws.gj + ((size_t)b * BW_MAXJ + k) * 3;
    gj[0] = gx; gj[1] = gy; gj[2] = gz;
    const int code = joint_src[k];
    if (code < TP_JSRC_REGRESSED(0)) {                 // posed joint: at most two output joints share a source (a + b commutes)
      atomicAdd(&s_gp[code * 3], gx); atomicAdd(&s_gp[code * 3 + 1], gy); atomicAdd(&s_gp[code * 3 + 2], gz);
    } else if (code < TP_JSRC_VERTEX(0)) {
      const int r = code - TP_JSRC_REGRESSED(0);
      atomicAdd(&s_ge[r * 3], gx); atomicAdd(&s_ge[r * 3 + 1], gy); atomicAdd(&s_ge[r * 3 + 2], gz);
    }
    if (b == 0) ws.pick_vid[k] = code >= TP_JSRC_VERTEX(0) ? code - TP_JSRC_VERTEX(0) : -1;
  }
  __syncthreads();
  for (int i = tid; i < 72; i += blockDim.x) ws.gposed[(size_t)b * 72 + i] = s_gp[i];
  for (int i = tid; i < BW_MAXREG * 3; i += blockDim.x) ws.gextra[(size_t)b * BW_MAXREG * 3 + i] = s_ge[i];
  if (tid == 0 && g_cam) {
    float sx = 0.f, sy = 0.f, sz = 0.f;
    if (g_kp2d && cam) {
      for (int k = 0; k < nj; ++k) { sx += s_gt[k * 3]; sy += s_gt[k * 3 + 1]; sz += s_gt[k * 3 + 2]; }
      const float den = 224.0f * cam[(int64_t)b * ld_cam] + 1e-9f;
      sz *= -2.0f * 5000.0f * 224.0f / (den * den);
    }
    g_cam[b * 3] = sz; g_cam[b * 3 + 1] = sx; g_cam[b * 3 + 2] = sy;
  }
}

// grid (vertex tiles, body groups): thread = vertex x BW_NB bodies.  Recomputes v_posed (blend rows are read once per 8
// bodies), assembles the vertex gradient (caller's g_verts + vertex picks + regressed joints), applies the skinning adjoint:
// g_vposed = T_R^T g_v (stored in the blend table's plane layout for the GEMM that follows) and the per-joint sums
// g_A_j = sum_v w_vj [g_v v_posed^T | g_v], reduced warp -> block in a fixed order (deterministic).
__global__ void __launch_bounds__(BW_VT) k_bw_lbs(const tp_smpl_model m, int n, const float* __restrict__ R, const float* __restrict__ betas,
                                                  int64_t ld_betas, const float* __restrict__ jreg, int nreg, int nj,
                                                  const float* __restrict__ g_verts, BwWs ws) {
  __shared__ float s_coef[BW_NB][220];
  __shared__ float s_A[BW_NB][24 * 12];
  __shared__ float s_ge[BW_NB][BW_MAXREG * 3];
  __shared__ float s_gj[BW_NB][BW_MAXJ * 3];
  __shared__ int s_pick[BW_MAXJ];
  __shared__ float s_slab[BW_VT / 32][288];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int v = blockIdx.x * BW_VT + tid, b0 = blockIdx.y * BW_NB;
  const int nb = min(BW_NB, n - b0);
  const int vp = m.vp;
  for (int i = tid; i < BW_NB * 220; i += BW_VT) {
    const int bb = i / 220, k = i - bb * 220;
    float c = 0.0f;
    if (bb < nb) {
      if (k < 207) { const int j = 1 + k / 9, e = k % 9; c = R[(size_t)(b0 + bb) * 216 + j * 9 + e] - ((e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f); }
      else if (k < 217) c = betas[(int64_t)(b0 + bb) * ld_betas + (k - 207)];
      else if (k == 217) c = 1.0f;
    }
    s_coef[bb][k] = c;
  }
  for (int i = tid; i < BW_NB * 288; i += BW_VT) { const int bb = i / 288; s_A[bb][i - bb * 288] = bb < nb ? ws.chain[(size_t)(b0 + bb) * BW_CHAIN + 432 + (i - bb * 288)] : 0.0f; }
  for (int i = tid; i < BW_NB * BW_MAXREG * 3; i += BW_VT) { const int bb = i / (BW_MAXREG * 3); s_ge[bb][i - bb * BW_MAXREG * 3] = bb < nb ? ws.gextra[(size_t)(b0 + bb) * BW_MAXREG * 3 + (i - bb * BW_MAXREG * 3)] : 0.0f; }
  for (int i = tid; i < BW_NB * BW_MAXJ * 3; i += BW_VT) { const int bb = i / (BW_MAXJ * 3); s_gj[bb][i - bb * BW_MAXJ * 3] = bb < nb ? ws.gj[(size_t)(b0 + bb) * BW_MAXJ * 3 + (i - bb * BW_MAXJ * 3)] : 0.0f; }
  for (int i = tid; i < BW_MAXJ; i += BW_VT) s_pick[i] = i < nj ? ws.pick_vid[i] : -1;
  __syncthreads();

  // v_posed of this vertex for the block's bodies
  float vpd[BW_NB][3];
#pragma unroll
  for (int bb = 0; bb < BW_NB; ++bb) vpd[bb][0] = vpd[bb][1] = vpd[bb][2] = 0.0f;
  for (int k = 0; k < 218; ++k) {
    const float x0 = m.blend[((size_t)k * 3 + 0) * vp + v], x1 = m.blend[((size_t)k * 3 + 1) * vp + v], x2 = m.blend[((size_t)k * 3 + 2) * vp + v];
#pragma unroll
    for (int bb = 0; bb < BW_NB; ++bb) {
      const float c = s_coef[bb][k];
      vpd[bb][0] = fmaf(c, x0, vpd[bb][0]); vpd[bb][1] = fmaf(c, x1, vpd[bb][1]); vpd[bb][2] = fmaf(c, x2, vpd[bb][2]);
    }
  }
  // picks that hit this vertex (at most a few of the nj output joints are vertex picks)
  int my_pick[2] = {-1, -1};
  for (int k = 0; k < nj; ++k)
    if (s_pick[k] == v) { if (my_pick[0] < 0) my_pick[0] = k; else my_pick[1] = k; }
  const bool real = v < m.n_verts;
  const int ks = m.ks;
  for (int bb = 0; bb < nb; ++bb) {
    const int b = b0 + bb;
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (real) {
      if (g_verts) { const float* g = g_verts + ((size_t)b * m.n_verts + v) * 3; gx = g[0]; gy = g[1]; gz = g[2]; }
      for (int r = 0; r < nreg; ++r) {
        const float w = jreg[(size_t)r * vp + v];
        gx = fmaf(w, s_ge[bb][r * 3], gx); gy = fmaf(w, s_ge[bb][r * 3 + 1], gy); gz = fmaf(w, s_ge[bb][r * 3 + 2], gz);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q)
        if (my_pick[q] >= 0) { gx += s_gj[bb][my_pick[q] * 3]; gy += s_gj[bb][my_pick[q] * 3 + 1]; gz += s_gj[bb][my_pick[q] * 3 + 2]; }
    }
    // T_R^T g_v
    float t00 = 0, t01 = 0, t02 = 0, t10 = 0, t11 = 0, t12 = 0, t20 = 0, t21 = 0, t22 = 0;
    for (int kk = 0; kk < ks; ++kk) {
      const int j = m.skin_idx[(size_t)v * ks + kk];
      const float w = m.skin_w[(size_t)v * ks + kk];
      const float* A = &s_A[bb][j * 12];
      t00 += w * A[0]; t01 += w * A[1]; t02 += w * A[2]; t10 += w * A[4]; t11 += w * A[5]; t12 += w * A[6]; t20 += w * A[8]; t21 += w * A[9]; t22 += w * A[10];
    }
    ws.gvp[((size_t)b * 3 + 0) * vp + v] = t00 * gx + t10 * gy + t20 * gz;
    ws.gvp[((size_t)b * 3 + 1) * vp + v] = t01 * gx + t11 * gy + t21 * gz;
    ws.gvp[((size_t)b * 3 + 2) * vp + v] = t02 * gx + t12 * gy + t22 * gz;
    // per-joint sums, warp-reduced joint by joint (joints no lane of the warp touches are skipped)
    for (int i = lane; i < 288; i += 32) s_slab[warp][i] = 0.0f;
    __syncwarp();
    for (int j = 0; j < 24; ++j) {
      float w = 0.0f;
      for (int kk = 0; kk < ks; ++kk)
        if (m.skin_idx[(size_t)v * ks + kk] == j) w += m.skin_w[(size_t)v * ks + kk];
      if (__ballot_sync(0xffffffffu, w != 0.0f) == 0u) continue;
      float c[12] = {w * gx * vpd[bb][0], w * gx * vpd[bb][1], w * gx * vpd[bb][2], w * gx,
                     w * gy * vpd[bb][0], w * gy * vpd[bb][1], w * gy * vpd[bb][2], w * gy,
                     w * gz * vpd[bb][0], w * gz * vpd[bb][1], w * gz * vpd[bb][2], w * gz};
#pragma unroll
      for (int e = 0; e < 12; ++e) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) c[e] += __shfl_xor_sync(0xffffffffu, c[e], off);
      }
      if (lane == 0) {
#pragma unroll
        for (int e = 0; e < 12; ++e) s_slab[warp][j * 12 + e] = c[e];
      }
    }
    __syncthreads();
    for (int i = tid; i < 288; i += BW_VT) {
      float s = 0.0f;
#pragma unroll
      for (int wq = 0; wq < BW_VT / 32; ++wq) s += s_slab[wq][i];
      ws.partial[((size_t)b * gridDim.x + blockIdx.x) * 288 + i] = s;
    }
    __syncthreads();
  }
}

// block per body: reduce the per-tile partials, then the kinematic-chain adjoint (thread 0)
__global__ void k_bw_chain(const tp_smpl_model m, int n, const float* __restrict__ R, int tiles, const float* __restrict__ g_R_extra,
                           float* __restrict__ g_R, float* __restrict__ g_betas, BwWs ws) {
  const int b = blockIdx.x, tid = threadIdx.x;
  __shared__ float s_gA[288];
  for (int i = tid; i < 288; i += blockDim.x) {
    float s = 0.0f;
    for (int t = 0; t < tiles; ++t) s += ws.partial[((size_t)b * tiles + t) * 288 + i];
    s_gA[i] = s;
  }
  __syncthreads();
  if (tid != 0) return;
  const float* ch = ws.chain + (size_t)b * BW_CHAIN;
  const float* J = ch; const float* RW = J + 72; const float* rel = RW + 216 + 72;
  const float* Rb = R + (size_t)b * 216;
  float gRW[24 * 9], gt[24 * 3], gJ[24 * 3];
  for (int i = 0; i < 24; ++i) {
    const float* gA = &s_gA[i * 12];          // rows r: [gA_R(r, 0..2) | gA_t(r)]
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) gRW[i * 9 + r * 3 + c] = gA[r * 4 + c] - gA[r * 4 + 3] * J[i * 3 + c];
      gt[i * 3 + r] = gA[r * 4 + 3] + ws.gposed[(size_t)b * 72 + i * 3 + r];
    }
    for (int c = 0; c < 3; ++c)
      gJ[i * 3 + c] = -(RW[i * 9 + c] * gA[3] + RW[i * 9 + 3 + c] * gA[7] + RW[i * 9 + 6 + c] * gA[11]);
  }
  float* gRb = g_R + (size_t)b * 216;
  for (int i = 23; i >= 1; --i) {
    const int p = m.parents[i];
    // gRW_p += gRW_i R_i^T + gt_i rel_i^T
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c)
        gRW[p * 9 + r * 3 + c] += gRW[i * 9 + r * 3] * Rb[i * 9 + c * 3] + gRW[i * 9 + r * 3 + 1] * Rb[i * 9 + c * 3 + 1] +
                                  gRW[i * 9 + r * 3 + 2] * Rb[i * 9 + c * 3 + 2] + gt[i * 3 + r] * rel[i * 3 + c];
    // gR_i = RW_p^T gRW_i ; grel_i = RW_p^T gt_i
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c)
        gRb[i * 9 + r * 3 + c] = RW[p * 9 + r] * gRW[i * 9 + c] + RW[p * 9 + 3 + r] * gRW[i * 9 + 3 + c] + RW[p * 9 + 6 + r] * gRW[i * 9 + 6 + c];
      const float grel = RW[p * 9 + r] * gt[i * 3] + RW[p * 9 + 3 + r] * gt[i * 3 + 1] + RW[p * 9 + 6 + r] * gt[i * 3 + 2];
      gJ[i * 3 + r] += grel;
      gJ[p * 3 + r] -= grel;
    }
    for (int c = 0; c < 3; ++c) gt[p * 3 + c] += gt[i * 3 + c];
  }
  for (int e = 0; e < 9; ++e) gRb[e] = gRW[e];
  for (int c = 0; c < 3; ++c) gJ[c] += gt[c];
  const float* gbl = ws.gbl + (size_t)b * 224;
  for (int i = 1; i < 24; ++i)
    for (int e = 0; e < 9; ++e) gRb[i * 9 + e] += gbl[(i - 1) * 9 + e];
  if (g_R_extra)
    for (int e = 0; e < 216; ++e) gRb[e] += g_R_extra[(size_t)b * 216 + e];
  for (int l = 0; l < 10; ++l) {
    float a = gbl[207 + l];
    for (int i = 0; i < 72; ++i) a += gJ[i] * m.j_shapedirs[i * 10 + l];
    g_betas[b * 10 + l] = a;
  }
}

static size_t align_up_sz(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace tp

using namespace tp;

extern "C" int tp_transpose_f32(const float* src, int64_t ld_src, int rows, int cols, void* dst, int64_t ld_dst, int dst_rows,
                                int dst_precision, int relu, void* stream) {
  TP_CHECK_ARG(src && dst && rows >= 0 && cols >= 0 && ld_src >= cols && ld_dst >= rows && dst_rows >= cols,
               "tp_transpose_f32: bad sizes (rows=%d cols=%d ld_src=%lld ld_dst=%lld dst_rows=%d)", rows, cols, (long long)ld_src,
               (long long)ld_dst, dst_rows);
  if (dst_rows == 0 || ld_dst == 0) return TP_OK;
  dim3 grid((unsigned)ceil_div(dst_rows, 32), (unsigned)ceil_div(ld_dst, 32)), block(32, 8);
  if (dst_precision == TP_PRECISION_BF16)
    k_transpose<__nv_bfloat16><<<grid, block, 0, (cudaStream_t)stream>>>(src, ld_src, rows, cols, (__nv_bfloat16*)dst, ld_dst, dst_rows, relu);
  else
    k_transpose<float><<<grid, block, 0, (cudaStream_t)stream>>>(src, ld_src, rows, cols, (float*)dst, ld_dst, dst_rows, relu);
  TP_LAUNCH_CHECK();
  return TP_OK;
}

extern "C" int tp_colsum_f32(const float* A, int64_t lda, int rows, int cols, float* out, float beta, void* stream) {
  TP_CHECK_ARG(A && out && rows >= 0 && cols >= 0 && lda >= cols, "tp_colsum_f32: bad arguments");
  if (cols == 0) return TP_OK;
  k_colsum<<<(unsigned)ceil_div(cols, 32), dim3(32, 8), 0, (cudaStream_t)stream>>>(A, lda, rows, cols, out, beta);
  TP_LAUNCH_CHECK();
  return TP_OK;
}

extern "C" int tp_mask_scale(float* a, int64_t ld, const float* mask, int64_t ldm, int rows, int cols, float scale, void* stream) {
  TP_CHECK_ARG(a && mask && rows >= 0 && cols >= 0 && ld >= cols && ldm >= cols, "tp_mask_scale: bad arguments");
  if ((int64_t)rows * cols == 0) return TP_OK;
  k_mask_scale<<<(unsigned)ceil_div((int64_t)rows * cols, 256), 256, 0, (cudaStream_t)stream>>>(a, ld, mask, ldm, rows, cols, scale);
  TP_LAUNCH_CHECK();
  return TP_OK;
}

extern "C" int tp_relu_backward(float* g, int64_t ld, const float* h, int64_t ldh, int rows, int cols, void* stream) {
  TP_CHECK_ARG(g && h && rows >= 0 && cols >= 0 && ld >= cols && ldh >= cols, "tp_relu_backward: bad arguments");
  if ((int64_t)rows * cols == 0) return TP_OK;
  k_relu_backward<<<(unsigned)ceil_div((int64_t)rows * cols, 256), 256, 0, (cudaStream_t)stream>>>(g, ld, h, ldh, rows, cols);
  TP_LAUNCH_CHECK();
  return TP_OK;
}

extern "C" int tp_axpby_f32(const float* src, int64_t lds, float* dst, int64_t ldd, int rows, int cols, float alpha, float beta, void* stream) {
  TP_CHECK_ARG(src && dst && rows >= 0 && cols >= 0 && lds >= cols && ldd >= cols, "tp_axpby_f32: bad arguments");
  if ((int64_t)rows * cols == 0) return TP_OK;
  k_axpby<<<(unsigned)ceil_div((int64_t)rows * cols, 256), 256, 0, (cudaStream_t)stream>>>(src, lds, dst, ldd, rows, cols, alpha, beta);
  TP_LAUNCH_CHECK();
  return TP_OK;
}

extern "C" int tp_gru_cell_backward(float* g_h, int64_t ldg, const float* gates, int64_t ld_gates, const float* h_prev, int64_t ldh,
                                    float* d_gi, int64_t ld_dgi, float* d_gh, int64_t ld_dgh, int B, int H, void* stream) {
  TP_CHECK_ARG(g_h && gates && d_gi && d_gh && B >= 1 && H >= 1, "tp_gru_cell_backward: null pointer / empty shape");
  TP_CHECK_ARG(ldg >= H && ld_gates >= 4 * (int64_t)H && ld_dgi >= 3 * (int64_t)H && ld_dgh >= 3 * (int64_t)H && (!h_prev || ldh >= H),
               "tp_gru_cell_backward: leading dimensions too small");
  k_gru_cell_backward<<<(unsigned)ceil_div((int64_t)B * H, 256), 256, 0, (cudaStream_t)stream>>>(g_h, ldg, gates, ld_gates, h_prev, ldh, d_gi,
                                                                                               ld_dgi, d_gh, ld_dgh, B, H);
  TP_LAUNCH_CHECK();
  return TP_OK;
}

extern "C" int tp_rot6d_backward(const float* x6, const float* g_R, float* g_x6, int64_t n, void* stream) {
  TP_CHECK_ARG(n >= 0, "tp_rot6d_backward: n < 0");
  if (n == 0) return TP_OK;
  TP_CHECK_ARG(x6 && g_R && g_x6, "tp_rot6d_backward: null pointer");
  k_rot6d_backward<<<(unsigned)ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(x6, g_R, g_x6, n);
  TP_LAUNCH_CHECK();
  return TP_OK;
}

extern "C" int tp_rotmat_to_angle_axis_backward(const float* R, const float* g_aa, int64_t ld_gaa, int per_row, float* g_R, int64_t n,
                                                int accumulate, void* stream) {
  TP_CHECK_ARG(n >= 0 && per_row >= 1, "tp_rotmat_to_angle_axis_backward: bad sizes");
  if (n == 0) return TP_OK;
  TP_CHECK_ARG(R && g_aa && g_R && ld_gaa >= 3 * (int64_t)per_row, "tp_rotmat_to_angle_axis_backward: null pointer / ld too small");
  k_r2aa_backward<<<(unsigned)ceil_div(n, 64), 64, 0, (cudaStream_t)stream>>>(R, g_aa, ld_gaa, g_R, n, per_row, accumulate);
  TP_LAUNCH_CHECK();
  return TP_OK;
}

extern "C" int tp_gemm_f32_splitk(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, const float* Cin,
                                  int64_t ldcin, float* C, int64_t ldc, int M, int N, int K, float alpha, float beta, int relu_a,
                                  int splits, void* workspace, size_t workspace_bytes, void* stream);
extern "C" size_t tp_gemm_f32_splitk_workspace_bytes(int M, int N, int splits);

static int bw_splits(int n) { return n <= 64 ? 32 : (n <= 256 ? 16 : 4); }

static size_t bw_layout(const tp_smpl_model* m, int n, BwWs* ws, unsigned char* base) {
  const int tiles = m->vp / BW_VT;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up_sz(bytes, 256); return base ? base + o : nullptr; };
  float* chain = (float*)take((size_t)n * BW_CHAIN * 4);
  float* gposed = (float*)take((size_t)n * 72 * 4);
  float* gextra = (float*)take((size_t)n * BW_MAXREG * 3 * 4);
  float* gj = (float*)take((size_t)n * BW_MAXJ * 3 * 4);
  int* pick = (int*)take(BW_MAXJ * 4);
  float* gvp = (float*)take((size_t)n * 3 * m->vp * 4);
  float* partial = (float*)take((size_t)n * tiles * 288 * 4);
  float* gbl = (float*)take((size_t)n * 224 * 4);
  const size_t gws = tp_gemm_f32_splitk_workspace_bytes(n, 217, bw_splits(n));
  void* gemm_ws = take(gws);
  if (ws) { *ws = BwWs{chain, gposed, gextra, gj, pick, gvp, partial, gbl, gemm_ws, gws}; }
  return off;
}

extern "C" size_t tp_smpl_backward_workspace_bytes(const tp_smpl_model* m, int n) {
  if (!m || n <= 0) return 256;
  return bw_layout(m, n, nullptr, nullptr);
}

extern "C" int tp_smpl_backward(const tp_smpl_model* m, int n, const float* R, const float* betas, int64_t ld_betas, const float* cam,
                                int64_t ld_cam, const float* jreg, int nreg, const int32_t* joint_src, int nj, const float* joints,
                                const float* g_verts, const float* g_joints, const float* g_kp2d, const float* g_R_extra,
                                float* g_R, float* g_betas, float* g_cam, void* workspace, size_t workspace_bytes, void* stream) {
  TP_CHECK_ARG(m && n >= 0, "tp_smpl_backward: null model / n < 0");
  if (n == 0) return TP_OK;
  TP_CHECK_ARG(R && betas && g_R && g_betas && workspace, "tp_smpl_backward: null pointer");
  TP_CHECK_ARG(nj >= 0 && nj <= BW_MAXJ && nreg >= 0 && nreg <= BW_MAXREG && (nj == 0 || (joint_src && joints)), "tp_smpl_backward: nj / nreg out of range");
  TP_CHECK_ARG(!g_kp2d || (cam && g_cam), "tp_smpl_backward: g_kp2d needs cam and g_cam");
  TP_CHECK_ARG(m->vp % BW_VT == 0 && m->ks >= 1, "tp_smpl_backward: vp must be a multiple of %d", BW_VT);
  TP_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && workspace_bytes >= tp_smpl_backward_workspace_bytes(m, n),
               "tp_smpl_backward: workspace misaligned or too small");
  cudaStream_t st = (cudaStream_t)stream;
  BwWs ws;
  bw_layout(m, n, &ws, reinterpret_cast<unsigned char*>(workspace));
  const int tiles = m->vp / BW_VT;
  k_bw_prepare<<<n, 64, 0, st>>>(*m, n, R, betas, ld_betas, cam, ld_cam, nreg, joint_src, nj, joints, g_joints, g_kp2d, g_cam, ws);
  TP_LAUNCH_CHECK();
  k_bw_lbs<<<dim3((unsigned)tiles, (unsigned)ceil_div(n, BW_NB)), BW_VT, 0, st>>>(*m, n, R, betas, ld_betas, jreg, nreg, nj, g_verts, ws);
  TP_LAUNCH_CHECK();
  TP_CUDA(cudaMemsetAsync(ws.gemm_ws, 0, 4096, st));          // split-K ticket counters
  int rc = tp_gemm_f32_splitk(ws.gvp, (int64_t)3 * m->vp, m->blend, (int64_t)3 * m->vp, nullptr, nullptr, 0, ws.gbl, 224, n, 217, 3 * m->vp,
                              1.0f, 0.0f, 0, bw_splits(n), ws.gemm_ws, ws.gemm_ws_bytes, stream);
  if (rc != TP_OK) return rc;
  k_bw_chain<<<n, 32, 0, st>>>(*m, n, R, tiles, g_R_extra, g_R, g_betas, ws);
  TP_LAUNCH_CHECK();
  return TP_OK;
}
