// K4 + K5 -- SMPL body-model forward in three launches (replaces ~40 ATen/cuBLAS launches of
// smplx.SMPL.forward / smplx.lbs.lbs (third-party, restated: SURVEY.md App. A.6), the wrapper
// lib/models/smpl.py:72-84, the H36M regression lib/models/spin.py:275-278, projection
// spin.py:280 and the theta assembly spin.py:282-285).
//
//  k_smpl_prepare  one warp per body, lane = joint: pose -> R (rot6d / Rodrigues / pass-through),
//                  J from betas (J_regressor folded into a 24x3x10 table at pack time),
//                  warp-shuffle kinematic-chain scan (level-synchronous over the tree depth),
//                  A_j = [W_R | W_t - W_R J_j], blend coefficients, R -> axis-angle, theta.
//  k_smpl_verts    thread = vertex, NB = 8 bodies per thread in registers: one 218-term
//                  contraction gives template + shape blend + pose blend (K4), then top-k
//                  skinning (K5).  Vertices are written once (coalesced, staged in shared
//                  memory) and the joint regressors are applied to the tile while it is still
//                  on chip -- vertices never round-trip HBM.
//  k_smpl_finalize per body: reduce regressor partials, compose the output joint set, project.
#include "rotations.cuh"
#include "skinny.cuh"
#include "umma.cuh"

namespace tp {

constexpr int kJ = 24;          // SMPL joints
constexpr int kCoef = 218;      // 207 pose-blend + 10 shape + 1 template
constexpr int kCoefLd = 224;    // row pitch of the coefficient scratch
constexpr int kVT = 128;        // vertices per tile (= threads per CTA in k_smpl_verts)
constexpr int kNB = 8;          // bodies per CTA in k_smpl_verts
constexpr int kMaxReg = 32;     // max rows of the caller's joint regressor

__device__ __forceinline__ void project_point_(const float* X, const float* cam, float* kp) {
  // lib/models/spin.py:307-351 (same operation order, see geometry.cu)
  float tz = 2.0f * 5000.0f / (224.0f * cam[0] + 1e-9f);
  float px = X[0] + cam[1], py = X[1] + cam[2], pz = X[2] + tz;
  px = px / pz; py = py / pz;
  kp[0] = (5000.0f * px) / 112.0f;
  kp[1] = (5000.0f * py) / 112.0f;
}

struct PrepArgs {
  const float* pose; int64_t ld_pose; int pose_kind;
  const float* betas; int64_t ld_betas;
  const float* cam; int64_t ld_cam;
  float* A;        // [n][24][12]
  float* posedJ;   // [n][24][3]
  float* coef;     // [n][kCoefLd]
  float* rotmat;   // [n][24][9] or null
  float* theta;    // [n][85] or null
  __nv_bfloat16* coef_tc;  // [n][256] bf16: pf(207) | beta_hi(10) | beta_lo(10) | beta_hi(10) | 0   (tensor-core path) or null
  unsigned char* coef_um;  // the same rows as the tcgen05 B-operand image [group of um_rows bodies][K block 4][row][128 B, 16-byte chunks
                           // XOR-swizzled by row & 7] (k_smpl_lbs_um: 32 rows, k_smpl_lbs_um2: 16) or null; bodies n..n_pad-1 get zero
                           // rows and zero transforms
  __nv_bfloat16* coef_cat; // [n][512] bf16 = [row | row] of coef_tc's rows (the fold GEMM contracts it with [M_hi | M_lo]) or null
  unsigned char* timg;     // k_smpl_lbs_um2: the joint transforms as the B operand of the skinning MMA: [8 bodies][row (body, e) 96][128 B]
                           // with k = joint: A_hi at k 0..23, A_lo at k 32..55, zero pads, same chunk swizzle; or null
  int n_pad, um_rows;
};

// byte offset of coefficient k of body b inside the tcgen05 B-operand image
__device__ __forceinline__ size_t coef_um_offset(int b, int k, int rows) {
  const int grp = b / rows, row = b - grp * rows;
  return (size_t)grp * rows * 512 + (size_t)(k >> 6) * rows * 128 + (size_t)row * 128 + (size_t)((((k & 63) >> 3) ^ (row & 7)) << 4) + (size_t)(k & 7) * 2;
}
// byte offset of (entry e of joint j, hi / lo) of body b inside the skinning-MMA operand image
__device__ __forceinline__ size_t timg_offset(int b, int e, int j, int lo) {
  const int n = (b & 7) * 12 + e, k = j + 32 * lo;
  return (size_t)(b >> 3) * 12288 + (size_t)n * 128 + (size_t)(((k >> 3) ^ (n & 7)) << 4) + (size_t)(k & 7) * 2;
}

__global__ void __launch_bounds__(128) k_smpl_prepare(const tp_smpl_model m, int n, const PrepArgs a) {
  // programmatic dependent launch: the vertex kernel may start now and load its resident blend tile while this
  // kernel runs; it waits (griddepcontrol.wait) before touching anything written here
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  pdl_wait();                                                             // pose / betas / cam may come from a PDL predecessor (the IEF kernel)
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= n) {                             // whole warp exits together
    if (a.coef_um && b < a.n_pad) {         // padding bodies of the last group: zero operand rows, zero transforms
      *reinterpret_cast<uint4*>(a.coef_um + coef_um_offset(b, (lane >> 3) * 64, a.um_rows) - (size_t)((b % a.um_rows) & 7) * 16 + (size_t)(lane & 7) * 16) =
          make_uint4(0u, 0u, 0u, 0u);
      for (int i = lane; i < kJ * 12; i += 32) a.A[(int64_t)b * kJ * 12 + i] = 0.0f;
      if (a.timg)
        for (int i = lane; i < 96; i += 32)
          *reinterpret_cast<uint4*>(a.timg + (size_t)(b >> 3) * 12288 + (size_t)((b & 7) * 12) * 128 + (size_t)i * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    return;
  }
  const int j = lane < kJ ? lane : kJ - 1;  // idle lanes shadow joint 23 (never written)
  const bool active = lane < kJ;

  float R[9];
  if (a.pose_kind == TP_POSE_ROTMAT) {
    const float* p = a.pose + (int64_t)b * a.ld_pose + j * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = p[k];
  } else if (a.pose_kind == TP_POSE_AXIS_ANGLE) {
    const float* p = a.pose + (int64_t)b * a.ld_pose + j * 3;
    float r[3] = {p[0], p[1], p[2]};
    rodrigues_smplx(r, R);
  } else {
    const float* p = a.pose + (int64_t)b * a.ld_pose + j * 6;
    float x[6] = {p[0], p[1], p[2], p[3], p[4], p[5]};
    rot6d_to_rotmat(x, R);
  }
  float beta[10];
#pragma unroll
  for (int l = 0; l < 10; ++l) beta[l] = a.betas[(int64_t)b * a.ld_betas + l];

  // rest-pose joint location: J = J_regressor.(v_template + shapedirs.beta), regressor pre-applied
  float Jx[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float s = m.j_template[j * 3 + c];
#pragma unroll
    for (int l = 0; l < 10; ++l) s = fmaf(m.j_shapedirs[(j * 3 + c) * 10 + l], beta[l], s);
    Jx[c] = s;
  }
  const int parent = m.parents[j];
  // depth in the kinematic tree by walking up through SHUFFLES (the table lives in the lanes already;
  // chasing m.parents[] through global memory costs one dependent L2 round trip per level)
  int depth = 0;
  {
    int anc = parent;
#pragma unroll 1
    for (int it = 0; it < kJ; ++it) {
      const bool up = anc >= 0;
      const int nxt = __shfl_sync(0xffffffffu, parent, up ? anc : 0);
      depth += up ? 1 : 0;
      anc = up ? nxt : -1;
      if (__all_sync(0xffffffffu, anc < 0)) break;
    }
  }
  int maxd = depth;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) maxd = max(maxd, __shfl_xor_sync(0xffffffffu, maxd, o));
  const int src = parent < 0 ? 0 : parent;

  // world transform W = [WR | Wt]; root: G_0 = [R_0 | J_0]
  float WR[9], Wt[3], rel[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float pj = __shfl_sync(0xffffffffu, Jx[c], src);
    rel[c] = parent < 0 ? Jx[c] : Jx[c] - pj;
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) WR[k] = R[k];
#pragma unroll
  for (int c = 0; c < 3; ++c) Wt[c] = rel[c];
  for (int d = 1; d <= maxd; ++d) {
    float PR[9], Pt[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) PR[k] = __shfl_sync(0xffffffffu, WR[k], src);
#pragma unroll
    for (int c = 0; c < 3; ++c) Pt[c] = __shfl_sync(0xffffffffu, Wt[c], src);
    if (depth == d) {  // W_i = W_parent @ G_i
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
          WR[r * 3 + c] = PR[r * 3 + 0] * R[0 * 3 + c] + PR[r * 3 + 1] * R[1 * 3 + c] + PR[r * 3 + 2] * R[2 * 3 + c];
        Wt[r] = PR[r * 3 + 0] * rel[0] + PR[r * 3 + 1] * rel[1] + PR[r * 3 + 2] * rel[2] + Pt[r];
      }
    }
  }
  float Av[12];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    Av[r * 4 + 0] = WR[r * 3 + 0];
    Av[r * 4 + 1] = WR[r * 3 + 1];
    Av[r * 4 + 2] = WR[r * 3 + 2];
    Av[r * 4 + 3] = Wt[r] - (WR[r * 3 + 0] * Jx[0] + WR[r * 3 + 1] * Jx[1] + WR[r * 3 + 2] * Jx[2]);
  }
  if (a.timg) {
    // ---- large-batch tcgen05 path: the body's operand rows are assembled in shared memory (2-byte stores) and leave as 16-byte
    // chunks: 32 for the coefficient row (one per K block and chunk position), 96 contiguous ones for the 12 transform rows
    __shared__ __align__(16) unsigned char stg_all[4][2048];
    unsigned char* stg = stg_all[threadIdx.x >> 5];
    __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(stg);            // [256] logical order
    unsigned char* trow = stg + 512;                                         // [12 rows][128 B] in its final (swizzled) layout
    const int n0 = (b & 7) * 12;
    if (active) {
      if (j >= 1) {
#pragma unroll
        for (int k = 0; k < 9; ++k) crow[(j - 1) * 9 + k] = __float2bfloat16_rn(R[k] - ((k == 0 || k == 4 || k == 8) ? 1.0f : 0.0f));
      } else {
#pragma unroll
        for (int l = 0; l < 10; ++l) {
          const __nv_bfloat16 hi = __float2bfloat16_rn(beta[l]);
          crow[207 + l] = hi; crow[217 + l] = __float2bfloat16_rn(beta[l] - __bfloat162float(hi)); crow[227 + l] = hi;
        }
        for (int k = 237; k < 256; ++k) crow[k] = __float2bfloat16_rn(0.0f);
      }
#pragma unroll
      for (int e = 0; e < 12; ++e) {
        const int n = n0 + e;
        const __nv_bfloat16 hi = __float2bfloat16_rn(Av[e]);
        unsigned char* row = trow + e * 128;
        *reinterpret_cast<__nv_bfloat16*>(row + ((((j >> 3)) ^ (n & 7)) << 4) + (j & 7) * 2) = hi;
        *reinterpret_cast<__nv_bfloat16*>(row + (((4 + (j >> 3)) ^ (n & 7)) << 4) + (j & 7) * 2) = __float2bfloat16_rn(Av[e] - __bfloat162float(hi));
      }
    } else {
      for (int i = lane - kJ; i < 24; i += 8) {            // zero pads: chunks 3 and 7 (k 24..31, 56..63) of the 12 rows
        const int e = i >> 1, n = n0 + e, chunk = (i & 1) ? 7 : 3;
        *reinterpret_cast<uint4*>(trow + e * 128 + ((chunk ^ (n & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    __syncwarp();
    {
      const int rows = a.um_rows, grp = b / rows, row = b - grp * rows, kb = lane >> 3, c = lane & 7;
      const uint4 v = *reinterpret_cast<const uint4*>(stg + kb * 128 + c * 16);
      *reinterpret_cast<uint4*>(a.coef_um + (size_t)grp * rows * 512 + (size_t)kb * rows * 128 + (size_t)row * 128 + (size_t)((c ^ (row & 7)) << 4)) = v;
      if (a.coef_cat) {                                     // [row | row]: A operand of the folded-regressor GEMM (K = 512)
        uint4* d = reinterpret_cast<uint4*>(a.coef_cat + (int64_t)b * 512);
        d[lane] = v; d[32 + lane] = v;
      }
      uint4* td = reinterpret_cast<uint4*>(a.timg + (size_t)(b >> 3) * 12288 + (size_t)n0 * 128);
#pragma unroll
      for (int i = 0; i < 3; ++i) td[lane + 32 * i] = *reinterpret_cast<const uint4*>(trow + (lane + 32 * i) * 16);
    }
  }
  if (!active) return;

  float* Ab = a.A + ((int64_t)b * kJ + j) * 12;
#pragma unroll
  for (int e = 0; e < 12; ++e) Ab[e] = Av[e];
#pragma unroll
  for (int r = 0; r < 3; ++r) a.posedJ[((int64_t)b * kJ + j) * 3 + r] = Wt[r];
  if (a.coef) {
    float* cf = a.coef + (int64_t)b * kCoefLd;
    if (j >= 1) {
#pragma unroll
      for (int k = 0; k < 9; ++k) cf[(j - 1) * 9 + k] = R[k] - ((k == 0 || k == 4 || k == 8) ? 1.0f : 0.0f);
    } else {
#pragma unroll
      for (int l = 0; l < 10; ++l) cf[207 + l] = beta[l];
      cf[217] = 1.0f;
      for (int k = kCoef; k < kCoefLd; ++k) cf[k] = 0.0f;
    }
  }
  if (a.coef_tc) {
    __nv_bfloat16* ct = a.coef_tc + (int64_t)b * 256;
    if (j >= 1) {
#pragma unroll
      for (int k = 0; k < 9; ++k) ct[(j - 1) * 9 + k] = __float2bfloat16_rn(R[k] - ((k == 0 || k == 4 || k == 8) ? 1.0f : 0.0f));
    } else {
#pragma unroll
      for (int l = 0; l < 10; ++l) {
        const __nv_bfloat16 hi = __float2bfloat16_rn(beta[l]);
        const __nv_bfloat16 lo = __float2bfloat16_rn(beta[l] - __bfloat162float(hi));
        ct[207 + l] = hi; ct[217 + l] = lo; ct[227 + l] = hi;
      }
      for (int k = 237; k < 256; ++k) ct[k] = __float2bfloat16_rn(0.0f);
    }
  }
  if (a.coef_um && !a.timg) {                               // first-generation tcgen05 path (k_smpl_lbs_um): scattered element stores
    if (j >= 1) {
#pragma unroll
      for (int k = 0; k < 9; ++k)
        *reinterpret_cast<__nv_bfloat16*>(a.coef_um + coef_um_offset(b, (j - 1) * 9 + k, a.um_rows)) =
            __float2bfloat16_rn(R[k] - ((k == 0 || k == 4 || k == 8) ? 1.0f : 0.0f));
    } else {
#pragma unroll
      for (int l = 0; l < 10; ++l) {
        const __nv_bfloat16 hi = __float2bfloat16_rn(beta[l]);
        const __nv_bfloat16 lo = __float2bfloat16_rn(beta[l] - __bfloat162float(hi));
        *reinterpret_cast<__nv_bfloat16*>(a.coef_um + coef_um_offset(b, 207 + l, a.um_rows)) = hi;
        *reinterpret_cast<__nv_bfloat16*>(a.coef_um + coef_um_offset(b, 217 + l, a.um_rows)) = lo;
        *reinterpret_cast<__nv_bfloat16*>(a.coef_um + coef_um_offset(b, 227 + l, a.um_rows)) = hi;
      }
      for (int k = 237; k < 256; ++k) *reinterpret_cast<__nv_bfloat16*>(a.coef_um + coef_um_offset(b, k, a.um_rows)) = __float2bfloat16_rn(0.0f);
    }
  }
  if (a.rotmat) {
#pragma unroll
    for (int k = 0; k < 9; ++k) a.rotmat[((int64_t)b * kJ + j) * 9 + k] = R[k];
  }
  if (a.theta) {
    float aa[3];
    rotmat_to_angle_axis(R, aa);
    float* th = a.theta + (int64_t)b * 85;
    th[3 + j * 3 + 0] = aa[0]; th[3 + j * 3 + 1] = aa[1]; th[3 + j * 3 + 2] = aa[2];
    if (j == 0) {
#pragma unroll
      for (int l = 0; l < 10; ++l) th[75 + l] = beta[l];
      if (a.cam) { th[0] = a.cam[(int64_t)b * a.ld_cam]; th[1] = a.cam[(int64_t)b * a.ld_cam + 1]; th[2] = a.cam[(int64_t)b * a.ld_cam + 2]; }
      else { th[0] = th[1] = th[2] = 0.0f; }
    }
  }
}

// shared-memory carve-up of k_smpl_verts (floats)
constexpr int kAPitch = 13;                              // joint stride in A_s (12 + 1: conflict-free gathers)
constexpr int kVsPitch = 28;                             // per-vertex stride of the transposed tile
constexpr int kSmCoef = kCoef * kNB;                     // coef_s[k][b]
constexpr int kSmA = kNB * kJ * kAPitch;                 // A_s[b][j][13]
constexpr int kSmVsT = kVT * kVsPitch;                   // vsT[v][c*NB + b]
constexpr int kSmVout = kNB * kVT * 3;                   // vout[b][v*3 + c]
constexpr int kSmPart = 4 * kMaxReg * kNB * 3;           // part[q][r][c*NB + b]
constexpr int kSmVertsFloats = kSmCoef + kSmA + kSmVsT + kSmVout + kSmPart;

__global__ void __launch_bounds__(kVT) k_smpl_verts(const tp_smpl_model m, int n, const float* __restrict__ coef,
                                                    const float* __restrict__ A, const float* __restrict__ jreg,
                                                    int nreg, float* __restrict__ verts, float* __restrict__ jpart,
                                                    int nsplit, int tiles_per_split, int ntiles) {
  extern __shared__ __align__(16) float sm[];
  float* coef_s = sm;
  float* A_s = coef_s + kSmCoef;
  float* vsT = A_s + kSmA;
  float* vout = vsT + kSmVsT;
  float* part = vout + kSmVout;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int body0 = blockIdx.x * kNB;
  const int split = blockIdx.y;
  const int vp = m.vp;

  for (int i = tid; i < kCoef * kNB; i += kVT) {
    int k = i / kNB, b = i - k * kNB;
    coef_s[i] = (body0 + b < n) ? coef[(int64_t)(body0 + b) * kCoefLd + k] : 0.0f;
  }
  for (int i = tid; i < kNB * kJ * 12; i += kVT) {
    int b = i / (kJ * 12), r = i - b * (kJ * 12), j = r / 12, e = r - j * 12;
    A_s[(b * kJ + j) * kAPitch + e] = (body0 + b < n) ? A[(int64_t)(body0 + b) * kJ * 12 + r] : 0.0f;
  }
  __syncthreads();

  float accj[kNB * 3];
#pragma unroll
  for (int i = 0; i < kNB * 3; ++i) accj[i] = 0.0f;

  const int tile_lo = split * tiles_per_split;
  const int tile_hi = min(ntiles, tile_lo + tiles_per_split);
  for (int tile = tile_lo; tile < tile_hi; ++tile) {
    const int v = tile * kVT + tid;  // < vp always (blend / skin tables are padded to vp)
    float acc[kNB][3];
#pragma unroll
    for (int b = 0; b < kNB; ++b) acc[b][0] = acc[b][1] = acc[b][2] = 0.0f;
    const float* bl = m.blend + v;
#pragma unroll 8   // 24 independent loads in flight per thread: this loop is latency-bound otherwise
    for (int k = 0; k < kCoef; ++k) {
      float d0 = __ldg(bl + (int64_t)(k * 3 + 0) * vp);
      float d1 = __ldg(bl + (int64_t)(k * 3 + 1) * vp);
      float d2 = __ldg(bl + (int64_t)(k * 3 + 2) * vp);
      const float4 c0 = *reinterpret_cast<const float4*>(&coef_s[k * kNB]);
      const float4 c1 = *reinterpret_cast<const float4*>(&coef_s[k * kNB + 4]);
      const float cc[kNB] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
      for (int b = 0; b < kNB; ++b) {
        acc[b][0] = fmaf(cc[b], d0, acc[b][0]);
        acc[b][1] = fmaf(cc[b], d1, acc[b][1]);
        acc[b][2] = fmaf(cc[b], d2, acc[b][2]);
      }
    }
    // linear blend skinning with the retained (top-k) weights of this vertex
    float T[kNB][12];
#pragma unroll
    for (int b = 0; b < kNB; ++b)
#pragma unroll
      for (int e = 0; e < 12; ++e) T[b][e] = 0.0f;
    for (int i = 0; i < m.ks; ++i) {
      const int jj = m.skin_idx[(int64_t)v * m.ks + i];
      const float w = m.skin_w[(int64_t)v * m.ks + i];
#pragma unroll
      for (int b = 0; b < kNB; ++b) {
        const float* Aj = &A_s[(b * kJ + jj) * kAPitch];
#pragma unroll
        for (int e = 0; e < 12; ++e) T[b][e] = fmaf(w, Aj[e], T[b][e]);
      }
    }
    const bool vvalid = v < m.n_verts;
    float o[kNB][3];
#pragma unroll
    for (int b = 0; b < kNB; ++b) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        float val = T[b][r * 4 + 0] * acc[b][0] + T[b][r * 4 + 1] * acc[b][1] + T[b][r * 4 + 2] * acc[b][2] + T[b][r * 4 + 3];
        o[b][r] = vvalid ? val : 0.0f;
        vout[(b * kVT + tid) * 3 + r] = o[b][r];
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      *reinterpret_cast<float4*>(&vsT[tid * kVsPitch + c * kNB]) = make_float4(o[0][c], o[1][c], o[2][c], o[3][c]);
      *reinterpret_cast<float4*>(&vsT[tid * kVsPitch + c * kNB + 4]) = make_float4(o[4][c], o[5][c], o[6][c], o[7][c]);
    }
    __syncthreads();
    if (verts) {
      const int64_t nv3 = (int64_t)m.n_verts * 3;
      const int64_t base = (int64_t)tile * kVT * 3;
#pragma unroll
      for (int b = 0; b < kNB; ++b) {
        if (body0 + b >= n) break;
        float* dst = verts + (int64_t)(body0 + b) * nv3 + base;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          int idx = tid + i * kVT;
          if (base + idx < nv3) dst[idx] = vout[b * kVT * 3 + idx];
        }
      }
    }
    if (lane < nreg) {  // joint regressors on the on-chip tile: lane = regressor row, warp = vertex quarter
      const float* jr = jreg + (int64_t)lane * vp + tile * kVT + warp * 32;
      for (int vv = 0; vv < 32; ++vv) {
        const float w = __ldg(jr + vv);
        const float* vs = &vsT[(warp * 32 + vv) * kVsPitch];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const float4 x = *reinterpret_cast<const float4*>(vs + q * 4);
          accj[q * 4 + 0] = fmaf(w, x.x, accj[q * 4 + 0]);
          accj[q * 4 + 1] = fmaf(w, x.y, accj[q * 4 + 1]);
          accj[q * 4 + 2] = fmaf(w, x.z, accj[q * 4 + 2]);
          accj[q * 4 + 3] = fmaf(w, x.w, accj[q * 4 + 3]);
        }
      }
    }
    __syncthreads();
  }
  if (nreg > 0) {
    if (lane < nreg) {
#pragma unroll
      for (int i = 0; i < kNB * 3; ++i) part[(warp * kMaxReg + lane) * kNB * 3 + i] = accj[i];
    }
    __syncthreads();
    for (int i = tid; i < nreg * kNB * 3; i += kVT) {
      int r = i / (kNB * 3), rem = i - r * (kNB * 3), c = rem / kNB, b = rem - c * kNB;
      if (body0 + b >= n) continue;
      float s = 0.0f;
#pragma unroll
      for (int q = 0; q < 4; ++q) s += part[(q * kMaxReg + r) * kNB * 3 + rem];
      jpart[(((int64_t)(body0 + b) * nsplit + split) * nreg + r) * 3 + c] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Tensor-core variant of the vertex kernel (bf16 mode; K4 of the north star): the 256-term
// contraction  [64 vertices x 3 coords] x [K = 256] x [32 bodies]  runs on mma.sync m16n8k16 with the
// vertex tile's blend fragments RESIDENT in shared memory (96 KB, loaded once per CTA and reused
// for every 32-body group the CTA walks through).  K layout: 207 pose-blend rows | 10 shape rows
// (beta_hi x S_hi) | 10 (beta_lo x S_hi) | 10 (beta_hi x S_lo) -- the shape blend keeps ~16 mantissa
// bits; the template is added in fp32.  Warp w: 16-vertex sub-tile w&3, body half w>>2; a thread ends up
// with x,y,z of 2 vertices x 4 bodies and skins them from registers.
constexpr int kTcVT = 64;                    // vertices per CTA
constexpr int kTcNB = 32;                    // bodies per group
constexpr int kTcK = 256;
constexpr int kTcCsPitch = kTcK + 32;        // bf16 elements per coefficient row (conflict-free 16-byte reads)
constexpr int kTcAStride = kJ * 12 + 4;      // floats per body in A_s
constexpr int kTcJtPitch = kTcVT + 4;     // 16-byte aligned rows: the regressor loop reads float4 (broadcast)
constexpr int kTcVoPitch = kTcVT * 3 + 4;  // floats per body in vout: 196 = 4 (mod 32) -> conflict-free float4 reads with lane = body
constexpr size_t kTcSmW = 12 * 8 * 1024;
constexpr size_t kTcSmC = (size_t)kTcNB * kTcCsPitch * 2;
constexpr size_t kTcSmA = (size_t)kTcNB * kTcAStride * 4;
constexpr size_t kTcSmV = (size_t)kTcNB * kTcVoPitch * 4;
constexpr size_t kTcSmJ = (size_t)kMaxReg * kTcJtPitch * 4;
constexpr size_t kTcSmem = kTcSmW + kTcSmC + kTcSmA + kTcSmV + kTcSmJ;

__global__ void __launch_bounds__(256, 1)
k_smpl_verts_tc(const tp_smpl_model m, int n, const __nv_bfloat16* __restrict__ coef_tc, const float* __restrict__ A,
                const float* __restrict__ jreg, int nreg, float* __restrict__ verts, float* __restrict__ jpart,
                int ntiles, int groups_per_cta) {
  extern __shared__ __align__(16) unsigned char tsm[];
  unsigned char* Wt = tsm;
  __nv_bfloat16* Cs = reinterpret_cast<__nv_bfloat16*>(tsm + kTcSmW);
  float* A_s = reinterpret_cast<float*>(tsm + kTcSmW + kTcSmC);
  float* vout = reinterpret_cast<float*>(tsm + kTcSmW + kTcSmC + kTcSmA);
  float* Jt = reinterpret_cast<float*>(tsm + kTcSmW + kTcSmC + kTcSmA + kTcSmV);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int st = warp & 3, nh = warp >> 2;
  const int tile = blockIdx.x;
  const int v0 = tile * kTcVT;

  // resident operands of this vertex tile: blend fragments (12 m-tiles x 8 blocks), regressor columns
  {
    const uint4* src = reinterpret_cast<const uint4*>(m.blend_tc) + (size_t)tile * (kTcSmW / 16);
    uint4* dst = reinterpret_cast<uint4*>(Wt);
    for (int i = tid; i < (int)(kTcSmW / 16); i += 256) dst[i] = ldg_stream16(src + i);
    for (int i = tid; i < nreg * kTcVT; i += 256) {
      const int r = i / kTcVT, v = i - r * kTcVT;
      Jt[r * kTcJtPitch + v] = jreg[(int64_t)r * m.vp + v0 + v];
    }
  }
  // per-thread vertex constants: template, retained skinning weights (ks <= 4 on this path)
  const int vA = v0 + st * 16 + g, vB = vA + 8;
  float tmpl[2][3];
  int sj[2][4]; float sw[2][4];
#pragma unroll
  for (int x = 0; x < 2; ++x) {
    const int v = x ? vB : vA;
#pragma unroll
    for (int c = 0; c < 3; ++c) tmpl[x][c] = m.template_pad[(int64_t)v * 3 + c];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      sj[x][k] = k < m.ks ? m.skin_idx[(int64_t)v * m.ks + k] : 0;
      sw[x][k] = k < m.ks ? m.skin_w[(int64_t)v * m.ks + k] : 0.0f;
    }
  }

  // everything above reads model constants only; the coefficients / transforms come from k_smpl_prepare
  asm volatile("griddepcontrol.wait;\n" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");      // k_smpl_finalize may be scheduled; it waits for this grid
  const int ngroups = (n + kTcNB - 1) / kTcNB;
  const int g_lo = blockIdx.y * groups_per_cta, g_hi = min(ngroups, g_lo + groups_per_cta);
  for (int grp = g_lo; grp < g_hi; ++grp) {
    const int body0 = grp * kTcNB;
    // (1) coefficients (bf16 rows) and skinning transforms of this body group
    for (int i = tid; i < kTcNB * (kTcK / 8); i += 256) {
      const int b = i / (kTcK / 8), c8 = i - b * (kTcK / 8);
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (body0 + b < n) v = __ldg(reinterpret_cast<const uint4*>(coef_tc + (int64_t)(body0 + b) * kTcK) + c8);
      *reinterpret_cast<uint4*>(Cs + (size_t)b * kTcCsPitch + c8 * 8) = v;
    }
    for (int i = tid; i < kTcNB * (kJ * 3); i += 256) {
      const int b = i / (kJ * 3), q = i - b * (kJ * 3);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (body0 + b < n) v = __ldg(reinterpret_cast<const float4*>(A + (int64_t)(body0 + b) * kJ * 12) + q);
      *reinterpret_cast<float4*>(A_s + (size_t)b * kTcAStride + q * 4) = v;
    }
    __syncthreads();
    // (2) blend contraction on tensor cores
    float acc[3][2][4];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int nn = 0; nn < 2; ++nn) acc[c][nn][0] = acc[c][nn][1] = acc[c][nn][2] = acc[c][nn][3] = 0.0f;
#pragma unroll
    for (int kb = 0; kb < kTcK / 32; ++kb) {
      uint4 bv[2];
#pragma unroll
      for (int nn = 0; nn < 2; ++nn)
        bv[nn] = *reinterpret_cast<const uint4*>(Cs + (size_t)((nh * 2 + nn) * 8 + g) * kTcCsPitch + kb * 32 + 8 * t);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const unsigned char* wp = Wt + (size_t)((st * 3 + c) * 8 + kb) * 1024 + lane * 16;
        const uint4 wa = *reinterpret_cast<const uint4*>(wp), wb = *reinterpret_cast<const uint4*>(wp + 512);
#pragma unroll
        for (int nn = 0; nn < 2; ++nn) {
          mma16816(acc[c][nn], wa, bv[nn].x, bv[nn].y);
          mma16816(acc[c][nn], wb, bv[nn].z, bv[nn].w);
        }
      }
    }
    // skinning from registers: element e of acc[c][nn] is (vertex x = e>>1, body (nh*2+nn)*8 + 2t + (e&1))
#pragma unroll
    for (int nn = 0; nn < 2; ++nn)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int x = e >> 1, bl = (nh * 2 + nn) * 8 + 2 * t + (e & 1);
        const float px = acc[0][nn][e] + tmpl[x][0], py = acc[1][nn][e] + tmpl[x][1], pz = acc[2][nn][e] + tmpl[x][2];
        float T[12];
#pragma unroll
        for (int q = 0; q < 12; ++q) T[q] = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4* Aj = reinterpret_cast<const float4*>(A_s + (size_t)bl * kTcAStride + sj[x][k] * 12);
          const float4 r0 = Aj[0], r1 = Aj[1], r2 = Aj[2];
          const float w = sw[x][k];
          T[0] = fmaf(w, r0.x, T[0]); T[1] = fmaf(w, r0.y, T[1]); T[2] = fmaf(w, r0.z, T[2]); T[3] = fmaf(w, r0.w, T[3]);
          T[4] = fmaf(w, r1.x, T[4]); T[5] = fmaf(w, r1.y, T[5]); T[6] = fmaf(w, r1.z, T[6]); T[7] = fmaf(w, r1.w, T[7]);
          T[8] = fmaf(w, r2.x, T[8]); T[9] = fmaf(w, r2.y, T[9]); T[10] = fmaf(w, r2.z, T[10]); T[11] = fmaf(w, r2.w, T[11]);
        }
        const int vloc = st * 16 + g + 8 * x;
        const bool vvalid = (v0 + vloc) < m.n_verts;
        float* o = vout + (size_t)bl * kTcVoPitch + vloc * 3;
        o[0] = vvalid ? (T[0] * px + T[1] * py + T[2] * pz + T[3]) : 0.0f;
        o[1] = vvalid ? (T[4] * px + T[5] * py + T[6] * pz + T[7]) : 0.0f;
        o[2] = vvalid ? (T[8] * px + T[9] * py + T[10] * pz + T[11]) : 0.0f;
      }
    __syncthreads();
    // (3) coalesced vertex store (768 contiguous bytes per body) + joint regressors on the on-chip tile
    if (verts) {
      const int64_t nv3 = (int64_t)m.n_verts * 3;
      const int64_t base = (int64_t)v0 * 3;
      for (int i = tid; i < kTcNB * (kTcVT * 3 / 2); i += 256) {
        const int b = i / (kTcVT * 3 / 2), q = i - b * (kTcVT * 3 / 2);
        if (body0 + b >= n) continue;
        const float2 v = *reinterpret_cast<const float2*>(vout + (size_t)b * kTcVoPitch + q * 2);
        float* dst = verts + (int64_t)(body0 + b) * nv3 + base + q * 2;
        if (base + q * 2 + 1 < nv3) *reinterpret_cast<float2*>(dst) = v;
        else if (base + q * 2 < nv3) dst[0] = v.x;
      }
    }
    // joint regressors: lane = body, each warp takes every 8th regressor row; the weights are a broadcast float4,
    // the body's vertices a conflict-free float4 (body pitch 196 floats) -- every shared-memory wavefront is full
    for (int r = warp; r < nreg; r += 8) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f;
      const float4* jr = reinterpret_cast<const float4*>(Jt + r * kTcJtPitch);
      const float4* vb = reinterpret_cast<const float4*>(vout + (size_t)lane * kTcVoPitch);
#pragma unroll 4
      for (int v4 = 0; v4 < kTcVT / 4; ++v4) {
        const float4 w = jr[v4];
        const float4 p0 = vb[3 * v4], p1 = vb[3 * v4 + 1], p2 = vb[3 * v4 + 2];      // x0 y0 z0 x1 | y1 z1 x2 y2 | z2 x3 y3 z3
        s0 = fmaf(w.x, p0.x, s0); s1 = fmaf(w.x, p0.y, s1); s2 = fmaf(w.x, p0.z, s2);
        s0 = fmaf(w.y, p0.w, s0); s1 = fmaf(w.y, p1.x, s1); s2 = fmaf(w.y, p1.y, s2);
        s0 = fmaf(w.z, p1.z, s0); s1 = fmaf(w.z, p1.w, s1); s2 = fmaf(w.z, p2.x, s2);
        s0 = fmaf(w.w, p2.y, s0); s1 = fmaf(w.w, p2.z, s1); s2 = fmaf(w.w, p2.w, s2);
      }
      if (body0 + lane < n) {
        float* dst = jpart + (((int64_t)(body0 + lane) * ntiles + tile) * nreg + r) * 3;
        dst[0] = s0; dst[1] = s1; dst[2] = s2;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// Large-batch path (config 4: SMPL standalone at 10^4..10^5 bodies).  The blend contraction
// [bodies, 256] x [256, 3 vp] runs as a tcgen05 GEMM (tp_gemm_bf16_tc, template as the bias) over a CHUNK of
// bodies whose fp32 result (85 MB per 1024 bodies) is consumed while most of it is still in L2, and this kernel skins it straight out of L2:
// thread = vertex, bodies looped.  All 32 lanes of a warp work on the SAME body, so the per-vertex gathers of the
// joint transforms hit distinct banks for distinct joints (row pitch 13) or broadcast -- 1.5 wavefronts per
// (vertex, body), the minimum for 4 x 48 bytes.  Vertices are written once, coalesced; the joint regressors are
// applied to the tile while it is in shared memory (lane = body, every regressor row per pass).
constexpr int kSkVT = 128;                         // vertices per CTA
#ifndef TP_SK_GB
#define TP_SK_GB 16
#endif
constexpr int kSkGB = TP_SK_GB;                    // bodies per stage (8 or 16)
constexpr int kSkAPitch = 13;                      // floats per joint row: bank (13 j + q) mod 32 is distinct for distinct joints (scalar loads);
                                                   // measured: pitch 14 + 8-byte loads has 42 % conflict wavefronts and costs a CTA per SM
constexpr int kSkABody = kJ * kSkAPitch;           // floats per body in A_s
constexpr int kSkVoPitch = kSkVT * 3 + 4;          // floats per body in vout (388 = 4 mod 32)
constexpr int kSkJtPitch = kSkVT + 4;
constexpr int kSkRC = 9;                           // regressor rows per accumulation pass
constexpr size_t kSkSmA = (size_t)2 * kSkGB * kSkABody * 4;
constexpr size_t kSkSmV = (size_t)kSkGB * kSkVoPitch * 4;
constexpr size_t kSkSmR = (size_t)4 * kSkRC * 3 * kSkGB * 4;
static size_t skin_smem(int nreg) { return kSkSmA + kSkSmV + kSkSmR + (size_t)(nreg > 0 ? nreg : 1) * kSkJtPitch * 4; }

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem, bool valid) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(sa), "l"(gmem), "r"(sz));
}

__global__ void __launch_bounds__(kSkVT, kSkGB == 8 ? 4 : 3)
k_smpl_skin(const tp_smpl_model m, int body_lo, int body_hi, int bodies_per_cta, const float* __restrict__ vposed, int64_t ldv,
            const float* __restrict__ A, const float* __restrict__ jreg, int nreg, float* __restrict__ verts,
            float* __restrict__ jpart, int ntiles) {
  extern __shared__ __align__(16) unsigned char ssm[];
  float* A_s = reinterpret_cast<float*>(ssm);                                   // [2][GB][24*13]
  float* vout = reinterpret_cast<float*>(ssm + kSkSmA);                         // [GB][388]
  float* red = reinterpret_cast<float*>(ssm + kSkSmA + kSkSmV);                 // [4][RC*3][GB]
  float* Jt = reinterpret_cast<float*>(ssm + kSkSmA + kSkSmV + kSkSmR);         // [nreg][132]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, v0 = tile * kSkVT, v = v0 + tid;
  const int b_lo = body_lo + blockIdx.y * bodies_per_cta, b_hi = min(body_hi, b_lo + bodies_per_cta);
  if (b_lo >= b_hi) return;
  for (int i = tid; i < nreg * kSkVT; i += kSkVT) {
    const int r = i / kSkVT, vv = i - r * kSkVT;
    Jt[r * kSkJtPitch + vv] = jreg[(int64_t)r * m.vp + v0 + vv];
  }
  int sj[4]; float sw[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    sj[k] = k < m.ks ? m.skin_idx[(int64_t)v * m.ks + k] * kSkAPitch : 0;
    sw[k] = k < m.ks ? m.skin_w[(int64_t)v * m.ks + k] : 0.0f;
  }
  const bool vvalid = v < m.n_verts;
  const int nstage = (b_hi - b_lo + kSkGB - 1) / kSkGB;
  auto load_A = [&](int stage) {            // joint transforms of one stage: [GB][24][12] -> row pitch 13, 4-byte async copies
    float* dst = A_s + (size_t)(stage & 1) * kSkGB * kSkABody;
    const int bb0 = b_lo + stage * kSkGB;
    int b = 0, r = tid;                     // r = element index inside a body (24 x 12 = 288)
#pragma unroll 4
    for (int it = 0; it < kSkGB * kJ * 12 / kSkVT; ++it) {
      if (r >= kJ * 12) { r -= kJ * 12; ++b; }
      const int j = (r * 171) >> 11, q = r - j * 12;             // r / 12 for r < 288
      const bool ok = bb0 + b < b_hi;
      cp_async4(dst + b * kSkABody + j * kSkAPitch + q, A + ((int64_t)(ok ? bb0 + b : b_lo) * kJ * 12 + r), ok);
      r += kSkVT;
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  load_A(0);
  if (nstage > 1) load_A(1); else asm volatile("cp.async.commit_group;\n" ::);
  for (int s = 0; s < nstage; ++s) {
    const int bb0 = b_lo + s * kSkGB;
    const int nb = min(kSkGB, b_hi - bb0);
    // blended rest-pose vertices of this stage (L2-resident GEMM output), all loads in flight before A is needed
    float px[kSkGB], py[kSkGB], pz[kSkGB];
#pragma unroll
    for (int b = 0; b < kSkGB; ++b) {
      px[b] = py[b] = pz[b] = 0.0f;
      if (b < nb) {
        const float* src = vposed + (int64_t)(bb0 + b - body_lo) * ldv + (int64_t)v * 3;
        px[b] = __ldcg(src); py[b] = __ldcg(src + 1); pz[b] = __ldcg(src + 2);
      }
    }
    asm volatile("cp.async.wait_group 1;\n" ::);
    __syncthreads();
    const float* As = A_s + (size_t)(s & 1) * kSkGB * kSkABody;
#pragma unroll
    for (int b = 0; b < kSkGB; ++b) {
      float T[12];
#pragma unroll
      for (int q = 0; q < 12; ++q) T[q] = 0.0f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float* Aj = As + b * kSkABody + sj[k];
#pragma unroll
        for (int q = 0; q < 12; ++q) T[q] = fmaf(sw[k], Aj[q], T[q]);
      }
      float* o = vout + b * kSkVoPitch + tid * 3;
      o[0] = vvalid ? (T[0] * px[b] + T[1] * py[b] + T[2] * pz[b] + T[3]) : 0.0f;
      o[1] = vvalid ? (T[4] * px[b] + T[5] * py[b] + T[6] * pz[b] + T[7]) : 0.0f;
      o[2] = vvalid ? (T[8] * px[b] + T[9] * py[b] + T[10] * pz[b] + T[11]) : 0.0f;
    }
    __syncthreads();
    if (s + 2 < nstage) load_A(s + 2); else asm volatile("cp.async.commit_group;\n" ::);
    // coalesced vertex store: 1536 contiguous bytes per body, one warp per body
    if (verts) {
      const int64_t nv3 = (int64_t)m.n_verts * 3, base = (int64_t)v0 * 3;
      for (int b = warp; b < nb; b += 4) {
        const float2* src = reinterpret_cast<const float2*>(vout + b * kSkVoPitch);
        float* dstb = verts + (int64_t)(bb0 + b) * nv3 + base;
#pragma unroll
        for (int k = 0; k < kSkVT * 3 / 64; ++k) {
          const int q = lane + 32 * k;
          const float2 val = src[q];
          if (base + q * 2 + 1 < nv3) *reinterpret_cast<float2*>(dstb + q * 2) = val;
          else if (base + q * 2 < nv3) dstb[q * 2] = val.x;
        }
      }
    }
    // joint regressors: lane = (body, vertex half); a warp's 32 vertices are read once per pass of kSkRC rows
    for (int r0 = 0; r0 < nreg; r0 += kSkRC) {
      constexpr int LPB = 32 / kSkGB;                       // lanes per body: each takes 8 / LPB of the warp's 8 vertex quads
      const int bl = lane % kSkGB, vh = lane / kSkGB;
      float acc[kSkRC][3];
#pragma unroll
      for (int rr = 0; rr < kSkRC; ++rr) acc[rr][0] = acc[rr][1] = acc[rr][2] = 0.0f;
#pragma unroll
      for (int i = 0; i < 8 / LPB; ++i) {
        const int v4 = warp * 8 + vh * (8 / LPB) + i;
        const float4* vb = reinterpret_cast<const float4*>(vout + bl * kSkVoPitch + v4 * 12);
        const float4 p0 = vb[0], p1 = vb[1], p2 = vb[2];                        // x0 y0 z0 x1 | y1 z1 x2 y2 | z2 x3 y3 z3
#pragma unroll
        for (int rr = 0; rr < kSkRC; ++rr) {
          if (r0 + rr < nreg) {
            const float4 w = *reinterpret_cast<const float4*>(Jt + (r0 + rr) * kSkJtPitch + v4 * 4);
            acc[rr][0] = fmaf(w.x, p0.x, acc[rr][0]); acc[rr][1] = fmaf(w.x, p0.y, acc[rr][1]); acc[rr][2] = fmaf(w.x, p0.z, acc[rr][2]);
            acc[rr][0] = fmaf(w.y, p0.w, acc[rr][0]); acc[rr][1] = fmaf(w.y, p1.x, acc[rr][1]); acc[rr][2] = fmaf(w.y, p1.y, acc[rr][2]);
            acc[rr][0] = fmaf(w.z, p1.z, acc[rr][0]); acc[rr][1] = fmaf(w.z, p1.w, acc[rr][1]); acc[rr][2] = fmaf(w.z, p2.x, acc[rr][2]);
            acc[rr][0] = fmaf(w.w, p2.y, acc[rr][0]); acc[rr][1] = fmaf(w.w, p2.z, acc[rr][1]); acc[rr][2] = fmaf(w.w, p2.w, acc[rr][2]);
          }
        }
      }
#pragma unroll
      for (int rr = 0; rr < kSkRC; ++rr)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float t = acc[rr][c];
#pragma unroll
          for (int o = kSkGB; o < 32; o <<= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
          if (vh == 0) red[(warp * kSkRC * 3 + rr * 3 + c) * kSkGB + bl] = t;
        }
      __syncthreads();
      for (int i = tid; i < kSkRC * 3 * kSkGB; i += kSkVT) {
        const int rc = i / kSkGB, b = i - rc * kSkGB, rr = rc / 3, c = rc - rr * 3;
        if (r0 + rr < nreg && b < nb) {
          const float t = (red[(0 * kSkRC * 3 + rc) * kSkGB + b] + red[(1 * kSkRC * 3 + rc) * kSkGB + b]) +
                          (red[(2 * kSkRC * 3 + rc) * kSkGB + b] + red[(3 * kSkRC * 3 + rc) * kSkGB + b]);
          jpart[(((int64_t)(bb0 + b) * ntiles + tile) * nreg + r0 + rr) * 3 + c] = t;
        }
      }
      __syncthreads();
    }
    __syncthreads();
  }
}

#include "smpl_um.inl"
#include "smpl_um2.inl"

// Folded joint regressors (large-batch path): sum_v J[r,v] verts[v] = sum_j ( R_j q[r,j] + t_j g0[r,j] ) with
// q[r,j] = sum_v J[r,v] w[v,j] p_v -- linear in the blend coefficients, so it comes out of one small GEMM over all bodies
// (q = [coef | coef] . [M_hi | M_lo]^T, M = (J (x) w) . Blend folded at pack time, split into two bf16 halves) instead of a pass
// over the 6890 skinned vertices of every body.
struct FoldArgs { const float* q; int64_t ldq; int lo_off; const float* A; const float* g0; };

__global__ void __launch_bounds__(128) k_smpl_finalize(int n, int n_verts, const float* __restrict__ posedJ,
                                                       const float* __restrict__ jpart, int nsplit, int nreg,
                                                       const float* __restrict__ verts, const int32_t* __restrict__ joint_src,
                                                       int nj, const float* __restrict__ cam, int64_t ld_cam,
                                                       float* __restrict__ joints, float* __restrict__ kp2d, const FoldArgs fold) {
  asm volatile("griddepcontrol.wait;\n" ::: "memory");      // no-op unless launched with programmatic serialization
  __shared__ float Jr[kMaxReg * 3];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (fold.q) {
    __shared__ float contrib[kMaxReg * kJ * 3];
    const float* qb = fold.q + (int64_t)b * fold.ldq;
    for (int i = tid; i < nreg * kJ; i += blockDim.x) {          // i = r * 24 + j
      const int j = i % kJ;
      const float* Aj = fold.A + ((int64_t)b * kJ + j) * 12;
      const float q0 = qb[i * 3], q1 = qb[i * 3 + 1], q2 = qb[i * 3 + 2];
      const float g = fold.g0[i];
#pragma unroll
      for (int c = 0; c < 3; ++c)
        contrib[i * 3 + c] = fmaf(Aj[c * 4], q0, fmaf(Aj[c * 4 + 1], q1, fmaf(Aj[c * 4 + 2], q2, Aj[c * 4 + 3] * g)));
    }
    __syncthreads();
    for (int i = tid; i < nreg * 3; i += blockDim.x) {
      const int r = i / 3, c = i - r * 3;
      float s = 0.0f;
      for (int j = 0; j < kJ; ++j) s += contrib[(r * kJ + j) * 3 + c];
      Jr[i] = s;
    }
  } else
  // regressor partials [tile][value]: warp w takes tiles w, w+4, ...; lane = value, so a tile is one contiguous
  // coalesced read and all of a warp's loads are independent (a few L2 round trips in total).  The four warp sums
  // are added in a fixed order -> deterministic.
  if (!fold.q) {
    __shared__ float Jw[4][kMaxReg * 3];
    const int warp = tid >> 5, lane = tid & 31, nval = nreg * 3;
    const float* base = jpart + (int64_t)b * nsplit * nval;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 4
    for (int sp = warp; sp < nsplit; sp += 4) {
      const float* row = base + (int64_t)sp * nval;
      if (lane < nval) a0 += __ldcg(row + lane);
      if (lane + 32 < nval) a1 += __ldcg(row + lane + 32);
      if (lane + 64 < nval) a2 += __ldcg(row + lane + 64);
    }
    if (warp < 4) { Jw[warp][lane] = a0; Jw[warp][lane + 32] = a1; Jw[warp][lane + 64] = a2; }
    __syncthreads();
    for (int i = tid; i < nval; i += blockDim.x) Jr[i] = (Jw[0][i] + Jw[1][i]) + (Jw[2][i] + Jw[3][i]);
  }
  __syncthreads();
  for (int o = tid; o < nj; o += blockDim.x) {
    const int code = joint_src[o];
    float X[3];
    if (code >= 1000) {
      const float* p = verts + ((int64_t)b * n_verts + (code - 1000)) * 3;
      X[0] = p[0]; X[1] = p[1]; X[2] = p[2];
    } else if (code >= 100) {
      X[0] = Jr[(code - 100) * 3]; X[1] = Jr[(code - 100) * 3 + 1]; X[2] = Jr[(code - 100) * 3 + 2];
    } else {
      const float* p = posedJ + ((int64_t)b * kJ + code) * 3;
      X[0] = p[0]; X[1] = p[1]; X[2] = p[2];
    }
    if (joints) {
      float* d = joints + ((int64_t)b * nj + o) * 3;
      d[0] = X[0]; d[1] = X[1]; d[2] = X[2];
    }
    if (kp2d && cam) {
      float c[3] = {cam[(int64_t)b * ld_cam], cam[(int64_t)b * ld_cam + 1], cam[(int64_t)b * ld_cam + 2]}, kp[2];
      project_point_(X, c, kp);
      kp2d[((int64_t)b * nj + o) * 2] = kp[0];
      kp2d[((int64_t)b * nj + o) * 2 + 1] = kp[1];
    }
  }
}

// Large-batch form of the finalize step with a folded regressor: ONE WARP per body, no block-wide barrier.  Lane = joint j: for
// every regressor row r the lane forms R_j q[r,j] + t_j g0[r,j] and the 24 lanes are summed by shuffles (fixed order); then the
// lanes compose the output joint set and project it.
__global__ void __launch_bounds__(128) k_smpl_finalize_fold(int n, int n_verts, const float* __restrict__ posedJ, int nreg,
                                                            const float* __restrict__ verts, const int32_t* __restrict__ joint_src, int nj,
                                                            const float* __restrict__ cam, int64_t ld_cam, float* __restrict__ joints,
                                                            float* __restrict__ kp2d, const FoldArgs fold) {
  asm volatile("griddepcontrol.wait;\n" ::: "memory");
  __shared__ float Jr_all[4][kMaxReg * 3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + warp;
  if (b >= n) return;
  float* Jr = Jr_all[warp];
  const int j = lane < kJ ? lane : kJ - 1;
  float Aj[12];
  {
    const float4* ap = reinterpret_cast<const float4*>(fold.A + ((int64_t)b * kJ + j) * 12);
    const float4 a0 = __ldg(ap), a1 = __ldg(ap + 1), a2 = __ldg(ap + 2);
    Aj[0] = a0.x; Aj[1] = a0.y; Aj[2] = a0.z; Aj[3] = a0.w; Aj[4] = a1.x; Aj[5] = a1.y; Aj[6] = a1.z; Aj[7] = a1.w;
    Aj[8] = a2.x; Aj[9] = a2.y; Aj[10] = a2.z; Aj[11] = a2.w;
  }
  const float* qb = fold.q + (int64_t)b * fold.ldq;
  for (int r = 0; r < nreg; ++r) {
    const int i = r * kJ + j;
    const float q0 = __ldcg(qb + i * 3), q1 = __ldcg(qb + i * 3 + 1), q2 = __ldcg(qb + i * 3 + 2), g = __ldg(fold.g0 + i);
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
    if (lane < kJ) {
      c0 = fmaf(Aj[0], q0, fmaf(Aj[1], q1, fmaf(Aj[2], q2, Aj[3] * g)));
      c1 = fmaf(Aj[4], q0, fmaf(Aj[5], q1, fmaf(Aj[6], q2, Aj[7] * g)));
      c2 = fmaf(Aj[8], q0, fmaf(Aj[9], q1, fmaf(Aj[10], q2, Aj[11] * g)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      c0 += __shfl_xor_sync(0xffffffffu, c0, o); c1 += __shfl_xor_sync(0xffffffffu, c1, o); c2 += __shfl_xor_sync(0xffffffffu, c2, o);
    }
    if (lane == 0) { Jr[r * 3] = c0; Jr[r * 3 + 1] = c1; Jr[r * 3 + 2] = c2; }
  }
  __syncwarp();
  for (int o = lane; o < nj; o += 32) {
    const int code = joint_src[o];
    float X[3];
    if (code >= 1000) {
      const float* p = verts + ((int64_t)b * n_verts + (code - 1000)) * 3;
      X[0] = p[0]; X[1] = p[1]; X[2] = p[2];
    } else if (code >= 100) {
      X[0] = Jr[(code - 100) * 3]; X[1] = Jr[(code - 100) * 3 + 1]; X[2] = Jr[(code - 100) * 3 + 2];
    } else {
      const float* p = posedJ + ((int64_t)b * kJ + code) * 3;
      X[0] = p[0]; X[1] = p[1]; X[2] = p[2];
    }
    if (joints) {
      float* d = joints + ((int64_t)b * nj + o) * 3;
      d[0] = X[0]; d[1] = X[1]; d[2] = X[2];
    }
    if (kp2d && cam) {
      float c[3] = {cam[(int64_t)b * ld_cam], cam[(int64_t)b * ld_cam + 1], cam[(int64_t)b * ld_cam + 2]}, kp[2];
      project_point_(X, c, kp);
      kp2d[((int64_t)b * nj + o) * 2] = kp[0];
      kp2d[((int64_t)b * nj + o) * 2 + 1] = kp[1];
    }
  }
}

static size_t al256(size_t v) { return (v + 255) / 256 * 256; }

struct SmplPlan { int ntiles, nsplit, tiles_per_split, ngroups; size_t off_A, off_J, off_coef, off_part, off_ctc, off_vposed, off_um, off_timg, off_q, total;
                  int tc, tc_tiles, tc_gsplit, tc_gpc, split, chunk, um, n_pad, nq_pad; };

// large-batch path: from this many bodies on, GEMM + skin over L2-resident chunks replaces the fused kernel
static int split_min_bodies() { static const int v = getenv("TP_SMPL_SPLIT_MIN") ? atoi(getenv("TP_SMPL_SPLIT_MIN")) : 1024; return v; }
// fused tcgen05 blend + skinning kernel (k_smpl_lbs_um) for the large-batch path; TP_SMPL_UM=0 falls back to GEMM + k_smpl_skin
// TP_SMPL_UM: 2 (default) = both contractions on tcgen05 (k_smpl_lbs_um2), 1 = blend on tcgen05 + shared-memory skinning gathers
// (k_smpl_lbs_um), 0 = GEMM + k_smpl_skin
static int um_enabled() { static const int v = getenv("TP_SMPL_UM") ? atoi(getenv("TP_SMPL_UM")) : 2; return v; }
static int split_chunk_bodies() { static const int v = getenv("TP_SMPL_CHUNK") ? atoi(getenv("TP_SMPL_CHUNK")) : 1024; return v < 16 ? 16 : v; }

static SmplPlan make_plan(const tp_smpl_model* m, int n, int nreg, int blend_mode) {
  SmplPlan p;
  p.tc = (blend_mode == 1 && m->blend_tc && m->template_pad && m->ks <= 4 && m->vp % kTcVT == 0) ? 1 : 0;
  p.tc_tiles = m->vp / kTcVT;
  {
    // body splits: as many CTAs as there are SM slots in a whole number of waves (1 CTA per SM);
    // among split counts up to 64 pick the one with the best last-wave fill
    const int groups = (n + kTcNB - 1) / kTcNB;
    const int sms = sm_count();
    int best = 1; double best_eff = 0.0;
    for (int sp = 1; sp <= 64 && sp <= groups; ++sp) {
      const int gpc = (groups + sp - 1) / sp, real = (groups + gpc - 1) / gpc;
      const long ctas = (long)p.tc_tiles * real;
      const long waves = (ctas + sms - 1) / sms;
      const double eff = (double)ctas / (double)(waves * sms);
      if (eff > best_eff + 1e-9) { best_eff = eff; best = real; }
    }
    p.tc_gpc = (groups + best - 1) / best;
    p.tc_gsplit = (groups + p.tc_gpc - 1) / p.tc_gpc;
  }
  p.ntiles = m->vp / kVT;
  p.ngroups = (n + kNB - 1) / kNB;
  int want = (2 * sm_count() + p.ngroups - 1) / p.ngroups;   // aim for >= 2 CTAs per SM
  if (want < 1) want = 1;
  if (want > p.ntiles) want = p.ntiles;
  p.tiles_per_split = (p.ntiles + want - 1) / want;
  p.nsplit = (p.ntiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.split = (p.tc && m->blend_km && m->vp % kSkVT == 0 && n >= split_min_bodies()) ? 1 : 0;
  p.um = (p.tc && m->blend_um && m->vp % kUsVT == 0 && nreg <= 16 && n >= split_min_bodies() && um_enabled()) ? 1 : 0;
  if (p.tc && m->blend_um && m->vp % kUsVT == 0 && n >= split_min_bodies() && um_enabled() >= 2 && m->skin_um) p.um = 2;   // any nreg <= kMaxReg
  if (p.um) p.split = 0;
  p.n_pad = p.um ? (n + kUsGB - 1) / kUsGB * kUsGB : n;
  size_t o = 0;
  p.off_A = o; o += al256((size_t)p.n_pad * kJ * 12 * 4);
  p.off_J = o; o += al256((size_t)n * kJ * 3 * 4);
  p.off_coef = o; o += al256((size_t)n * kCoefLd * 4);
  const int part_tiles = p.um ? m->vp / kUsVT : (p.tc ? p.tc_tiles : p.nsplit);
  p.off_part = o; o += al256((size_t)n * part_tiles * (nreg > 0 ? nreg : 1) * 3 * 4);
  p.off_ctc = o; o += p.tc ? al256((size_t)n * kTcK * 2 * 2) : 0;      // x2: the [row | row] form of the fold GEMM
  p.chunk = split_chunk_bodies() < n ? split_chunk_bodies() : n;
  p.off_vposed = o; o += p.split ? al256((size_t)p.chunk * m->vp * 3 * 4) : 0;
  p.off_um = o; o += p.um ? al256((size_t)(p.n_pad / kUsGB) * kUsBBytes) : 0;
  p.off_timg = o; o += p.um == 2 ? al256((size_t)(p.n_pad / kU2SB) * kU2TimgBytes) : 0;
  p.nq_pad = (nreg * kJ * 3 + 15) / 16 * 16;                  // folded-regressor GEMM output: [n][M_hi part | M_lo part] fp32
  p.off_q = o; o += (p.um == 2 && nreg > 0) ? al256((size_t)n * p.nq_pad * 4) : 0;
  p.total = o;
  return p;
}

}  // namespace tp

using namespace tp;

extern "C" size_t tp_smpl_workspace_bytes(const tp_smpl_model* m, int n, int nreg, int blend_mode) {
  if (!m || n <= 0 || m->vp <= 0) return 0;
  return make_plan(m, n, nreg, blend_mode).total;
}

extern "C" int tp_smpl_forward(const tp_smpl_model* m, int n, const float* pose, int64_t ld_pose, int pose_kind,
                               const float* betas, int64_t ld_betas, const float* cam, int64_t ld_cam,
                               const float* jreg, int nreg, const int32_t* joint_src, int nj,
                               float* verts, float* joints, float* kp2d, float* rotmat, float* theta, int blend_mode,
                               void* workspace, size_t workspace_bytes, void* stream) {
  return tp_smpl_forward_ex(m, n, pose, ld_pose, pose_kind, betas, ld_betas, cam, ld_cam, jreg, nreg, nullptr, joint_src, nj, verts, joints,
                            kp2d, rotmat, theta, blend_mode, workspace, workspace_bytes, stream);
}

extern "C" int tp_smpl_forward_ex(const tp_smpl_model* m, int n, const float* pose, int64_t ld_pose, int pose_kind,
                                  const float* betas, int64_t ld_betas, const float* cam, int64_t ld_cam,
                                  const float* jreg, int nreg, const tp_smpl_regfold* fold, const int32_t* joint_src, int nj,
                                  float* verts, float* joints, float* kp2d, float* rotmat, float* theta, int blend_mode,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  TP_CHECK_ARG(m != nullptr, "tp_smpl_forward: null model");
  TP_CHECK_ARG(n >= 0, "tp_smpl_forward: n=%d", n);
  if (n == 0) return TP_OK;
  TP_CHECK_ARG(m->blend && m->j_template && m->j_shapedirs && m->parents && m->skin_idx && m->skin_w,
               "tp_smpl_forward: model has null tables");
  TP_CHECK_ARG(m->vp % kVT == 0 && m->n_verts > 0 && m->n_verts <= m->vp, "tp_smpl_forward: bad n_verts/vp (%d/%d)", m->n_verts, m->vp);
  TP_CHECK_ARG(m->ks >= 1 && m->ks <= kJ, "tp_smpl_forward: ks=%d out of range", m->ks);
  TP_CHECK_ARG(pose && betas, "tp_smpl_forward: null pose/betas");
  TP_CHECK_ARG(pose_kind >= 0 && pose_kind <= 2, "tp_smpl_forward: bad pose_kind %d", pose_kind);
  TP_CHECK_ARG(nreg >= 0 && nreg <= kMaxReg && (nreg == 0 || jreg), "tp_smpl_forward: nreg=%d (max %d) / null jreg", nreg, kMaxReg);
  TP_CHECK_ARG(nj >= 0 && (nj == 0 || joint_src), "tp_smpl_forward: null joint_src");
  TP_CHECK_ARG(aligned16(m->blend), "tp_smpl_forward: blend table must be 16-byte aligned");
  TP_CHECK_ARG(blend_mode == 0 || blend_mode == 1, "tp_smpl_forward: bad blend_mode %d", blend_mode);
  SmplPlan pl = make_plan(m, n, nreg, blend_mode);
  // the folded regressor serves the large-batch tcgen05 path only; it must describe the same nreg rows as jreg
  const bool use_fold = pl.um == 2 && nreg > 0 && fold && fold->m_km && fold->g0 && fold->nreg == nreg && fold->nq_pad == pl.nq_pad;
  TP_CHECK_ARG(!fold || fold->nreg == nreg, "tp_smpl_forward_ex: fold->nreg=%d does not match nreg=%d", fold ? fold->nreg : 0, nreg);
  if (pl.um == 2 && nreg > 16 && !use_fold) {                 // the epilogue regressor handles 16 rows: more need the fold
    return fail(TP_ERR_UNSUPPORTED, "tp_smpl_forward: the large-batch tcgen05 path needs a folded regressor (tp_smpl_regfold) for nreg=%d > 16", nreg);
  }
  TP_CHECK_ARG(workspace && workspace_bytes >= pl.total, "tp_smpl_forward: workspace too small (%zu < %zu)", workspace_bytes, pl.total);
  TP_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "tp_smpl_forward: workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  PrepArgs pa;
  pa.pose = pose; pa.ld_pose = ld_pose; pa.pose_kind = pose_kind;
  pa.betas = betas; pa.ld_betas = ld_betas; pa.cam = cam; pa.ld_cam = ld_cam;
  pa.A = reinterpret_cast<float*>(ws + pl.off_A);
  pa.posedJ = reinterpret_cast<float*>(ws + pl.off_J);
  pa.coef = pl.um ? nullptr : reinterpret_cast<float*>(ws + pl.off_coef);      // the fp32 rows feed the non-tcgen05 vertex kernels only
  pa.rotmat = rotmat; pa.theta = theta;
  pa.coef_tc = (pl.tc && !pl.um) ? reinterpret_cast<__nv_bfloat16*>(ws + pl.off_ctc) : nullptr;
  pa.coef_cat = use_fold ? reinterpret_cast<__nv_bfloat16*>(ws + pl.off_ctc) : nullptr;
  pa.coef_um = pl.um ? ws + pl.off_um : nullptr;
  pa.timg = pl.um == 2 ? ws + pl.off_timg : nullptr;
  pa.n_pad = pl.n_pad; pa.um_rows = pl.um == 2 ? kU2GB : kUsGB;
  float* jpart = reinterpret_cast<float*>(ws + pl.off_part);

  {
    PdlConfig lc(dim3((unsigned)ceil_div(pl.n_pad, 4)), dim3(128), 0, st);
    TP_CUDA(cudaLaunchKernelEx(&lc.cfg, k_smpl_prepare, *m, n, pa));
  }
  TP_LAUNCH_CHECK();
  const bool need_verts_pass = verts != nullptr || nreg > 0;
  if (need_verts_pass && pl.um == 2) {
    U2Params up;
    up.m = *m; up.n = n; up.ngroups = pl.n_pad / kU2GB; up.ntiles = m->vp / kUsVT; up.nreg = use_fold ? 0 : nreg;
    up.coef_img = ws + pl.off_um; up.timg = ws + pl.off_timg; up.jreg = jreg; up.verts = verts; up.jpart = jpart;
    if (use_fold) {        // q = coef . [M_hi ; M_lo]^T (+ the template part as the bias of the hi columns): one tcgen05 GEMM over all bodies
      tp_gemm_seg sg = tp_gemm_seg{};
      sg.m_start = 0; sg.m_rows = n; sg.n_start = 0; sg.n_cols = pl.nq_pad; sg.out = reinterpret_cast<float*>(ws + pl.off_q); sg.ldc = pl.nq_pad;
      sg.bias = fold->q_bias;
      int rc = tp_gemm_bf16_tc(pa.coef_cat, n, fold->m_km, pl.nq_pad, 2 * kTcK, &sg, 1, stream);
      if (rc != TP_OK) return rc;
    }
    TP_CUDA(cudaFuncSetAttribute(k_smpl_lbs_um2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kU2Smem));
    const long long items = (long long)up.ntiles * up.ngroups;
    const int grid = (int)(items < sm_count() ? items : sm_count());
    PdlConfig lc(dim3((unsigned)grid), dim3(kU2Threads), kU2Smem, st);
    TP_CUDA(cudaLaunchKernelEx(&lc.cfg, k_smpl_lbs_um2, up));
    TP_LAUNCH_CHECK();
  } else if (need_verts_pass && pl.um) {
    UsParams up;
    up.m = *m; up.n = n; up.ngroups = pl.n_pad / kUsGB; up.ntiles = m->vp / kUsVT; up.nreg = nreg;
    up.coef_img = ws + pl.off_um; up.A = pa.A; up.jreg = jreg; up.verts = verts; up.jpart = jpart;
    TP_CUDA(cudaFuncSetAttribute(k_smpl_lbs_um, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUsSmem));
    const long long items = (long long)up.ntiles * up.ngroups;
    const int grid = (int)(items < sm_count() ? items : sm_count());
    PdlConfig lc(dim3((unsigned)grid), dim3(kUsThreads), kUsSmem, st);
    TP_CUDA(cudaLaunchKernelEx(&lc.cfg, k_smpl_lbs_um, up));
    TP_LAUNCH_CHECK();
  } else if (need_verts_pass && pl.split) {
    // chunks of bodies: tcgen05 GEMM (blend + template) into an L2-resident scratch, then the skinning kernel
    float* vposed = reinterpret_cast<float*>(ws + pl.off_vposed);
    const int64_t ldv = (int64_t)m->vp * 3;
    const size_t smem = skin_smem(nreg);
    TP_CUDA(cudaFuncSetAttribute(k_smpl_skin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    TP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_smpl_skin, kSkVT, smem));
    const int vt = m->vp / kSkVT, slots = (per_sm > 0 ? per_sm : 1) * sm_count();
    for (int c0 = 0; c0 < n; c0 += pl.chunk) {
      const int cb = n - c0 < pl.chunk ? n - c0 : pl.chunk;
      tp_gemm_seg sg = tp_gemm_seg{};
      sg.m_start = c0; sg.m_rows = cb; sg.n_start = 0; sg.n_cols = m->vp * 3;
      sg.out = vposed; sg.ldc = ldv; sg.bias = m->template_pad;
      int rc = tp_gemm_bf16_tc(pa.coef_tc, n, m->blend_km, m->vp * 3, kTcK, &sg, 1, stream);
      if (rc != TP_OK) return rc;
      int splits = slots / vt;                                 // one wave of 3 CTAs per SM
      const int stages = (cb + kSkGB - 1) / kSkGB;
      if (splits > stages) splits = stages;
      if (splits < 1) splits = 1;
      const int bpc = ((stages + splits - 1) / splits) * kSkGB;
      dim3 grid((unsigned)vt, (unsigned)((cb + bpc - 1) / bpc));
      k_smpl_skin<<<grid, kSkVT, smem, st>>>(*m, c0, c0 + cb, bpc, vposed, ldv, pa.A, jreg, nreg, verts, jpart, vt);
      TP_LAUNCH_CHECK();
    }
  } else if (need_verts_pass && pl.tc) {
    TP_CUDA(cudaFuncSetAttribute(k_smpl_verts_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
    dim3 grid((unsigned)pl.tc_tiles, (unsigned)pl.tc_gsplit);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = kTcSmem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    TP_CUDA(cudaLaunchKernelEx(&cfg, k_smpl_verts_tc, *m, n, (const __nv_bfloat16*)pa.coef_tc, (const float*)pa.A, jreg, nreg, verts,
                               jpart, pl.tc_tiles, pl.tc_gpc));
    TP_LAUNCH_CHECK();
  } else if (need_verts_pass) {
    constexpr size_t smem = (size_t)kSmVertsFloats * sizeof(float);
    TP_CUDA(cudaFuncSetAttribute(k_smpl_verts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)pl.ngroups, (unsigned)pl.nsplit);
    k_smpl_verts<<<grid, kVT, smem, st>>>(*m, n, pa.coef, pa.A, jreg, nreg, verts, jpart, pl.nsplit,
                                          pl.tiles_per_split, pl.ntiles);
    TP_LAUNCH_CHECK();
  }
  if (nj > 0 && (joints || kp2d)) {
    TP_CHECK_ARG(verts != nullptr, "tp_smpl_forward: verts is required when joints are requested (vertex picks read it)");
    {
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3((unsigned)n); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 0; cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
      cfg.attrs = attr; cfg.numAttrs = 1;
      FoldArgs fa;
      fa.q = use_fold ? reinterpret_cast<const float*>(ws + pl.off_q) : nullptr;
      fa.ldq = pl.nq_pad; fa.lo_off = 0; fa.A = pa.A; fa.g0 = use_fold ? fold->g0 : nullptr;
      if (use_fold) {
        cfg.gridDim = dim3((unsigned)ceil_div(n, 4));
        TP_CUDA(cudaLaunchKernelEx(&cfg, k_smpl_finalize_fold, n, (int)m->n_verts, (const float*)pa.posedJ, nreg, (const float*)verts,
                                   joint_src, nj, cam, ld_cam, joints, kp2d, fa));
      } else
      TP_CUDA(cudaLaunchKernelEx(&cfg, k_smpl_finalize, n, (int)m->n_verts, (const float*)pa.posedJ, (const float*)jpart,
                                 (int)((pl.split || pl.um) ? m->vp / kSkVT : (pl.tc ? pl.tc_tiles : pl.nsplit)), nreg, (const float*)verts,
                                 joint_src, nj, cam, ld_cam, joints, kp2d, fa));
    }
    TP_LAUNCH_CHECK();
  }
  return TP_OK;
}
