// K3 -- encoder linear heads + IEF Regressor (reference lib/models/tepose.py:79-85 and
// lib/models/spin.py:250-261).  v1: a fixed sequence of fp32 FFMA GEMM launches on one
// stream (graph-capturable); fc1 is evaluated as  x.W1x^T (once)  +  [pose|shape|cam].W1p^T
// (per iteration) -- algebraically identical to fc1(cat[x, pose, shape, cam]).
#include "skinny.cuh"
#include "ief_cluster.inl"
#include "ief_heads.inl"
#include <stdlib.h>

namespace tp {

__global__ void k_broadcast_rows(const float* __restrict__ src, float* __restrict__ dst, int rows, int width, int src_rows) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * width) dst[i] = (src_rows == 1) ? src[i % width] : src[i];
}

static size_t al256(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace tp

using namespace tp;

#define TP_TRY(x) do { int _rc = (x); if (_rc != TP_OK) return _rc; } while (0)

// One nn.Linear at skinny M.  fp32: W is [N,K] row-major fp32 (FFMA, split-K).  bf16: W is the
// fragment-packed bf16 copy (tp_pack_mma_a_bf16) streamed by the tensor-core skinny kernel.
static const size_t kSplitScratch = 4096 + ((size_t)8 << 20);

static int linear(int precision, const float* A, int64_t lda, const void* W, const float* bias, const float* Cin,
                  int64_t ldcin, float* C, int64_t ldc, int M, int N, int K, float alpha, float beta, int relu_a,
                  void* scratch, void* stream, const void* Alp = nullptr, int64_t ldalp = 0, void* Clp = nullptr,
                  int64_t ldclp = 0) {
  const int sms = tp::sm_count();
  if (precision == TP_PRECISION_BF16) {
    // the tensor-core skinny kernel covers <= 64 rows per launch; more rows go in 64-row chunks
    // (the packed weights stay L2-resident between chunks)
    const int groups = (N + 127) / 128, kb = (K + 31) / 32;
    const unsigned char* alp = reinterpret_cast<const unsigned char*>(Alp);
    unsigned char* clp = reinterpret_cast<unsigned char*>(Clp);
    for (int m0 = 0; m0 < M; m0 += 64) {
      const int mc = M - m0 < 64 ? M - m0 : 64;
      int splits = sms / groups;                     // one wave: groups * splits <= #SMs (the CTAs hold 1 CTA/SM worth of registers)
      if (splits > kb / 4) splits = kb / 4;
      if (splits < 1 || !scratch) splits = 1;
      while (splits > 1 && tp_skinny_bf16_workspace_bytes(mc, N, splits) > kSplitScratch) --splits;
      TP_TRY(tp_skinny_bf16_ex(A + (int64_t)m0 * lda, lda, alp ? alp + (int64_t)m0 * ldalp * 2 : nullptr, ldalp, mc, K, W, N, bias,
                               Cin ? Cin + (int64_t)m0 * ldcin : nullptr, ldcin, C + (int64_t)m0 * ldc, ldc,
                               clp ? clp + (int64_t)m0 * ldclp * 2 : nullptr, ldclp, alpha, beta, relu_a, splits, 0, scratch,
                               kSplitScratch, stream));
    }
    return TP_OK;
  }
  const int bm = M <= 32 ? 32 : 64;
  const int tiles = ((N + 31) / 32) * ((M + bm - 1) / bm);
  int splits = (3 * sms + tiles - 1) / tiles;
  if (splits > K / 128) splits = K / 128;
  if (splits > 16) splits = 16;
  if (splits < 1 || !scratch) splits = 1;
  while (splits > 1 && tp_gemm_f32_splitk_workspace_bytes(M, N, splits) > kSplitScratch) --splits;
  return tp_gemm_f32_splitk(A, lda, reinterpret_cast<const float*>(W), K, bias, Cin, ldcin, C, ldc, M, N, K, alpha, beta,
                            relu_a, splits, scratch, kSplitScratch, stream);
}

extern "C" size_t tp_encoder_heads_workspace_bytes(int B) { (void)B; return kSplitScratch; }

// Eval-mode heads as ONE skinny GEMM: feat = relu([h_fwd | h_rec]) . [0.5 W_fwd | 0.5 W_rec]^T + 0.5 (b_fwd + b_rec)
// (halving is exact in fp32/bf16, so this is the same arithmetic as (lin_fwd + lin_rec) / 2 up to summation order).
// h_cat [B, 3H] holds y[-1] of gru_fwd in columns [0,H) and y_rec[0] in [H,3H); w_cat is the packed [2048, 3H] matrix.
extern "C" int tp_encoder_heads_cat(int precision, const void* w_cat, const float* b_cat, const float* h_cat, int64_t ld_h,
                                    int B, int H, float* feat, void* feat_bf16, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  TP_CHECK_ARG(precision == TP_PRECISION_FP32 || precision == TP_PRECISION_BF16, "tp_encoder_heads_cat: bad precision");
  TP_CHECK_ARG(w_cat && b_cat && h_cat && feat, "tp_encoder_heads_cat: null pointer");
  TP_CHECK_ARG(B >= 1 && H >= 4 && H % 4 == 0, "tp_encoder_heads_cat: bad sizes B=%d H=%d", B, H);
  TP_CHECK_ARG(workspace && workspace_bytes >= kSplitScratch && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
               "tp_encoder_heads_cat: workspace too small / misaligned");
  TP_CUDA(cudaMemsetAsync(workspace, 0, 4096, (cudaStream_t)stream));
  void* sc = (B > 64 && precision != TP_PRECISION_BF16) ? nullptr : workspace;   // fp32 split-K scratch is sized for <= 64 rows
  return linear(precision, h_cat, ld_h, w_cat, b_cat, nullptr, 0, feat, 2048, B, 2048, 3 * H, 1.f, 0.f, 1, sc, stream,
                nullptr, 0, precision == TP_PRECISION_BF16 ? feat_bf16 : nullptr, 2048);
}

extern "C" int tp_encoder_heads(int precision, const void* w_fwd, const float* b_fwd, const void* w_rec,
                                const float* b_rec, const float* h_fwd, int64_t ld_hf, const float* h_rec, int64_t ld_hr,
                                int B, int H, int is_train, float* feat, void* feat_bf16, void* workspace,
                                size_t workspace_bytes, void* stream) {
  TP_CHECK_ARG(precision == TP_PRECISION_FP32 || precision == TP_PRECISION_BF16, "tp_encoder_heads: bad precision");
  const int P = precision;
  TP_CHECK_ARG(w_fwd && b_fwd && w_rec && b_rec && h_fwd && h_rec && feat, "tp_encoder_heads: null pointer");
  TP_CHECK_ARG(B >= 1 && H >= 4 && H % 4 == 0, "tp_encoder_heads: bad sizes B=%d H=%d", B, H);
  TP_CHECK_ARG(workspace && workspace_bytes >= kSplitScratch && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
               "tp_encoder_heads: workspace too small / misaligned");
  TP_CUDA(cudaMemsetAsync(workspace, 0, 4096, (cudaStream_t)stream));
  void* sc = workspace;
  if (B > 64 && P != TP_PRECISION_BF16) sc = nullptr;   // the fp32 split-K scratch is sized for <= 64 rows
  if (!is_train) {
    // (linear_fwd(relu(hF)) + linear_rec(relu(hR))) / 2  -- halving each term first is exact in fp32
    TP_TRY(linear(P, h_fwd, ld_hf, w_fwd, b_fwd, nullptr, 0, feat, 2048, B, 2048, H, 0.5f, 0.f, 1, sc, stream));
    TP_TRY(linear(P, h_rec, ld_hr, w_rec, b_rec, feat, 2048, feat, 2048, B, 2048, 2 * H, 0.5f, 1.f, 1, sc, stream, nullptr, 0,
                  P == TP_PRECISION_BF16 ? feat_bf16 : nullptr, 2048));
  } else {
    // stacked [B,2,2048]: row b holds the fwd features then the rec features
    TP_TRY(linear(P, h_fwd, ld_hf, w_fwd, b_fwd, nullptr, 0, feat, 4096, B, 2048, H, 1.f, 0.f, 1, sc, stream, nullptr, 0,
                  P == TP_PRECISION_BF16 ? feat_bf16 : nullptr, 4096));
    TP_TRY(linear(P, h_rec, ld_hr, w_rec, b_rec, nullptr, 0, feat + 2048, 4096, B, 2048, 2 * H, 1.f, 0.f, 1, sc, stream, nullptr, 0,
                  (P == TP_PRECISION_BF16 && feat_bf16) ? reinterpret_cast<unsigned char*>(feat_bf16) + 2048 * 2 : nullptr, 4096));
  }
  return TP_OK;
}


// ------------------------------------------------------------------------------------------------
// Fused IEF (bf16 mode, <= 32 rows): the whole chain  base = fc1_x(feat);  3 x { fc1_p, fc2, dec }
// as ONE persistent cooperative kernel.  64 CTAs, CTA c owns the 16-row weight tile c of every layer
// (8 warps split K, partials meet in shared memory); activations travel between layers as bf16 rows
// in global memory (L2), layers are separated by a release/acquire grid barrier and the next layer's
// weight fragments are already in flight when a CTA arrives at the barrier.  Every CTA needs the WHOLE
// activation block of a layer (kIefRep > 1 would give groups of CTAs their own copy).
namespace tp {

constexpr int kIefMaxLayers = 16;
constexpr int kIefRep = 1;        // replicas of every inter-layer activation block; >1 spreads the 64 reader CTAs over
                                  // copies (measured: no gain -- the staging phase was bound by constant-cache misses, not L2)
struct IefLayer {
  const __nv_bfloat16* A; int lda; int K; int rep_in;     // rep_in: element stride between input replicas (0 = single copy)
  const uint4* Wp; int N; int wkb;                        // wkb: 32-column blocks per 16-row tile of the packed matrix (tile stride)
  int local_next;                                         // 1: the next layer only consumes what THIS CTA (and thread) wrote -> no grid barrier
  const float* bias; const float* Cin; int ldcin;
  float* C; int ldc; __nv_bfloat16* Clp; int ldclp; int rep_out;   // Clp is written kIefRep times, rep_out elements apart
};
struct IefFusedParams {
  IefLayer layer[kIefMaxLayers];
  int nlayers, M;
  unsigned int* barrier; int barrier_shards;
  int narrow_from, narrow_ctas;   // layers >= narrow_from need only the first narrow_ctas CTAs: the others leave, and the
                                  // barriers of that phase count narrow_ctas arrivals on a second counter (word 240 of the slot)
  int direct_kb;      // layers with at most this many 32-column blocks load their B fragments straight into registers (0 = never)
  const float* feat; const __nv_bfloat16* feat_lp; __nv_bfloat16* feat_cvt;   // feat_cvt: kIefRep bf16 replicas of feat
  const float* init; int init_rows; float* psc; __nv_bfloat16* psc_lp;
  const float* hcat; int64_t ld_h; int KH; __nv_bfloat16* hcat_cvt;   // fused heads: relu(h_cat [M,KH]) -> bf16 in the prologue
  long long* trace;   // debug: [grid][layers][4] clock stamps
};
#define IEF_TRACE(slot) do { if (p.trace && threadIdx.x == 0) p.trace[((size_t)blockIdx.x * kIefMaxLayers + l) * 8 + (slot)] = clock64(); } while (0)

template <int NT>
__global__ void __launch_bounds__(kSkThreads, 1) k_ief_fused(const IefFusedParams p) {
  constexpr int NB = NT * 8, RPW = 17;
  extern __shared__ __align__(16) unsigned char ief_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int ut = blockIdx.x;
  unsigned int epoch = 0;
  // layer table -> shared memory: indexing the kernel-parameter copy with the (dynamic) layer index
  // costs a constant-cache round trip per field
  __shared__ IefLayer sL[kIefMaxLayers];
  for (int i = tid; i < (int)(sizeof(IefLayer) * kIefMaxLayers / 4); i += kSkThreads)
    reinterpret_cast<int*>(sL)[i] = reinterpret_cast<const int*>(p.layer)[i];
  __syncthreads();
  pdl_wait();                    // PDL: the encoder states / features come from the previous kernel
  pdl_launch_dependents();

  // prologue: IEF state <- init (fp32 + bf16 copy), optional feat conversion
  if (p.psc)           // (null: the iterations, and their state, belong to k_ief_cluster)
  for (int i = blockIdx.x * kSkThreads + tid; i < p.M * 160; i += gridDim.x * kSkThreads) {
    const float v = p.init_rows == 1 ? p.init[i % 160] : p.init[i];
    p.psc[i] = v;
#pragma unroll
    for (int r = 0; r < kIefRep; ++r) p.psc_lp[(size_t)r * p.M * 160 + i] = __float2bfloat16_rn(v);
  }
  // fused heads: the encoder states, relu'd, as bf16 rows (lib/models/tepose.py:79-80)
  if (p.hcat)
    for (int i = blockIdx.x * kSkThreads + tid; i < p.M * p.KH; i += gridDim.x * kSkThreads) {
      const int m = i / p.KH, c = i - m * p.KH;
      p.hcat_cvt[i] = __float2bfloat16_rn(fmaxf(__ldcg(p.hcat + (int64_t)m * p.ld_h + c), 0.0f));
    }
  // feat (bf16 from the heads, or fp32) -> kIefRep bf16 replicas
  if (p.feat || p.feat_lp)
  for (int i = blockIdx.x * kSkThreads + tid; i < p.M * 2048; i += gridDim.x * kSkThreads) {
    const __nv_bfloat16 v = p.feat_lp ? p.feat_lp[i] : __float2bfloat16_rn(p.feat[i]);
#pragma unroll
    for (int r = 0; r < kIefRep; ++r) p.feat_cvt[(size_t)r * p.M * 2048 + i] = v;
  }
  grid_barrier_sh(p.barrier, ++epoch, p.barrier_shards);

  uint4 wa[kSkPF], wb[kSkPF];
  auto prefetch = [&](const IefLayer& L) {
    const int nkb = (L.K + 31) / 32;
    const int b_lo = (warp * nkb) / 8, nb = ((warp + 1) * nkb) / 8 - b_lo;
    const uint4* wp = L.Wp + ((int64_t)ut * L.wkb + b_lo) * 64 + lane;
#pragma unroll
    for (int q = 0; q < kSkPF; ++q)
      if (q < nb) { wa[q] = ldg_stream16(wp + (int64_t)q * 64); wb[q] = ldg_stream16(wp + (int64_t)q * 64 + 32); }
  };
  if (ut < (sL[0].N + 15) / 16) prefetch(sL[0]);

  unsigned int epoch2 = 0;
  auto barrier_narrow = [&]() {          // grid_barrier() among the first narrow_ctas CTAs
    __syncthreads();
    if (tid == 0) {
      unsigned int* ctr = p.barrier + 32 * (kBarrierShards - 1) + 16;      // word 240 of the 1 KB slot: not one of grid_barrier_sh's shard counters (base + 32 k)
      const unsigned int target = ++epoch2 * (unsigned int)p.narrow_ctas;
      unsigned int seen;
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(ctr) : "memory");
      do {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];\n" : "=r"(seen) : "l"(ctr) : "memory");
      } while (seen < target);
#ifdef TP_BARRIER_ACQUIRE_FENCE
      asm volatile("fence.acq_rel.gpu;\n" ::: "memory");
#endif
    }
    __syncthreads();
  };
  for (int l = 0; l < p.nlayers; ++l) {
    if (l == p.narrow_from && (int)blockIdx.x >= p.narrow_ctas) return;      // no work left for this CTA
    const IefLayer L = sL[l];
    IEF_TRACE(0);
    if (ut < (L.N + 15) / 16) {
      const int nkb = (L.K + 31) / 32;
      const int pitch = ((nkb * 32 + 63) / 64) * 64 + 32;
      __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(ief_smem);
      float* red = reinterpret_cast<float*>(ief_smem + (size_t)NB * pitch * 2);
      const int b_lo = (warp * nkb) / 8, nb = ((warp + 1) * nkb) / 8 - b_lo;
      // epilogue operands do not depend on this layer's matmul: fetch them first
      constexpr int E = (NB * 16 + kSkThreads - 1) / kSkThreads;
      float cin[E], bia[E];
#pragma unroll
      for (int i = 0; i < E; ++i) {
        const int idx = tid + i * kSkThreads, m = idx >> 4, r = idx & 15, nn = ut * 16 + r;
        const bool ok = idx < NB * 16 && m < p.M && nn < L.N;
        cin[i] = (ok && L.Cin) ? __ldcg(L.Cin + (int64_t)m * L.ldcin + nn) : 0.0f;
        bia[i] = (ok && L.bias) ? __ldg(L.bias + nn) : 0.0f;
      }
      // activations of this layer (bf16 rows written by the previous layer / the prologue): own replica
      const __nv_bfloat16* Ain = L.A + (size_t)(blockIdx.x % kIefRep) * L.rep_in;
      float acc[NT][4];
#pragma unroll
      for (int n = 0; n < NT; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.0f;
      if (nkb <= p.direct_kb) {
        // a warp's B fragments are <= 8 k-blocks x NT row groups of 16 bytes -- load them straight from L2 into
        // registers (one round trip per 4 k-blocks), no shared-memory staging and no CTA-wide sync before the MMAs
        // (two rounds of <= 4 k-blocks per warp, so K = 2048 layers take this path too)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h * 4 < nb) {
            uint4 bv[4][NT];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
              for (int n = 0; n < NT; ++n) {
                const int r = n * 8 + g;
                bv[q][n] = make_uint4(0u, 0u, 0u, 0u);
                if (h * 4 + q < nb && r < p.M)
                  bv[q][n] = __ldcg(reinterpret_cast<const uint4*>(Ain + (int64_t)r * L.lda + (b_lo + h * 4 + q) * 32 + 8 * t));
              }
            if (h == 0) IEF_TRACE(1);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (h * 4 + q < nb) {
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                  mma16816(acc[n], wa[h * 4 + q], bv[q][n].x, bv[q][n].y);
                  mma16816(acc[n], wb[h * 4 + q], bv[q][n].z, bv[q][n].w);
                }
              }
            }
          }
        }
      } else {
      // every thread keeps ALL its 16-byte loads of a half block in flight before the first store:
      // this phase is a pure L2 round trip, so bytes in flight decide its duration
      const int c8n = nkb * 4;                                  // 16-byte groups per row
      const int per_row = (c8n + 31) / 32;                      // groups per lane per row (<= 8 for K <= 2048)
      for (int r0 = warp; r0 < NB; r0 += 2 * (kSkThreads / 32)) {
        uint4 v[2][8];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int r = r0 + rr * (kSkThreads / 32);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int col = (lane + 32 * k) * 8;
            v[rr][k] = make_uint4(0u, 0u, 0u, 0u);
            if (k < per_row && r < NB && r < p.M && col < L.K) v[rr][k] = __ldcg(reinterpret_cast<const uint4*>(Ain + (int64_t)r * L.lda + col));   // L2-coherent (barrier contract)
          }
        }
        if (r0 == warp) IEF_TRACE(4);          // all loads of the first half issued
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int r = r0 + rr * (kSkThreads / 32);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int col = (lane + 32 * k) * 8;
            if (k < per_row && r < NB && col < nkb * 32) *reinterpret_cast<uint4*>(As + (size_t)r * pitch + col) = v[rr][k];
          }
        }
        if (r0 == warp) IEF_TRACE(5);          // first half landed and stored
      }
      IEF_TRACE(6);                              // this warp done staging
      __syncthreads();
      IEF_TRACE(1);
#pragma unroll
      for (int q = 0; q < kSkPF; ++q) {
        if (q < nb) {
#pragma unroll
          for (int n = 0; n < NT; ++n) {
            const uint4 bv = *reinterpret_cast<const uint4*>(As + (size_t)(n * 8 + g) * pitch + (b_lo + q) * 32 + 8 * t);
            mma16816(acc[n], wa[q], bv.x, bv.y);
            mma16816(acc[n], wb[q], bv.z, bv.w);
          }
        }
      }
      }
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        float* r0 = red + ((size_t)warp * NB + n * 8 + 2 * t) * RPW + g;
        r0[0] = acc[n][0]; r0[RPW] = acc[n][1]; r0[8] = acc[n][2]; r0[RPW + 8] = acc[n][3];
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < E; ++i) {
        const int idx = tid + i * kSkThreads, m = idx >> 4, r = idx & 15, nn = ut * 16 + r;
        if (idx < NB * 16 && m < p.M && nn < L.N) {
          float sum = 0.0f;
#pragma unroll
          for (int w = 0; w < 8; ++w) sum += red[((size_t)w * NB + m) * RPW + r];
          const float v = sum + bia[i] + cin[i];
          if (L.C) L.C[(int64_t)m * L.ldc + nn] = v;
          if (L.Clp) {
            const __nv_bfloat16 vb = __float2bfloat16_rn(v);
#pragma unroll
            for (int rr = 0; rr < kIefRep; ++rr) L.Clp[(size_t)rr * L.rep_out + (int64_t)m * L.ldclp + nn] = vb;
          }
        }
      }
      __syncthreads();     // As / red are reused by the next layer
    }
    IEF_TRACE(2);
    if (l + 1 < p.nlayers) {
      if (ut < (sL[l + 1].N + 15) / 16) prefetch(sL[l + 1]);   // weights do not depend on the barrier
      if (!L.local_next) {
        if (l + 1 > p.narrow_from) barrier_narrow();      // producer and consumers of layer l's output are all narrow-phase CTAs
        else grid_barrier_sh(p.barrier, ++epoch, p.barrier_shards);
      }
    }
    IEF_TRACE(3);
  }
}

}  // namespace tp

struct HeadsArgs { const void* w_cat; const float* b_cat; const float* h_cat; int64_t ld_h; int H; };

// The iterations of the IEF loop as one 16-CTA cluster with the weights on chip (ief_cluster.inl).  Needs a GPC with 16 free SMs
// and 223 KB of shared memory per CTA; TP_IEF_NO_CLUSTER=1 keeps the iterations inside the grid-barrier kernel.
static int g_ief_cluster = 1;       // tp_set_ief_cluster()
extern "C" int tp_set_ief_cluster(int enable) { const int was = g_ief_cluster; g_ief_cluster = enable ? 1 : 0; return was; }

static bool ief_cluster_available() {
  if (!g_ief_cluster) return false;
  static const int ok = [] {
    if (getenv("TP_IEF_NO_CLUSTER")) return 0;
    if (cudaFuncSetAttribute(tp::k_ief_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
        cudaFuncSetAttribute(tp::k_ief_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tp::kClSmemBytes) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    cudaLaunchConfig_t cfg = cudaLaunchConfig_t{};
    cudaLaunchAttribute at[1];
    cfg.gridDim = dim3(tp::kClCtas); cfg.blockDim = dim3(tp::kClThreads); cfg.dynamicSmemBytes = tp::kClSmemBytes;
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = tp::kClCtas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, tp::k_ief_cluster, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n >= 1 ? 1 : 0;
  }();
  return ok != 0;
}

// heads + fc1's feature part as the split-K kernel (ief_heads.inl); the caller has checked ief_cluster_available().
// scratch: >= 4 MB (two partial-sum areas of kHbPartBytes)
static int heads_base_launch(const tp_ief_weights* w, const HeadsArgs* hd, const float* feat, const void* feat_bf16, int N,
                             __nv_bfloat16* feat_rep, float* base, unsigned char* scratch, unsigned int* barrier, int barrier_shards,
                             cudaStream_t st) {
  tp::HeadsBaseParams q;
  memset(&q, 0, sizeof(q));
  int pitch = 256 + 32;
  if (hd) {
    q.hcat = hd->h_cat; q.ld_h = hd->ld_h; q.KH = 3 * hd->H;
    q.w_cat = reinterpret_cast<const uint4*>(hd->w_cat); q.b_cat = hd->b_cat;
    if (q.KH / 8 + 32 > pitch) pitch = q.KH / 8 + 32;
  } else {
    q.feat = feat; q.feat_lp = reinterpret_cast<const __nv_bfloat16*>(feat_bf16);
  }
  q.feat_rep = feat_rep; q.w1x = reinterpret_cast<const uint4*>(w->w1x); q.b1 = w->b1; q.base = base; q.M = N;
  q.part1 = reinterpret_cast<float*>(scratch); q.part2 = reinterpret_cast<float*>(scratch + tp::kHbPartBytes);
  q.barrier = barrier; q.barrier_shards = barrier_shards; q.trace = tp::trace_ptr();
  const size_t smem = tp::kHbOffAs + (size_t)32 * pitch * 2;
  TP_CUDA(cudaFuncSetAttribute(tp::k_heads_base, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tp::PdlConfig lc(dim3(tp::kHbGrid), dim3(tp::kHbThreads), smem, st, /*cooperative=*/true);   // the CTAs of a group wait for each other
  TP_CUDA(cudaLaunchKernelEx(&lc.cfg, tp::k_heads_base, q));
  tp::count_launch();
  return TP_OK;
}

static int ief_cluster_launch(const tp_ief_weights* w, const float* base, int N, const float* init, int init_rows, int n_iter,
                              float* psc, cudaStream_t st) {
  tp::IefClParams q;
  memset(&q, 0, sizeof(q));
  q.base = base; q.w1p = reinterpret_cast<const uint4*>(w->w1p); q.w2 = reinterpret_cast<const uint4*>(w->w2);
  q.wdec = reinterpret_cast<const uint4*>(w->wdec); q.b2 = w->b2; q.bdec = w->bdec;
  q.init = init; q.init_rows = init_rows; q.psc = psc; q.M = N; q.n_iter = n_iter;
  static const int rpc_env = getenv("TP_IEF_CLUSTER_ROWS") ? atoi(getenv("TP_IEF_CLUSTER_ROWS")) : 8;
  q.rows_per_cluster = (rpc_env == 8 || rpc_env == 16 || rpc_env == 32) ? rpc_env : 8;
  const int nclusters = (N + q.rows_per_cluster - 1) / q.rows_per_cluster;
  q.trace = tp::trace_ptr();
  TP_CUDA(cudaFuncSetAttribute(tp::k_ief_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  TP_CUDA(cudaFuncSetAttribute(tp::k_ief_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tp::kClSmemBytes));
  cudaLaunchConfig_t cfg = cudaLaunchConfig_t{};
  cudaLaunchAttribute at[2];
  cfg.gridDim = dim3(tp::kClCtas * nclusters); cfg.blockDim = dim3(tp::kClThreads); cfg.dynamicSmemBytes = tp::kClSmemBytes; cfg.stream = st;
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = tp::kClCtas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = tp::pdl_enabled() ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 2;
  TP_CUDA(cudaLaunchKernelEx(&cfg, tp::k_ief_cluster, q));
  tp::count_launch();
  return TP_OK;
}

static int ief_fused(const tp_ief_weights* w, const float* feat, const void* feat_bf16, int N, const float* init,
                     int init_rows, int n_iter, float* psc, unsigned char* ws, cudaStream_t st, const HeadsArgs* hd = nullptr,
                     void* barrier = nullptr) {
  // workspace (see tp_ief_workspace_bytes): [base fp32 | ...per-layer buffers of the unfused path... | scratch];
  // the fused kernel keeps its barrier counter and all replica buffers in the 8 MB scratch region
  const size_t slab = al256((size_t)N * 1024 * sizeof(float));
  const size_t slab_lp = al256((size_t)N * 1024 * 2), slab_p = al256((size_t)N * 160 * 2);
  float* base = reinterpret_cast<float*>(ws);
  unsigned char* sc = ws + 3 * slab + 2 * slab_lp + slab_p;
  const size_t n1024 = (size_t)N * 1024, n160 = (size_t)N * 160, n2048 = (size_t)N * 2048;
  __nv_bfloat16* rep = reinterpret_cast<__nv_bfloat16*>(sc + 4096);
  __nv_bfloat16* u1_lp = rep;                                   // [kIefRep][N,1024]
  __nv_bfloat16* u2_lp = u1_lp + kIefRep * n1024;               // [kIefRep][N,1024]
  __nv_bfloat16* psc_lp = u2_lp + kIefRep * n1024;              // [kIefRep][N,160]
  __nv_bfloat16* feat_rep = psc_lp + kIefRep * n160;            // [kIefRep][N,2048]
  __nv_bfloat16* hcat_cvt = feat_rep + kIefRep * n2048;         // [N,3H] relu'd encoder states (fused heads)
  float* featf = reinterpret_cast<float*>(ws + slab);           // [N,2048] fp32 head accumulator (the u1 | u2 slabs of the unfused path)
  IefFusedParams p;
  memset(&p, 0, sizeof(p));
  p.M = N; p.barrier = reinterpret_cast<unsigned int*>(barrier ? barrier : sc);
  static const int shards_env = getenv("TP_BARRIER_SHARDS") ? atoi(getenv("TP_BARRIER_SHARDS")) : 1;
  p.barrier_shards = barrier ? (shards_env >= 1 && shards_env <= tp::kBarrierShards ? shards_env : 1) : 1;
  p.feat = feat; p.feat_lp = reinterpret_cast<const __nv_bfloat16*>(feat_bf16); p.feat_cvt = feat_rep;
  const bool cluster_tail = n_iter <= tp::kClMaxIter && ief_cluster_available();
  p.init = init; p.init_rows = init_rows; p.psc = cluster_tail ? nullptr : psc; p.psc_lp = psc_lp;
  p.trace = tp::trace_ptr();
  static const bool no_direct = getenv("TP_IEF_NO_DIRECT") != nullptr;
  static const int direct_env = getenv("TP_IEF_DIRECT_KB") ? atoi(getenv("TP_IEF_DIRECT_KB")) : 32;
  p.direct_kb = no_direct ? 0 : direct_env;     // 32: layers with K <= 1024 load their B fragments straight from L2 (64 = K <= 2048 too: measured, no difference)
  int n = 0;
  auto add = [&](const __nv_bfloat16* A, int lda, int K, size_t rep_in, const void* Wp, int Nn, const float* bias,
                 const float* Cin, int ldcin, float* C, int ldc, __nv_bfloat16* Clp, int ldclp, size_t rep_out) {
    IefLayer& L = p.layer[n++];
    L.A = A; L.lda = lda; L.K = K; L.rep_in = (int)rep_in; L.Wp = reinterpret_cast<const uint4*>(Wp); L.N = Nn; L.wkb = (K + 31) / 32;
    L.bias = bias; L.Cin = Cin; L.ldcin = ldcin; L.C = C; L.ldc = ldc; L.Clp = Clp; L.ldclp = ldclp; L.rep_out = (int)rep_out;
  };
  int grid = 64;
  p.narrow_from = 0; p.narrow_ctas = 64;
  if (hd) {
    // eval heads as leading layers: feat = relu(h_cat) . w_cat^T + b_cat, the K = 3H reduction cut into slices of <= 2048
    // columns that accumulate through the fp32 buffer (one grid barrier each); 2048 output rows = 128 weight tiles
    const int KH = 3 * hd->H, wkb = KH / 32;
    p.hcat = hd->h_cat; p.ld_h = hd->ld_h; p.KH = KH; p.hcat_cvt = hcat_cvt;
    p.feat = nullptr; p.feat_lp = nullptr;
    const int nsl = (KH + 2047) / 2048;
    for (int sl = 0; sl < nsl; ++sl) {
      const int k0 = sl * 2048, kk = KH - k0 < 2048 ? KH - k0 : 2048;
      add(hcat_cvt + k0, KH, kk, 0, reinterpret_cast<const uint4*>(hd->w_cat) + (size_t)(k0 / 32) * 64, 2048, sl == 0 ? hd->b_cat : nullptr,
          sl > 0 ? featf : nullptr, 2048, sl + 1 < nsl ? featf : nullptr, 2048, sl + 1 == nsl ? feat_rep : nullptr, 2048, n2048);
      p.layer[n - 1].wkb = wkb;
      p.layer[n - 1].local_next = sl + 1 < nsl ? 1 : 0;     // K slices of one output tile accumulate inside the owning CTA
    }
    grid = 128;
    static const bool no_narrow = getenv("TP_IEF_NO_NARROW") != nullptr;
    p.narrow_from = n;                     // the IEF layers proper (1024 or 160 output rows) live on CTAs 0..63
    if (no_narrow) p.narrow_ctas = 128;
  }
  add(feat_rep, 2048, 2048, n2048, w->w1x, 1024, w->b1, nullptr, 0, base, 1024, nullptr, 0, 0);
  for (int it = 0; it < (cluster_tail ? 0 : n_iter); ++it) {
    add(psc_lp, 160, 160, n160, w->w1p, 1024, nullptr, base, 1024, nullptr, 0, u1_lp, 1024, n1024);
    add(u1_lp, 1024, 1024, n1024, w->w2, 1024, w->b2, nullptr, 0, nullptr, 0, u2_lp, 1024, n1024);
    add(u2_lp, 1024, 1024, n1024, w->wdec, 160, w->bdec, psc, 160, psc, 160, psc_lp, 160, n160);
  }
  p.nlayers = n;
  if (!barrier) TP_CUDA(cudaMemsetAsync(sc, 0, 1024, st));
  static const bool no_splitk = getenv("TP_IEF_NO_SPLITK") != nullptr;
  if (cluster_tail && !no_splitk && tp::sm_count() >= tp::kHbGrid && (!hd || ((3 * hd->H) % 1024 == 0 && 3 * hd->H <= 8192))) {
    // (scratch: the replica buffers end below 4 MB -- see the check in tp_heads_ief_forward -- the partial sums take [4 MB, 8 MB))
    TP_TRY(heads_base_launch(w, hd, feat, feat_bf16, N, feat_rep, base, sc + ((size_t)4 << 20), p.barrier, p.barrier_shards, st));
    return ief_cluster_launch(w, base, N, init, init_rows, n_iter, psc, st);
  }
  const int nb = N <= 8 ? 8 : 32;
  const size_t smem = (size_t)nb * (2048 + 32) * 2 + (size_t)8 * nb * 17 * 4;
  if (nb == 8) {
    TP_CUDA(cudaFuncSetAttribute(tp::k_ief_fused<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tp::PdlConfig lc(dim3(grid), dim3(tp::kSkThreads), smem, st, /*cooperative=*/true);
    TP_CUDA(cudaLaunchKernelEx(&lc.cfg, tp::k_ief_fused<1>, p));
  } else {
    TP_CUDA(cudaFuncSetAttribute(tp::k_ief_fused<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tp::PdlConfig lc(dim3(grid), dim3(tp::kSkThreads), smem, st, /*cooperative=*/true);
    TP_CUDA(cudaLaunchKernelEx(&lc.cfg, tp::k_ief_fused<4>, p));
  }
  tp::count_launch();
  if (cluster_tail) return ief_cluster_launch(w, base, N, init, init_rows, n_iter, psc, st);
  return TP_OK;
}

extern "C" size_t tp_ief_workspace_bytes(int n_rows) {
  const size_t n = (size_t)(n_rows > 0 ? n_rows : 0);
  return 3 * al256(n * 1024 * sizeof(float)) + 2 * al256(n * 1024 * 2) + al256(n * 160 * 2) + kSplitScratch;
}

extern "C" int tp_ief_forward(int precision, const tp_ief_weights* w, const float* feat, const void* feat_bf16, int n_rows, const float* init,
                              int init_rows, int n_iter, float* psc, void* workspace, size_t workspace_bytes, void* stream) {
  TP_CHECK_ARG(precision == TP_PRECISION_FP32 || precision == TP_PRECISION_BF16, "tp_ief_forward: bad precision");
  const int P = precision;
  TP_CHECK_ARG(w && feat && init && psc, "tp_ief_forward: null pointer");
  TP_CHECK_ARG(n_rows >= 1 && n_iter >= 0, "tp_ief_forward: bad sizes n_rows=%d n_iter=%d", n_rows, n_iter);
  TP_CHECK_ARG(init_rows == 1 || init_rows == n_rows, "tp_ief_forward: init_rows must be 1 or n_rows");
  TP_CHECK_ARG(w->w1x && w->b1 && w->w1p && w->w2 && w->b2 && w->wdec && w->bdec, "tp_ief_forward: null weight");
  TP_CHECK_ARG(workspace && workspace_bytes >= tp_ief_workspace_bytes(n_rows), "tp_ief_forward: workspace too small");
  TP_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "tp_ief_forward: workspace must be 256-byte aligned");
  const int N = n_rows;
  static const bool unfused = getenv("TP_IEF_UNFUSED") != nullptr;
  if (P == TP_PRECISION_BF16 && N <= 32 && n_iter >= 0 && 1 + 3 * n_iter <= kIefMaxLayers && !unfused && sm_count() >= 64)
    return ief_fused(w, feat, feat_bf16, N, init, init_rows, n_iter, psc, reinterpret_cast<unsigned char*>(workspace),
                     (cudaStream_t)stream);
  const size_t slab = al256((size_t)N * 1024 * sizeof(float));
  float* base = reinterpret_cast<float*>(workspace);
  float* u1 = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + slab);
  float* u2 = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + 2 * slab);
  // bf16 copies of the activations that feed the next layer (written by the producing epilogue)
  const size_t slab_lp = al256((size_t)N * 1024 * 2), slab_p = al256((size_t)N * 160 * 2);
  unsigned char* wsb = reinterpret_cast<unsigned char*>(workspace) + 3 * slab;
  void* u1_lp = P == TP_PRECISION_BF16 ? wsb : nullptr;
  void* u2_lp = P == TP_PRECISION_BF16 ? wsb + slab_lp : nullptr;
  void* psc_lp = P == TP_PRECISION_BF16 ? wsb + 2 * slab_lp : nullptr;
  void* sc = wsb + 2 * slab_lp + slab_p;
  TP_CUDA(cudaMemsetAsync(sc, 0, 4096, (cudaStream_t)stream));
  if (N > 64 && P != TP_PRECISION_BF16) sc = nullptr;   // fp32 split-K scratch is sized for <= 64 rows
  TP_TRY(linear(P, feat, 2048, w->w1x, w->b1, nullptr, 0, base, 1024, N, 1024, 2048, 1.f, 0.f, 0, sc, stream,
                P == TP_PRECISION_BF16 ? feat_bf16 : nullptr, 2048));
  k_broadcast_rows<<<(unsigned)ceil_div((int64_t)N * 160, 256), 256, 0, (cudaStream_t)stream>>>(init, psc, N, 160, init_rows);
  TP_LAUNCH_CHECK();
  for (int it = 0; it < n_iter; ++it) {
    // the first iteration reads the fp32 init; later ones the bf16 copy the previous dec layer left
    TP_TRY(linear(P, psc, 160, w->w1p, nullptr, base, 1024, u1, 1024, N, 1024, 160, 1.f, 1.f, 0, sc, stream,
                  it > 0 ? psc_lp : nullptr, 160, u1_lp, 1024));
    TP_TRY(linear(P, u1, 1024, w->w2, w->b2, nullptr, 0, u2, 1024, N, 1024, 1024, 1.f, 0.f, 0, sc, stream, u1_lp, 1024,
                  u2_lp, 1024));
    TP_TRY(linear(P, u2, 1024, w->wdec, w->bdec, psc, 160, psc, 160, N, 160, 1024, 1.f, 1.f, 0, sc, stream, u2_lp, 1024,
                  psc_lp, 160));
  }
  return TP_OK;
}

// Eval heads + IEF in ONE persistent kernel (bf16 mode, <= 32 rows): the heads GEMM becomes the leading layers of
// k_ief_fused, so the [N,2048] feature never leaves the kernel as a separate launch.
extern "C" int tp_heads_ief_forward(const void* w_cat, const float* b_cat, const float* h_cat, int64_t ld_h, int H,
                                    const tp_ief_weights* w, int n_rows, const float* init, int init_rows, int n_iter, float* psc,
                                    void* workspace, size_t workspace_bytes, void* barrier, void* stream) {
  TP_CHECK_ARG(w_cat && b_cat && h_cat && w && init && psc, "tp_heads_ief_forward: null pointer");
  TP_CHECK_ARG(n_rows >= 1 && n_rows <= 32, "tp_heads_ief_forward: n_rows=%d (1..32; use tp_encoder_heads_cat + tp_ief_forward beyond)", n_rows);
  TP_CHECK_ARG(H >= 32 && H % 32 == 0, "tp_heads_ief_forward: H=%d must be a multiple of 32", H);
  TP_CHECK_ARG(n_iter >= 0 && (3 * H + 2047) / 2048 + 1 + 3 * n_iter <= kIefMaxLayers, "tp_heads_ief_forward: too many layers (H=%d, n_iter=%d)", H, n_iter);
  TP_CHECK_ARG(init_rows == 1 || init_rows == n_rows, "tp_heads_ief_forward: init_rows must be 1 or n_rows");
  TP_CHECK_ARG(w->w1x && w->b1 && w->w1p && w->w2 && w->b2 && w->wdec && w->bdec, "tp_heads_ief_forward: null weight");
  TP_CHECK_ARG(workspace && workspace_bytes >= tp_ief_workspace_bytes(n_rows), "tp_heads_ief_forward: workspace too small");
  TP_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "tp_heads_ief_forward: workspace must be 256-byte aligned");
  TP_CHECK_ARG((size_t)4096 + (size_t)n_rows * (2 * 1024 + 160 + 2048 + 3 * (size_t)H) * 2 <= ((size_t)4 << 20), "tp_heads_ief_forward: H too large for the scratch region");
  if (sm_count() < 128) return fail(TP_ERR_UNSUPPORTED, "tp_heads_ief_forward needs 128 co-resident CTAs");
  HeadsArgs hd{w_cat, b_cat, h_cat, ld_h, H};
  return ief_fused(w, nullptr, nullptr, n_rows, init, init_rows, n_iter, psc, reinterpret_cast<unsigned char*>(workspace),
                   (cudaStream_t)stream, &hd, barrier);
}
