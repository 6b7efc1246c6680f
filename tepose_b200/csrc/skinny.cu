// Weight-streaming "skinny" GEMM for M <= 64 activation rows (nn.Linear at batch 32/64, the
// live-stream input projection at batch 1):
//     C[M,N] = alpha * ( act(A)[M,K] . W[N,K]^T + bias ) + beta * Cin
// At these sizes the op is bound by streaming W once from HBM, not by math, so:
//   * W is pre-packed (tp_pack_mma_a_bf16) in mma.sync m16n8k16 A-fragment order: every lane
//     fetches its 4 fragment registers with ONE 16-byte load, a warp reads 512 contiguous bytes;
//   * a CTA owns 128 weight rows (8 warps x one 16-row tile) and one K slice (split-K over
//     blockIdx.y keeps all SMs streaming); each lane keeps up to 16 fragment loads in flight;
//   * the activation slice is converted to bf16 once per CTA into shared memory (relu fused);
//   * split-K partials are reduced in split order by the last CTA to arrive (deterministic).
// Tensor throughput is irrelevant here (M <= 64), which is why this path uses mma.sync rather
// than tcgen05: no TMEM round trip, accumulators stay in registers for the fused epilogue.
#include "skinny.cuh"
#include <stdlib.h>
#include <string.h>

namespace tp {

// W [rows, cols] fp32 (row stride ld) -> bf16 fragments  dst[ut = ceil(rows/16)][kb = ceil(cols/32)][q=2][lane=32][word=4].
// Lane (g = lane/4, t = lane%4), word w = 2*e2 + hi holds
//   W[ut*16 + hi*8 + g][kb*32 + 8t + 4q + 2*e2 + {0,1}]   (zero beyond rows / cols).
// The K order inside a 32-column block is permuted so that the matching B fragments of both
// k-subtiles are one plain 16-byte read of row-major activations at column kb*32 + 8t.
__global__ void k_pack_mma_a(const float* __restrict__ w, int64_t ld, int rows, int cols, uint32_t* __restrict__ dst) {
  const int KB = (cols + 31) / 32, UT = (rows + 15) / 16;
  const int64_t words = (int64_t)UT * KB * 256;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < words; i += (int64_t)gridDim.x * blockDim.x) {
    const int wd = (int)(i & 3), lane = (int)((i >> 2) & 31), q = (int)((i >> 7) & 1);
    const int64_t rest = i >> 8;
    const int kb = (int)(rest % KB), ut = (int)(rest / KB);
    const int e2 = wd >> 1, hi = wd & 1, g = lane >> 2, t = lane & 3;
    const int row = ut * 16 + hi * 8 + g;
    const int col = kb * 32 + 8 * t + 4 * q + 2 * e2;
    float v0 = 0.f, v1 = 0.f;
    if (row < rows) {
      if (col < cols) v0 = w[(int64_t)row * ld + col];
      if (col + 1 < cols) v1 = w[(int64_t)row * ld + col + 1];
    }
    __nv_bfloat162 v = __floats2bfloat162_rn(v0, v1);
    dst[i] = *reinterpret_cast<uint32_t*>(&v);
  }
}

template <int NT>
__global__ void __launch_bounds__(kSkThreads, 1) k_skinny_bf16(const SkArgs a) {
  constexpr int NB = NT * 8;
  extern __shared__ __align__(16) unsigned char sk_smem[];
  __shared__ int s_last;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int kb_lo = blockIdx.y * a.kb_per_split;
  const int kb_hi = min(a.kb_total, kb_lo + a.kb_per_split);
  const int nkb = max(0, kb_hi - kb_lo);
  const int pitch = ((nkb * 32 + 63) / 64) * 64 + 32;          // bf16 elements per staged row
  __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(sk_smem);
  const int ut = blockIdx.x * (kSkRows / 16) + warp;
  const bool live = ut < a.ut_total;

  // W fragments: issue the first ring of loads before touching the activations
  uint4 wa[kSkPF], wb[kSkPF];
  const uint4* wp = a.Wp + ((int64_t)ut * a.kb_total + kb_lo) * 64 + lane;
  if (live) {
#pragma unroll
    for (int q = 0; q < kSkPF; ++q)
      if (q < nkb) { wa[q] = ldg_stream16(wp + (int64_t)q * 64); wb[q] = ldg_stream16(wp + (int64_t)q * 64 + 32); }
  }
  // programmatic dependent launch: everything above only touched constant weights; the activations
  // (and Cin / tickets) come from the preceding kernel in the stream
  asm volatile("griddepcontrol.wait;\n" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  stage_activations<NB>(a, As, pitch, kb_lo, nkb, warp, lane);
  __syncthreads();

  float acc[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.0f;
  if (live) {
    for (int blk = 0; blk < nkb; blk += kSkPF) {
#pragma unroll
      for (int q = 0; q < kSkPF; ++q) {
        const int cur = blk + q;
        if (cur < nkb) {
#pragma unroll
          for (int n = 0; n < NT; ++n) {
            const uint4 bv = *reinterpret_cast<const uint4*>(As + (size_t)(n * 8 + g) * pitch + cur * 32 + 8 * t);
            mma16816(acc[n], wa[q], bv.x, bv.y);
            mma16816(acc[n], wb[q], bv.z, bv.w);
          }
          if (cur + kSkPF < nkb) {
            wa[q] = ldg_stream16(wp + (int64_t)(cur + kSkPF) * 64);
            wb[q] = ldg_stream16(wp + (int64_t)(cur + kSkPF) * 64 + 32);
          }
        }
      }
    }
  }

  // accumulator fragment: c0 (row g, col 2t), c1 (g, 2t+1), c2 (g+8, 2t), c3 (g+8, 2t+1);
  // "row" is the weight row (output column nn), "col" the activation row m.
  if (gridDim.y == 1) {
    if (live) {
      float cin[NT][4], bia[2] = {0.f, 0.f};
      const int nn0 = ut * 16 + g;
      if (a.bias) {
        if (nn0 < a.N) bia[0] = __ldg(a.bias + nn0);
        if (nn0 + 8 < a.N) bia[1] = __ldg(a.bias + nn0 + 8);
      }
      // all addend loads first (C may alias Cin, so the compiler cannot hoist them itself)
#pragma unroll
      for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int m = n * 8 + 2 * t + (e & 1), nn = nn0 + (e >> 1) * 8;
          cin[n][e] = (a.Cin && m < a.M && nn < a.N) ? __ldcg(a.Cin + (int64_t)m * a.ldcin + nn) : 0.0f;
        }
#pragma unroll
      for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int m = n * 8 + 2 * t + (e & 1), nn = nn0 + (e >> 1) * 8;
          if (m < a.M && nn < a.N) {
            const float v = (acc[n][e] + bia[e >> 1]) * a.alpha + a.beta * cin[n][e];
            a.C[(int64_t)m * a.ldc + nn] = v;
            if (a.Clp) a.Clp[(int64_t)m * a.ldclp + nn] = __float2bfloat16_rn(v);
          }
        }
    }
    return;
  }
  // split-K: partial tile [NB][128] per (split, n-group), reduced by the last CTA of the group
  float* mine = a.part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (NB * kSkRows);
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    const int m = n * 8 + 2 * t, nl = warp * 16 + g;
    __stcg(&mine[(m) * kSkRows + nl], acc[n][0]);
    __stcg(&mine[(m + 1) * kSkRows + nl], acc[n][1]);
    __stcg(&mine[(m) * kSkRows + nl + 8], acc[n][2]);
    __stcg(&mine[(m + 1) * kSkRows + nl + 8], acc[n][3]);
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = (atomicAdd(&a.tickets[blockIdx.x], 1u) == gridDim.y - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // element e = tid + 256*i: all EPT loads of one split are in flight together (the naive
  // element-outer loop costs EPT * splits serialized L2 round trips)
  constexpr int EPT = NB * kSkRows / kSkThreads;
  float sum[EPT], cin[EPT];
#pragma unroll
  for (int i = 0; i < EPT; ++i) sum[i] = 0.0f;
  const size_t zstride = (size_t)gridDim.x * (NB * kSkRows);
  const float* p0 = a.part + (size_t)blockIdx.x * (NB * kSkRows) + tid;
  for (unsigned z = 0; z < gridDim.y; ++z) {
    float v[EPT];
#pragma unroll
    for (int i = 0; i < EPT; ++i) v[i] = __ldcg(p0 + z * zstride + i * kSkThreads);
#pragma unroll
    for (int i = 0; i < EPT; ++i) sum[i] += v[i];
  }
  const int nl = tid % kSkRows, nn = blockIdx.x * kSkRows + nl;
  const float bia = (a.bias && nn < a.N) ? __ldg(a.bias + nn) : 0.0f;
#pragma unroll
  for (int i = 0; i < EPT; ++i) {
    const int m = (tid + i * kSkThreads) / kSkRows;
    cin[i] = (a.Cin && m < a.M && nn < a.N) ? __ldcg(a.Cin + (int64_t)m * a.ldcin + nn) : 0.0f;
  }
#pragma unroll
  for (int i = 0; i < EPT; ++i) {
    const int m = (tid + i * kSkThreads) / kSkRows;
    if (m < a.M && nn < a.N) {
      const float v = (sum[i] + bia) * a.alpha + a.beta * cin[i];
      a.C[(int64_t)m * a.ldc + nn] = v;
      if (a.Clp) a.Clp[(int64_t)m * a.ldclp + nn] = __float2bfloat16_rn(v);
    }
  }
  if (tid == 0) a.tickets[blockIdx.x] = 0;
}

// Single-phase variant for small layers (N <= ~1024): one 16-row weight tile per CTA, the 8 warps
// split K, partial sums meet in shared memory, no cross-CTA reduction and no second dependent pass.
template <int NT>
__global__ void __launch_bounds__(kSkThreads, 1) k_skinny_mt(const SkArgs a) {
  constexpr int NB = NT * 8, RPW = 17;
  extern __shared__ __align__(16) unsigned char sk_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int nkb = a.kb_total;
  const int pitch = ((nkb * 32 + 63) / 64) * 64 + 32;
  __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(sk_smem);
  float* red = reinterpret_cast<float*>(sk_smem + (size_t)NB * pitch * 2);      // [8][NB][17]
  const int ut = blockIdx.x;
  const int b_lo = (warp * nkb) / 8, b_hi = ((warp + 1) * nkb) / 8, nb = b_hi - b_lo;

  uint4 wa[kSkPF], wb[kSkPF];
  const uint4* wp = a.Wp + ((int64_t)ut * a.kb_total + b_lo) * 64 + lane;
#pragma unroll
  for (int q = 0; q < kSkPF; ++q)
    if (q < nb) { wa[q] = ldg_stream16(wp + (int64_t)q * 64); wb[q] = ldg_stream16(wp + (int64_t)q * 64 + 32); }
  asm volatile("griddepcontrol.wait;\n" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  stage_activations<NB>(a, As, pitch, 0, nkb, warp, lane);
  __syncthreads();

  float acc[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.0f;
  for (int blk = 0; blk < nb; blk += kSkPF) {
#pragma unroll
    for (int q = 0; q < kSkPF; ++q) {
      const int cur = blk + q;
      if (cur < nb) {
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          const uint4 bv = *reinterpret_cast<const uint4*>(As + (size_t)(n * 8 + g) * pitch + (b_lo + cur) * 32 + 8 * t);
          mma16816(acc[n], wa[q], bv.x, bv.y);
          mma16816(acc[n], wb[q], bv.z, bv.w);
        }
        if (cur + kSkPF < nb) {
          wa[q] = ldg_stream16(wp + (int64_t)(cur + kSkPF) * 64);
          wb[q] = ldg_stream16(wp + (int64_t)(cur + kSkPF) * 64 + 32);
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
  constexpr int EPT = NB * 16 / kSkThreads;                     // 2 (NB = 32) or 4 (NB = 64); NB = 8 -> 1 with a guard
  float sum[EPT > 0 ? EPT : 1], cin[EPT > 0 ? EPT : 1];
  constexpr int E = EPT > 0 ? EPT : 1;
#pragma unroll
  for (int i = 0; i < E; ++i) {
    const int idx = tid + i * kSkThreads, m = idx >> 4, r = idx & 15, nn = ut * 16 + r;
    sum[i] = 0.0f; cin[i] = 0.0f;
    if (idx < NB * 16) {
#pragma unroll
      for (int w = 0; w < 8; ++w) sum[i] += red[((size_t)w * NB + m) * RPW + r];
      if (a.Cin && m < a.M && nn < a.N) cin[i] = __ldcg(a.Cin + (int64_t)m * a.ldcin + nn);
    }
  }
#pragma unroll
  for (int i = 0; i < E; ++i) {
    const int idx = tid + i * kSkThreads, m = idx >> 4, r = idx & 15, nn = ut * 16 + r;
    if (idx < NB * 16 && m < a.M && nn < a.N) {
      const float v = (sum[i] + (a.bias ? __ldg(a.bias + nn) : 0.0f)) * a.alpha + a.beta * cin[i];
      a.C[(int64_t)m * a.ldc + nn] = v;
      if (a.Clp) a.Clp[(int64_t)m * a.ldclp + nn] = __float2bfloat16_rn(v);
    }
  }
}

}  // namespace tp

using namespace tp;

extern "C" size_t tp_pack_mma_a_bytes(int rows, int cols) {
  return (size_t)((rows + 15) / 16) * ((cols + 31) / 32) * 1024;
}

extern "C" int tp_pack_mma_a_bf16(const float* w, int64_t ld, int rows, int cols, void* dst, void* stream) {
  TP_CHECK_ARG(w && dst && rows > 0 && cols > 0 && ld >= cols, "tp_pack_mma_a_bf16: bad arguments");
  TP_CHECK_ARG(aligned16(dst), "tp_pack_mma_a_bf16: dst must be 16-byte aligned");
  k_pack_mma_a<<<1024, 256, 0, (cudaStream_t)stream>>>(w, ld, rows, cols, reinterpret_cast<uint32_t*>(dst));
  TP_LAUNCH_CHECK();
  return TP_OK;
}

extern "C" size_t tp_skinny_bf16_workspace_bytes(int M, int N, int splits) {
  const size_t nb = M <= 8 ? 8 : (M <= 32 ? 32 : 64);
  const size_t groups = ((size_t)N + kSkRows - 1) / kSkRows;
  return kSkTicketBytes + (size_t)(splits > 1 ? splits : 0) * groups * nb * kSkRows * sizeof(float);
}

template <typename KernelT>
static int launch_pdl(KernelT kfn, dim3 grid, size_t smem, cudaStream_t st, const SkArgs& a) {
  TP_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = dim3(kSkThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;      // tp_set_pdl(): the one switch of every PDL launch
  cfg.attrs = attr; cfg.numAttrs = 1;
  TP_CUDA(cudaLaunchKernelEx(&cfg, kfn, a));
  count_launch();
  return TP_OK;
}

// mode: 0 = auto, 1 = force split kernel, 2 = force single-phase kernel
static int skinny_dispatch(const float* A, int64_t lda, const void* Alp, int64_t ldalp, int M, int K, const void* Wp, int N,
                           const float* bias, const float* Cin, int64_t ldcin, float* C, int64_t ldc, void* Clp,
                           int64_t ldclp, float alpha, float beta, int relu_a, int splits, int mode, void* workspace,
                           size_t workspace_bytes, void* stream) {
  TP_CHECK_ARG((A || Alp) && Wp && C, "tp_skinny_bf16: null pointer");
  TP_CHECK_ARG(M >= 1 && M <= 64, "tp_skinny_bf16: M=%d must be in [1,64]", M);
  TP_CHECK_ARG(N >= 1 && K >= 4 && K % 4 == 0, "tp_skinny_bf16: need K %% 4 == 0 (K=%d)", K);
  TP_CHECK_ARG(!A || (lda % 4 == 0 && aligned16(A)), "tp_skinny_bf16: A must be 16-byte aligned with lda %% 4 == 0");
  TP_CHECK_ARG(!Alp || (K % 8 == 0 && ldalp % 8 == 0 && aligned16(Alp)), "tp_skinny_bf16: bf16 A needs K, ld %% 8 == 0 and 16-byte alignment");
  TP_CHECK_ARG(aligned16(Wp), "tp_skinny_bf16: Wp must be 16-byte aligned");
  SkArgs a;
  a.A = A; a.lda = lda; a.M = M; a.K = K; a.N = N;
  a.Alp = reinterpret_cast<const __nv_bfloat16*>(Alp); a.ldalp = ldalp;
  a.Clp = reinterpret_cast<__nv_bfloat16*>(Clp); a.ldclp = ldclp;
  a.Wp = reinterpret_cast<const uint4*>(Wp);
  a.ut_total = (N + 15) / 16; a.kb_total = (K + 31) / 32;
  a.bias = bias; a.Cin = Cin; a.ldcin = ldcin; a.C = C; a.ldc = ldc; a.alpha = alpha; a.beta = beta; a.relu_a = relu_a;
  a.part = nullptr; a.tickets = nullptr; a.kb_per_split = a.kb_total;
  const int nb = M <= 8 ? 8 : (M <= 32 ? 32 : 64);
  cudaStream_t st = (cudaStream_t)stream;

  // single-phase kernel: whole K staged per CTA
  const int pitch_all = ((a.kb_total * 32 + 63) / 64) * 64 + 32;
  const size_t smem_mt = (size_t)nb * pitch_all * 2 + (size_t)8 * nb * 17 * 4;
  const bool mt_ok = smem_mt <= 200 * 1024;
  if (mode == 2 || (mode == 0 && mt_ok && N <= 1024)) {
    if (!mt_ok) return fail(TP_ERR_UNSUPPORTED, "tp_skinny_bf16: K=%d too large for the single-phase kernel", K);
    dim3 grid((unsigned)a.ut_total);
    if (nb == 8) return launch_pdl(k_skinny_mt<1>, grid, smem_mt, st, a);
    if (nb == 32) return launch_pdl(k_skinny_mt<4>, grid, smem_mt, st, a);
    return launch_pdl(k_skinny_mt<8>, grid, smem_mt, st, a);
  }

  const int groups = (N + kSkRows - 1) / kSkRows;
  if (splits < 1) splits = 1;
  if (splits > a.kb_total) splits = a.kb_total;
  a.kb_per_split = (a.kb_total + splits - 1) / splits;
  splits = (a.kb_total + a.kb_per_split - 1) / a.kb_per_split;
  if (splits > 1) {
    const size_t need = kSkTicketBytes + (size_t)splits * groups * nb * kSkRows * sizeof(float);
    TP_CHECK_ARG(workspace && workspace_bytes >= need && (size_t)groups * 4 <= kSkTicketBytes && aligned16(workspace),
                 "tp_skinny_bf16: split-K workspace too small (%zu < %zu)", workspace_bytes, need);
    a.tickets = reinterpret_cast<unsigned int*>(workspace);
    a.part = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + kSkTicketBytes);
  }
  const int pitch = ((a.kb_per_split * 32 + 63) / 64) * 64 + 32;
  const size_t smem = (size_t)nb * pitch * 2;
  if (smem > 200 * 1024) return fail(TP_ERR_UNSUPPORTED, "tp_skinny_bf16: K slice of %d columns does not fit in shared memory; raise splits", a.kb_per_split * 32);
  dim3 grid((unsigned)groups, (unsigned)splits);
  if (nb == 8) return launch_pdl(k_skinny_bf16<1>, grid, smem, st, a);
  if (nb == 32) return launch_pdl(k_skinny_bf16<4>, grid, smem, st, a);
  return launch_pdl(k_skinny_bf16<8>, grid, smem, st, a);
}

extern "C" int tp_skinny_bf16(const float* A, int64_t lda, int M, int K, const void* Wp, int N, const float* bias,
                              const float* Cin, int64_t ldcin, float* C, int64_t ldc, float alpha, float beta,
                              int relu_a, int splits, void* workspace, size_t workspace_bytes, void* stream) {
  TP_CHECK_ARG(A != nullptr, "tp_skinny_bf16: null A");
  return skinny_dispatch(A, lda, nullptr, 0, M, K, Wp, N, bias, Cin, ldcin, C, ldc, nullptr, 0, alpha, beta, relu_a,
                         splits, splits > 1 ? 1 : 0, workspace, workspace_bytes, stream);
}

extern "C" int tp_skinny_bf16_ex(const float* A, int64_t lda, const void* A_bf16, int64_t lda_bf16, int M, int K,
                                 const void* Wp, int N, const float* bias, const float* Cin, int64_t ldcin, float* C,
                                 int64_t ldc, void* C_bf16, int64_t ldc_bf16, float alpha, float beta, int relu_a,
                                 int splits, int mode, void* workspace, size_t workspace_bytes, void* stream) {
  return skinny_dispatch(A, lda, A_bf16, lda_bf16, M, K, Wp, N, bias, Cin, ldcin, C, ldc, C_bf16, ldc_bf16, alpha, beta,
                         relu_a, splits, mode, workspace, workspace_bytes, stream);
}
