// fp32 FFMA GEMM  C = alpha*(act(A).W^T + bias) + beta*Cin   (A [M,K], W [N,K], both K-major).
// Strict-parity path for nn.Linear / the GRU input projection: plain fp32 FMAs, fp32
// accumulate, no tensor cores.  Operand tiles stay K-major in shared memory (cp.async,
// 3 stages); each thread owns rows {ty + i*RY} x {tx + j*RX} so that the float4 reads along
// K are bank-conflict free with a row pitch of BK+4 floats.
#include "common.cuh"

namespace tp {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int BM, int BN, int TM, int TN, bool RELU_A>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
k_gemm_f32(const float* __restrict__ A, int64_t lda, const float* __restrict__ W, int64_t ldw,
           const float* __restrict__ bias, const float* Cin, int64_t ldcin, float* C, int64_t ldc,
           int M, int N, int K, float alpha, float beta, float* part, unsigned int* tickets) {
  constexpr int BK = 32, LD = BK + 4, STAGES = 3;
  constexpr int RY = BM / TM, RX = BN / TN, NT = RY * RX;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                         // [STAGES][BM][LD]
  float* Ws = smem + STAGES * BM * LD;      // [STAGES][BN][LD]
  const int tid = threadIdx.x, tx = tid % RX, ty = tid / RX;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  // split-K: blockIdx.z owns k-tiles [kt0, kt0 + nkt)
  const int nkt_all = (K + BK - 1) / BK;
  const int per = (nkt_all + gridDim.z - 1) / gridDim.z;
  const int kt0 = blockIdx.z * per;
  const int nkt = max(0, min(per, nkt_all - kt0));

  auto load_stage = [&](int stage, int kt) {
    const int k0 = (kt0 + kt) * BK;
    for (int c = tid; c < (BM + BN) * (BK / 4); c += NT) {
      int r = c / (BK / 4), q = c % (BK / 4);
      int k = k0 + q * 4;
      if (r < BM) {
        int m = m0 + r;
        bool ok = (m < M) && (k < K);
        cp_async16(&As[(stage * BM + r) * LD + q * 4], ok ? (A + (int64_t)m * lda + k) : A, ok);
      } else {
        int rr = r - BM, n = n0 + rr;
        bool ok = (n < N) && (k < K);
        cp_async16(&Ws[(stage * BN + rr) * LD + q * 4], ok ? (W + (int64_t)n * ldw + k) : W, ok);
      }
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nkt) load_stage(s, s);
    cp_async_commit();
  }
  for (int kt = 0; kt < nkt; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {  // prefetch tile kt+STAGES-1 into the slot freed at iteration kt-1
      int nk = kt + STAGES - 1;
      if (nk < nkt) load_stage(nk % STAGES, nk);
      cp_async_commit();
    }
    const float* as = As + (kt % STAGES) * BM * LD;
    const float* ws = Ws + (kt % STAGES) * BN * LD;
#pragma unroll
    for (int q = 0; q < BK / 4; ++q) {
      float4 a[TM], w[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        a[i] = *reinterpret_cast<const float4*>(&as[(ty + i * RY) * LD + q * 4]);
        if (RELU_A) { a[i].x = fmaxf(a[i].x, 0.f); a[i].y = fmaxf(a[i].y, 0.f); a[i].z = fmaxf(a[i].z, 0.f); a[i].w = fmaxf(a[i].w, 0.f); }
      }
#pragma unroll
      for (int j = 0; j < TN; ++j) w[j] = *reinterpret_cast<const float4*>(&ws[(tx + j * RX) * LD + q * 4]);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          acc[i][j] = fmaf(a[i].x, w[j].x, acc[i][j]);
          acc[i][j] = fmaf(a[i].y, w[j].y, acc[i][j]);
          acc[i][j] = fmaf(a[i].z, w[j].z, acc[i][j]);
          acc[i][j] = fmaf(a[i].w, w[j].w, acc[i][j]);
        }
    }
  }
  cp_async_wait<0>();

  if (gridDim.z > 1) {
    // publish this split's partial tile; the last CTA to arrive reduces all splits in split order
    // (deterministic) and applies the epilogue.  Tickets reset themselves for the next launch.
    __shared__ int s_last;
    const size_t tile = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    float* mine = part + ((size_t)blockIdx.z * gridDim.x * gridDim.y + tile) * (BM * BN);
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) __stcg(&mine[(ty + i * RY) * BN + tx + j * RX], acc[i][j]);
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&tickets[tile], 1u) == gridDim.z - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;
    for (unsigned z = 0; z < gridDim.z; ++z) {
      const float* src = part + ((size_t)z * gridDim.x * gridDim.y + tile) * (BM * BN);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] += __ldcg(&src[(ty + i * RY) * BN + tx + j * RX]);
    }
    if (tid == 0) tickets[tile] = 0;
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty + i * RY;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx + j * RX;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      v *= alpha;
      if (Cin) v += beta * Cin[(int64_t)m * ldcin + n];
      C[(int64_t)m * ldc + n] = v;
    }
  }
}

constexpr size_t kSplitTicketBytes = 4096;   // zero-initialised ticket counters at the head of the scratch

template <int BM, int BN, int TM, int TN>
static int launch(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, const float* Cin,
                  int64_t ldcin, float* C, int64_t ldc, int M, int N, int K, float alpha, float beta, int relu_a,
                  cudaStream_t st, int splits, void* ws, size_t ws_bytes) {
  constexpr int LD = 36, STAGES = 3;
  constexpr size_t smem = (size_t)STAGES * (BM + BN) * LD * sizeof(float);
  dim3 grid((unsigned)ceil_div(N, BN), (unsigned)ceil_div(M, BM));
  dim3 block((BM / TM) * (BN / TN));
  float* part = nullptr;
  unsigned int* tickets = nullptr;
  const int nkt_all = (K + 31) / 32;
  if (splits > nkt_all) splits = nkt_all;
  if (splits > 1) {
    const size_t tiles = (size_t)grid.x * grid.y;
    const size_t need = kSplitTicketBytes + (size_t)splits * tiles * BM * BN * sizeof(float);
    if (!ws || ws_bytes < need || tiles * sizeof(unsigned) > kSplitTicketBytes || (reinterpret_cast<uintptr_t>(ws) & 15))
      splits = 1;   // not enough scratch: fall back to the single-pass kernel (same result)
    else {
      tickets = reinterpret_cast<unsigned int*>(ws);
      part = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(ws) + kSplitTicketBytes);
      grid.z = (unsigned)splits;
    }
  }
  if (relu_a) {
    auto kfn = k_gemm_f32<BM, BN, TM, TN, true>;
    TP_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kfn<<<grid, block, smem, st>>>(A, lda, W, ldw, bias, Cin, ldcin, C, ldc, M, N, K, alpha, beta, part, tickets);
  } else {
    auto kfn = k_gemm_f32<BM, BN, TM, TN, false>;
    TP_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kfn<<<grid, block, smem, st>>>(A, lda, W, ldw, bias, Cin, ldcin, C, ldc, M, N, K, alpha, beta, part, tickets);
  }
  TP_LAUNCH_CHECK();
  return TP_OK;
}

}  // namespace tp

static int gemm_f32_dispatch(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                             const float* Cin, int64_t ldcin, float* C, int64_t ldc, int M, int N, int K,
                             float alpha, float beta, int relu_a, void* stream, int splits, void* ws, size_t ws_bytes,
                             const char* who) {
  using namespace tp;
  TP_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "%s: negative size", who);
  if (M == 0 || N == 0) return TP_OK;
  TP_CHECK_ARG(A && W && C, "%s: null pointer", who);
  TP_CHECK_ARG(K % 4 == 0 && lda % 4 == 0 && ldw % 4 == 0, "%s: K=%d lda=%lld ldw=%lld must be multiples of 4", who,
               K, (long long)lda, (long long)ldw);
  TP_CHECK_ARG(aligned16(A) && aligned16(W), "%s: A/W must be 16-byte aligned", who);
  cudaStream_t st = (cudaStream_t)stream;
  if (M <= 32) return launch<32, 32, 2, 4>(A, lda, W, ldw, bias, Cin, ldcin, C, ldc, M, N, K, alpha, beta, relu_a, st, splits, ws, ws_bytes);
  if (M <= 64) return launch<64, 32, 4, 4>(A, lda, W, ldw, bias, Cin, ldcin, C, ldc, M, N, K, alpha, beta, relu_a, st, splits, ws, ws_bytes);
  return launch<128, 64, 8, 4>(A, lda, W, ldw, bias, Cin, ldcin, C, ldc, M, N, K, alpha, beta, relu_a, st, splits, ws, ws_bytes);
}

extern "C" int tp_gemm_f32(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                           const float* Cin, int64_t ldcin, float* C, int64_t ldc, int M, int N, int K,
                           float alpha, float beta, int relu_a, void* stream) {
  return gemm_f32_dispatch(A, lda, W, ldw, bias, Cin, ldcin, C, ldc, M, N, K, alpha, beta, relu_a, stream, 1, nullptr, 0,
                           "tp_gemm_f32");
}

extern "C" size_t tp_gemm_f32_splitk_workspace_bytes(int M, int N, int splits) {
  const size_t bm = M <= 32 ? 32 : (M <= 64 ? 64 : 128), bn = M <= 64 ? 32 : 64;
  const size_t tiles = ((size_t)N + bn - 1) / bn * (((size_t)M + bm - 1) / bm);
  return tp::kSplitTicketBytes + (size_t)(splits > 1 ? splits : 1) * tiles * bm * bn * sizeof(float);
}

extern "C" int tp_gemm_f32_splitk(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                                  const float* Cin, int64_t ldcin, float* C, int64_t ldc, int M, int N, int K,
                                  float alpha, float beta, int relu_a, int splits, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  return gemm_f32_dispatch(A, lda, W, ldw, bias, Cin, ldcin, C, ldc, M, N, K, alpha, beta, relu_a, stream, splits,
                           workspace, workspace_bytes, "tp_gemm_f32_splitk");
}
