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
           int M, int N, int K, float alpha, float beta) {
  constexpr int BK = 32, LD = BK + 4, STAGES = 3;
  constexpr int RY = BM / TM, RX = BN / TN, NT = RY * RX;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                         // [STAGES][BM][LD]
  float* Ws = smem + STAGES * BM * LD;      // [STAGES][BN][LD]
  const int tid = threadIdx.x, tx = tid % RX, ty = tid / RX;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nkt = (K + BK - 1) / BK;

  auto load_stage = [&](int stage, int kt) {
    const int k0 = kt * BK;
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

template <int BM, int BN, int TM, int TN>
static int launch(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, const float* Cin,
                  int64_t ldcin, float* C, int64_t ldc, int M, int N, int K, float alpha, float beta, int relu_a,
                  cudaStream_t st) {
  constexpr int LD = 36, STAGES = 3;
  constexpr size_t smem = (size_t)STAGES * (BM + BN) * LD * sizeof(float);
  dim3 grid((unsigned)ceil_div(N, BN), (unsigned)ceil_div(M, BM));
  dim3 block((BM / TM) * (BN / TN));
  if (relu_a) {
    auto kfn = k_gemm_f32<BM, BN, TM, TN, true>;
    TP_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kfn<<<grid, block, smem, st>>>(A, lda, W, ldw, bias, Cin, ldcin, C, ldc, M, N, K, alpha, beta);
  } else {
    auto kfn = k_gemm_f32<BM, BN, TM, TN, false>;
    TP_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kfn<<<grid, block, smem, st>>>(A, lda, W, ldw, bias, Cin, ldcin, C, ldc, M, N, K, alpha, beta);
  }
  TP_LAUNCH_CHECK();
  return TP_OK;
}

}  // namespace tp

extern "C" int tp_gemm_f32(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                           const float* Cin, int64_t ldcin, float* C, int64_t ldc, int M, int N, int K,
                           float alpha, float beta, int relu_a, void* stream) {
  using namespace tp;
  TP_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "tp_gemm_f32: negative size");
  if (M == 0 || N == 0) return TP_OK;
  TP_CHECK_ARG(A && W && C, "tp_gemm_f32: null pointer");
  TP_CHECK_ARG(K % 4 == 0 && lda % 4 == 0 && ldw % 4 == 0, "tp_gemm_f32: K=%d lda=%lld ldw=%lld must be multiples of 4",
               K, (long long)lda, (long long)ldw);
  TP_CHECK_ARG(aligned16(A) && aligned16(W), "tp_gemm_f32: A/W must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (M <= 32) return launch<32, 32, 2, 4>(A, lda, W, ldw, bias, Cin, ldcin, C, ldc, M, N, K, alpha, beta, relu_a, st);
  if (M <= 64) return launch<64, 32, 4, 4>(A, lda, W, ldw, bias, Cin, ldcin, C, ldc, M, N, K, alpha, beta, relu_a, st);
  return launch<128, 64, 8, 4>(A, lda, W, ldw, bias, Cin, ldcin, C, ldc, M, N, K, alpha, beta, relu_a, st);
}
