// K3 -- encoder linear heads + IEF Regressor (reference lib/models/tepose.py:79-85 and
// lib/models/spin.py:250-261).  v1: a fixed sequence of fp32 FFMA GEMM launches on one
// stream (graph-capturable); fc1 is evaluated as  x.W1x^T (once)  +  [pose|shape|cam].W1p^T
// (per iteration) -- algebraically identical to fc1(cat[x, pose, shape, cam]).
#include "common.cuh"

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
  if (precision == TP_PRECISION_BF16 && M <= 64) {
    const int groups = (N + 127) / 128, kb = (K + 31) / 32;
    int splits = (sms + groups - 1) / groups;
    if (splits > kb / 4) splits = kb / 4;
    if (splits < 1 || !scratch) splits = 1;
    while (splits > 1 && tp_skinny_bf16_workspace_bytes(M, N, splits) > kSplitScratch) --splits;
    return tp_skinny_bf16_ex(A, lda, Alp, ldalp, M, K, W, N, bias, Cin, ldcin, C, ldc, Clp, ldclp, alpha, beta, relu_a,
                             splits, 0, scratch, kSplitScratch, stream);
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
  if (B > 64) {   // the split-K scratch is sized for <= 64 rows
    sc = nullptr;
  }
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
  if (N > 64) sc = nullptr;
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
