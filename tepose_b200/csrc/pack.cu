// Operand packing: [B,T,K] fp32 -> time-major, zero-padded [T*B, Kp] fp32 / bf16.
#include "common.cuh"
#include <cuda_fp16.h>

namespace tp {

__device__ __forceinline__ float in_f32(float v) { return v; }
__device__ __forceinline__ float in_f32(__half v) { return __half2float(v); }

template <typename OutT, typename InT = float>
__global__ void k_pack_rows(const InT* __restrict__ src, int64_t stride_b, int64_t stride_t,
                            int rows_b, int rows_t, int k, OutT* __restrict__ dst, int kp, int relu,
                            unsigned int* __restrict__ zero, int zero_words) {
  pdl_launch_dependents();            // the GEMM that consumes dst may set itself up now (it waits before reading)
  // first kernel of a step: it also clears the grid-barrier slots of the persistent kernels that follow (saves a fill node)
  if (zero && blockIdx.x == 0)
    for (int i = threadIdx.x; i < zero_words; i += blockDim.x) zero[i] = 0u;
  int row = blockIdx.x;               // t * rows_b + b
  int t = row / rows_b, b = row - t * rows_b;
  const InT* s = src + (int64_t)b * stride_b + (int64_t)t * stride_t;
  OutT* d = dst + (int64_t)row * kp;
  if constexpr (sizeof(OutT) == 2) {
    // two columns per thread and store (kp is even); the loads of five column pairs are issued before the first store, so a
    // 2176-wide row is ONE memory round trip per thread instead of five dependent ones
    constexpr int U = 5;
    for (int c0 = 2 * threadIdx.x; c0 < kp; c0 += U * 2 * blockDim.x) {
      float v0[U], v1[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = c0 + u * 2 * blockDim.x;
        v0[u] = (c < k) ? in_f32(s[c]) : 0.0f;
        v1[u] = (c + 1 < k) ? in_f32(s[c + 1]) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = c0 + u * 2 * blockDim.x;
        if (c < kp) {
          float a = v0[u], b = v1[u];
          if (relu) { a = fmaxf(a, 0.0f); b = fmaxf(b, 0.0f); }
          *reinterpret_cast<__nv_bfloat162*>(d + c) = __floats2bfloat162_rn(a, b);
        }
      }
    }
  } else {
    for (int c = threadIdx.x; c < kp; c += blockDim.x) {
      float v = (c < k) ? in_f32(s[c]) : 0.0f;
      if (relu) v = fmaxf(v, 0.0f);
      d[c] = v;
    }
  }
}

// Inverse of k_pack_rows with the residual of lib/models/vibe.py:60-62 folded in:
// out[b*rows_t + t, c] = y[t*rows_b + b, c] (+ x[b, t, c]); an optional bf16 copy rides along.
__global__ void k_unpack_rows_residual(const float* __restrict__ y, int64_t ld_y, const float* __restrict__ x, int64_t stride_b,
                                       int64_t stride_t, int rows_b, int rows_t, int k, float* __restrict__ out,
                                       __nv_bfloat16* __restrict__ out_lp) {
  int row = blockIdx.x;               // b * rows_t + t
  int b = row / rows_t, t = row - b * rows_t;
  const float* s = y + ((int64_t)t * rows_b + b) * ld_y;
  const float* r = x ? x + (int64_t)b * stride_b + (int64_t)t * stride_t : nullptr;
  float* d = out + (int64_t)row * k;
  for (int c = 2 * threadIdx.x; c < k; c += 2 * blockDim.x) {     // k is even
    float v0 = s[c], v1 = s[c + 1];
    if (r) { v0 += r[c]; v1 += r[c + 1]; }
    *reinterpret_cast<float2*>(d + c) = make_float2(v0, v1);
    if (out_lp) *reinterpret_cast<__nv_bfloat162*>(out_lp + (int64_t)row * k + c) = __floats2bfloat162_rn(v0, v1);
  }
}

}  // namespace tp

extern "C" int tp_pack_rows_ex(const float* src, int64_t stride_b, int64_t stride_t, int rows_b, int rows_t,
                               int k, void* dst, int kp, int dst_precision, int relu, void* zero, size_t zero_bytes, void* stream);

extern "C" int tp_pack_rows(const float* src, int64_t stride_b, int64_t stride_t, int rows_b, int rows_t,
                            int k, void* dst, int kp, int dst_precision, int relu, void* stream) {
  return tp_pack_rows_ex(src, stride_b, stride_t, rows_b, rows_t, k, dst, kp, dst_precision, relu, nullptr, 0, stream);
}

template <typename InT>
static int pack_rows_impl(const InT* src, int64_t stride_b, int64_t stride_t, int rows_b, int rows_t,
                          int k, void* dst, int kp, int dst_precision, int relu, void* zero, size_t zero_bytes, void* stream) {
  using namespace tp;
  TP_CHECK_ARG(!zero || ((reinterpret_cast<uintptr_t>(zero) & 3) == 0 && zero_bytes % 4 == 0 && zero_bytes <= (1u << 20)),
               "tp_pack_rows_ex: zero region must be 4-byte aligned, a multiple of 4 bytes and <= 1 MB");
  unsigned int* zp = reinterpret_cast<unsigned int*>(zero);
  const int zw = zero ? (int)(zero_bytes / 4) : 0;
  TP_CHECK_ARG(rows_b >= 0 && rows_t >= 0 && k >= 0 && kp >= k, "tp_pack_rows: bad sizes");
  TP_CHECK_ARG(kp % 8 == 0, "tp_pack_rows: kp=%d must be a multiple of 8", kp);
  if (rows_b == 0 || rows_t == 0 || kp == 0) {
    if (zero && zw) TP_CUDA(cudaMemsetAsync(zero, 0, zero_bytes, (cudaStream_t)stream));
    return TP_OK;
  }
  TP_CHECK_ARG(src && dst, "tp_pack_rows: null pointer");
  unsigned grid = (unsigned)(rows_b * rows_t);
  if (dst_precision == TP_PRECISION_BF16)
    k_pack_rows<__nv_bfloat16, InT><<<grid, 256, 0, (cudaStream_t)stream>>>(src, stride_b, stride_t, rows_b, rows_t, k,
                                                                           (__nv_bfloat16*)dst, kp, relu, zp, zw);
  else
    k_pack_rows<float, InT><<<grid, 256, 0, (cudaStream_t)stream>>>(src, stride_b, stride_t, rows_b, rows_t, k,
                                                                   (float*)dst, kp, relu, zp, zw);
  TP_LAUNCH_CHECK();
  return TP_OK;
}

extern "C" int tp_pack_rows_ex(const float* src, int64_t stride_b, int64_t stride_t, int rows_b, int rows_t,
                               int k, void* dst, int kp, int dst_precision, int relu, void* zero, size_t zero_bytes, void* stream) {
  return pack_rows_impl<float>(src, stride_b, stride_t, rows_b, rows_t, k, dst, kp, dst_precision, relu, zero, zero_bytes, stream);
}

extern "C" int tp_pack_rows_f16(const void* src, int64_t stride_b, int64_t stride_t, int rows_b, int rows_t,
                                int k, void* dst, int kp, int dst_precision, int relu, void* zero, size_t zero_bytes, void* stream) {
  return pack_rows_impl<__half>(reinterpret_cast<const __half*>(src), stride_b, stride_t, rows_b, rows_t, k, dst, kp, dst_precision, relu,
                                zero, zero_bytes, stream);
}

namespace tp {
__global__ void k_split3(const float* __restrict__ src, int64_t ld, int rows, int k, __nv_bfloat16* __restrict__ dst) {
  const int64_t n4 = (int64_t)rows * (k / 4);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / (k / 4)), c = (int)(i - (int64_t)r * (k / 4)) * 4;
    const float4 v = *reinterpret_cast<const float4*>(src + (int64_t)r * ld + c);
    const float x[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { hi[e] = __float2bfloat16_rn(x[e]); lo[e] = __float2bfloat16_rn(x[e] - __bfloat162float(hi[e])); }
    __nv_bfloat16* d = dst + (int64_t)r * 3 * k + c;
    *reinterpret_cast<uint2*>(d) = *reinterpret_cast<const uint2*>(hi);
    *reinterpret_cast<uint2*>(d + k) = *reinterpret_cast<const uint2*>(lo);
    *reinterpret_cast<uint2*>(d + 2 * k) = *reinterpret_cast<const uint2*>(hi);
  }
}
}  // namespace tp

extern "C" int tp_split3_bf16(const float* src, int64_t ld_src, int rows, int k, void* dst, void* stream) {
  TP_CHECK_ARG(src && dst && rows > 0 && k > 0 && k % 8 == 0 && ld_src % 4 == 0, "tp_split3_bf16: need k %% 8 == 0, ld %% 4 == 0 (rows=%d k=%d)", rows, k);
  TP_CHECK_ARG(tp::aligned16(src) && tp::aligned16(dst), "tp_split3_bf16: pointers must be 16-byte aligned");
  const int64_t n4 = (int64_t)rows * (k / 4);
  const unsigned grid = (unsigned)(n4 / 256 + 1 < 2048 ? n4 / 256 + 1 : 2048);
  tp::k_split3<<<grid, 256, 0, (cudaStream_t)stream>>>(src, ld_src, rows, k, reinterpret_cast<__nv_bfloat16*>(dst));
  TP_LAUNCH_CHECK();
  return TP_OK;
}

extern "C" int tp_unpack_rows_residual(const float* y, int64_t ld_y, const float* x, int64_t stride_b, int64_t stride_t,
                                       int rows_b, int rows_t, int k, float* out, void* out_bf16, void* stream) {
  using namespace tp;
  TP_CHECK_ARG(rows_b >= 0 && rows_t >= 0 && k >= 0 && k % 2 == 0 && ld_y >= k && ld_y % 2 == 0,
               "tp_unpack_rows_residual: bad sizes (k=%d must be even, ld_y >= k)", k);
  if (rows_b == 0 || rows_t == 0 || k == 0) return TP_OK;
  TP_CHECK_ARG(y && out, "tp_unpack_rows_residual: null pointer");
  TP_CHECK_ARG(((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(out)) & 7) == 0 &&
               (!x || ((reinterpret_cast<uintptr_t>(x) & 3) == 0)), "tp_unpack_rows_residual: misaligned pointer");
  k_unpack_rows_residual<<<(unsigned)(rows_b * rows_t), 256, 0, (cudaStream_t)stream>>>(
      y, ld_y, x, stride_b, stride_t, rows_b, rows_t, k, out, reinterpret_cast<__nv_bfloat16*>(out_bf16));
  TP_LAUNCH_CHECK();
  return TP_OK;
}
