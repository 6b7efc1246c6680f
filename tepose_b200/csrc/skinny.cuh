// Shared device pieces of the skinny (M <= 64) tensor-core GEMM kernels: fragment loads, mma wrapper,
// argument block, activation staging.  Used by skinny.cu and by the fused IEF kernel in regressor.cu.
#pragma once
#include "common.cuh"

namespace tp {

constexpr int kSkThreads = 256;
constexpr int kSkRows = 128;          // weight rows per CTA
constexpr int kSkPF = 8;              // 32-column blocks in flight per warp
constexpr size_t kSkTicketBytes = 4096;

__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void mma16816(float* c, const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

struct SkArgs {
  const float* A; int64_t lda; int M, K, N;
  const uint4* Wp; int ut_total, kb_total, kb_per_split;
  const float* bias; const float* Cin; int64_t ldcin; float* C; int64_t ldc;
  float alpha, beta; int relu_a;
  float* part; unsigned int* tickets;
  const __nv_bfloat16* Alp; int64_t ldalp;   // optional bf16 copy of A (read instead of A when non-null)
  __nv_bfloat16* Clp; int64_t ldclp;         // optional bf16 copy of the output (next layer's Alp)
};

// act(A)[0:NB, 32*kb_lo : 32*(kb_lo+nkb)] -> bf16 rows of `pitch` elements in shared memory.
// Reads the bf16 copy when the producer left one (no conversion, half the bytes), else converts fp32.
template <int NB>
__device__ __forceinline__ void stage_activations(const SkArgs& a, __nv_bfloat16* As, int pitch, int kb_lo, int nkb,
                                                  int warp, int lane) {
  if (a.Alp) {
    const int c8n = nkb * 4;                                   // 16-byte (8 x bf16) groups per row
#pragma unroll 4
    for (int r = warp; r < NB; r += kSkThreads / 32) {
#pragma unroll 2
      for (int c8 = lane; c8 < c8n; c8 += 32) {
        const int col = kb_lo * 32 + c8 * 8;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (r < a.M && col < a.K) v = *reinterpret_cast<const uint4*>(a.Alp + (int64_t)r * a.ldalp + col);  // K % 8 == 0
        if (a.relu_a) {
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
          const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 4; ++i) h[i] = __hmax2(h[i], z);
        }
        *reinterpret_cast<uint4*>(As + (size_t)r * pitch + c8 * 8) = v;
      }
    }
    return;
  }
  const int c4n = nkb * 8;                                     // float4 groups per row
#pragma unroll 4
  for (int r = warp; r < NB; r += kSkThreads / 32) {
#pragma unroll 2
    for (int c4 = lane; c4 < c4n; c4 += 32) {
      const int col = kb_lo * 32 + c4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < a.M && col < a.K) v = *reinterpret_cast<const float4*>(a.A + (int64_t)r * a.lda + col);  // K % 4 == 0
      if (a.relu_a) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
      uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      *reinterpret_cast<uint2*>(As + (size_t)r * pitch + c4 * 4) = pk;
    }
  }
}

}  // namespace tp
