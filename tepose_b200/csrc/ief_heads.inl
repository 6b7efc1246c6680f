// Front of K3 (bf16 mode, <= 32 rows): the encoder heads and the iteration-invariant part of fc1, split-K over groups of 8 CTAs.
//
// Reference: lib/models/tepose.py:79-85 (feat = (linear_fwd(relu(y[-1])) + linear_rec(relu(y_rec[0]))) / 2, as ONE GEMM over
// [h_fwd | h_rec] with the halved, concatenated weights -- see tp_encoder_heads_cat) and lib/models/spin.py:252 (fc1 on the feature
// columns of xc: base = W1[:, :2048] . feat + b1).
//
// The row-split version (k_ief_fused: CTA = 16 weight rows, all of K) made every CTA stage the WHOLE activation block (384 KB per
// CTA, 49 MB of L2->SM traffic for 25 MB of weights) and streamed its weights with 8 KB in flight per warp: 2.3 TB/s.  Here
//   * a group of 8 consecutive CTAs shares 128 weight rows and each CTA takes one eighth of K: it stages [32 x K/8] activations
//     straight from the fp32 encoder states (relu + bf16 on the way in: no conversion pass, no grid barrier before the first MMA);
//   * the CTA's whole weight slice (8 tiles x K/8 = 192 KB at H = 2048) is requested up front with cp.async.bulk into shared memory
//     (160 KB resident, the rest recycled) -- the weights are constants, so the requests go out BEFORE griddepcontrol.wait and the
//     25 MB stream is one HBM latency plus the transfer instead of a chain of 8 KB round trips;
//   * the eight partial sums of a 16-row tile meet in global memory (L2): every CTA stores its partials, arrives on the group's
//     counter (release) and, once the 8 arrivals are in, sums the tile it owns in a fixed order (L2-coherent loads);
//   * one grid barrier later the same scheme runs fc1's feature part (4 tiles x K/8 per CTA, owner = (tile, half of the batch)).
#pragma once
#include "skinny.cuh"
#include "umma.cuh"
#include "ief_cluster.inl"

namespace tp {

constexpr int kHbThreads = 256, kHbGroup = 8, kHbGrid = 128;
constexpr int kHbSlots = 5;                                // 4 KB weight chunks (4 k-blocks of one tile) resident per warp
constexpr uint32_t kHbOffW = 0;                            // [8 warps][5 slots][4 KB]
constexpr uint32_t kHbOffBar = 8 * kHbSlots * 4096;        // mbarriers: [8 warps][8 chunks] phase 1, [8 warps] phase 2
constexpr uint32_t kHbOffAs = kHbOffBar + 8 * (64 + 8);    // staged activations, bf16 rows of (K/8 + 32) elements
constexpr size_t kHbPartBytes = (size_t)16 * 16 * 4 * 2048;   // one partial-sum scratch: [16 groups][<=16 sources][<=8 tiles ...] (2 MB)

struct HeadsBaseParams {
  const float* hcat; int64_t ld_h; int KH;                // relu(hcat [M, KH]) is the heads' input (KH = 3H); null: feat / feat_lp given
  const uint4* w_cat; const float* b_cat;                 // packed [2048 x KH]
  const float* feat; const __nv_bfloat16* feat_lp;        // heads skipped: the feature [M, 2048] (bf16 copy preferred)
  __nv_bfloat16* feat_rep;                                // [M, 2048] bf16 scratch (phase 1 -> phase 2)
  const uint4* w1x; const float* b1;                      // packed [1024 x 2048]
  float* base;                                            // [M, 1024] out
  float* part1; float* part2;                             // partial sums: [16][8][8 tiles][512] and [16][16][4 tiles][512] floats
  int M;
  unsigned int* barrier; int barrier_shards;              // zeroed 1 KB slot: grid barrier (words 32 k) + the group counters (words 1.., 33..)
  long long* trace;
};

#define HB_TRACE(slot) do { if (p.trace && tid == 0) p.trace[(size_t)blockIdx.x * 16 + (slot)] = clock64(); } while (0)

// the 8 CTAs of a group meet on a counter: arrive (release, after the CTA's stores) and wait for all 8 (then read with __ldcg)
__device__ __forceinline__ void hb_group_sync(unsigned int* ctr) {
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int seen;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(ctr) : "memory");
    do {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];\n" : "=r"(seen) : "l"(ctr) : "memory");
    } while (seen < (unsigned)kHbGroup);
#ifdef TP_BARRIER_ACQUIRE_FENCE
    asm volatile("fence.acq_rel.gpu;\n" ::: "memory");
#endif
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kHbThreads, 1) k_heads_base(const HeadsBaseParams p) {
  extern __shared__ __align__(128) unsigned char hb_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int k = blockIdx.x % kHbGroup;                    // K split of this CTA
  const int rb = blockIdx.x / kHbGroup;                   // row block of the group
  const uint32_t sm = smem_u32(hb_smem);
  const int ntc = (p.M + 7) >> 3;
  __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(hb_smem + kHbOffAs);
  const bool heads = p.hcat != nullptr;
  auto bar1 = [&](int w, int c) { return sm + kHbOffBar + 8u * (w * 8 + c); };
  auto bar2 = [&](int w) { return sm + kHbOffBar + 8u * (64 + w); };
  const uint32_t wslot = sm + kHbOffW + warp * (kHbSlots * 4096);
  // phase-2 role: tile t2 of the group's four, K half kh of this CTA's 8 k-blocks; owner CTA of (tile, batch half) = 2 tile + half
  const int t2 = warp & 3, kh = warp >> 2;
  const int own_t = k >> 1, own_np = k & 1;
  const int own_cnt = ntc - 2 * own_np < 0 ? 0 : (ntc - 2 * own_np > 2 ? 2 : ntc - 2 * own_np);   // 8-row groups this CTA owns in phase 2

  const int nkb1 = heads ? p.KH / 32 / kHbGroup : 0;      // k-blocks of this CTA in phase 1 (a multiple of 4)
  const int nch = nkb1 >> 2;                              // 4 KB chunks per warp
  const int pitch1 = nkb1 * 32 + 32;
  const uint4* wp1 = heads ? p.w_cat + ((size_t)(8 * rb + warp) * (p.KH / 32) + (size_t)k * nkb1) * 64 : nullptr;
  const uint4* wp2 = p.w1x + ((size_t)(4 * rb + t2) * 64 + (size_t)k * 8 + kh * 4) * 64;     // 64 k-blocks per tile (K = 2048)
  auto fetch = [&](uint32_t bar, uint32_t dst, const void* src) {       // one 4 KB chunk, completion on `bar`
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(4096) : "memory");
    cl_bulk_g2s(dst, src, 4096, bar);
  };
  if (tid == 0) {
    for (int i = 0; i < 72; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(sm + kHbOffBar + 8u * i));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (lane == 0) {                                         // weights are constants: request them before the predecessor has finished
    if (heads) {
      for (int c = 0; c < nch && c < kHbSlots; ++c) fetch(bar1(warp, c), wslot + c * 4096, wp1 + (size_t)c * 256);
    } else {
      fetch(bar2(warp), wslot, wp2);
    }
  }
  pdl_wait();                                             // the encoder states / the feature come from the previous kernel
  pdl_launch_dependents();
  HB_TRACE(0);
  if (heads) {
    // ------------------------------------------------------------ phase 1: heads
    // relu(hcat[:, K slice]) -> bf16 rows in shared memory; float4 loads, up to 24 in flight per thread (one L2 round trip at H = 2048)
    const int c4n = nkb1 * 8;                             // float4 per row
    const int total = 8 * ntc * c4n;
    const float* src = p.hcat + (size_t)k * nkb1 * 32;
    constexpr int kInFlight = 24;
    for (int i0 = tid; i0 < total; i0 += kInFlight * kHbThreads) {
      float4 v[kInFlight];
#pragma unroll
      for (int u = 0; u < kInFlight; ++u) {
        const int i = i0 + u * kHbThreads, m = i / c4n, c4 = i - m * c4n;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < total && m < p.M) v[u] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)m * p.ld_h + c4 * 4));
      }
#pragma unroll
      for (int u = 0; u < kInFlight; ++u) {
        const int i = i0 + u * kHbThreads, m = i / c4n, c4 = i - m * c4n;
        if (i < total) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(fmaxf(v[u].x, 0.f), fmaxf(v[u].y, 0.f)), hi = __floats2bfloat162_rn(fmaxf(v[u].z, 0.f), fmaxf(v[u].w, 0.f));
          *reinterpret_cast<uint2*>(As + (size_t)m * pitch1 + c4 * 4) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        }
      }
    }
    __syncthreads();
    HB_TRACE(1);
    float acc[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.0f;
    for (int c = 0; c < nch; ++c) {
      mbar_wait(reinterpret_cast<uint64_t*>(hb_smem + kHbOffBar + 8 * (warp * 8 + c)), 0);
      const uint32_t ws = wslot + (c % kHbSlots) * 4096 + lane * 16;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 a = lds16(ws + q * 1024), b = lds16(ws + q * 1024 + 512);
        const int kb = 4 * c + q;
#pragma unroll
        for (int n = 0; n < 4; ++n)
          if (n < ntc) {
            const uint4 bv = *reinterpret_cast<const uint4*>(As + (size_t)(n * 8 + g) * pitch1 + kb * 32 + 8 * t);
            mma16816(acc[n], a, bv.x, bv.y);
            mma16816(acc[n], b, bv.z, bv.w);
          }
      }
      if (c + kHbSlots < nch) {                            // this chunk's slot takes chunk c + 5 (its fragments are in registers / consumed)
        __syncwarp();
        if (lane == 0) fetch(bar1(warp, c + kHbSlots), wslot + (c % kHbSlots) * 4096, wp1 + (size_t)(c + kHbSlots) * 256);
      }
    }
    __syncwarp();
    if (lane == 0) fetch(bar2(warp), wslot, wp2);          // phase-2 weights of this warp: in flight across the group sync and the grid barrier
    HB_TRACE(2);
    // partial sums of tile `warp` -> global [group][this K split][tile][n][lane][4]
    {
      float* dst = p.part1 + (((size_t)rb * 8 + k) * 8 + warp) * 512 + lane * 4;
#pragma unroll
      for (int n = 0; n < 4; ++n)
        if (n < ntc) __stcg(reinterpret_cast<float4*>(dst + n * 128), make_float4(acc[n][0], acc[n][1], acc[n][2], acc[n][3]));
    }
    hb_group_sync(p.barrier + 1 + rb);
    HB_TRACE(3);
    // owner of tile 8 rb + k: sum the 8 partials in K order, add the bias, leave the feature as bf16 rows
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int f = tid + 256 * r, n = f >> 7, ln = (f >> 2) & 31, j = f & 3;
      const int m = n * 8 + 2 * (ln & 3) + (j & 1), row = (8 * rb + k) * 16 + (ln >> 2) + 8 * (j >> 1);
      if (n < ntc && m < p.M) {
        float s = 0.0f;
#pragma unroll
        for (int src8 = 0; src8 < 8; ++src8) s += __ldcg(p.part1 + (((size_t)rb * 8 + src8) * 8 + k) * 512 + f);
        p.feat_rep[(size_t)m * 2048 + row] = __float2bfloat16_rn(s + p.b_cat[row]);
      }
    }
    grid_barrier_sh(p.barrier, 1, p.barrier_shards);
  }
  HB_TRACE(4);
  // -------------------------------------------------------------- phase 2: base = W1x . feat + b1
  {
    constexpr int pitch2 = 256 + 32;
    // this CTA's 256 feature columns of every row: 32 x 512 B
    const int total = 8 * ntc * 32;
    for (int i = tid; i < total; i += kHbThreads) {
      const int m = i >> 5, c8 = i & 31;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (m < p.M) {
        if (heads || p.feat_lp) {
          v = __ldcg(reinterpret_cast<const uint4*>((heads ? p.feat_rep : p.feat_lp) + (size_t)m * 2048 + k * 256 + c8 * 8));
        } else {
          const float4 f0 = __ldcg(reinterpret_cast<const float4*>(p.feat + (size_t)m * 2048 + k * 256 + c8 * 8));
          const float4 f1 = __ldcg(reinterpret_cast<const float4*>(p.feat + (size_t)m * 2048 + k * 256 + c8 * 8 + 4));
          __nv_bfloat162 q0 = __floats2bfloat162_rn(f0.x, f0.y), q1 = __floats2bfloat162_rn(f0.z, f0.w);
          __nv_bfloat162 q2 = __floats2bfloat162_rn(f1.x, f1.y), q3 = __floats2bfloat162_rn(f1.z, f1.w);
          v = make_uint4(*reinterpret_cast<uint32_t*>(&q0), *reinterpret_cast<uint32_t*>(&q1), *reinterpret_cast<uint32_t*>(&q2), *reinterpret_cast<uint32_t*>(&q3));
        }
      }
      *reinterpret_cast<uint4*>(As + (size_t)m * pitch2 + c8 * 8) = v;
    }
    __syncthreads();
    HB_TRACE(5);
    float acc[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.0f;
    mbar_wait(reinterpret_cast<uint64_t*>(hb_smem + kHbOffBar + 8 * (64 + warp)), 0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 a = lds16(wslot + q * 1024 + lane * 16), b = lds16(wslot + q * 1024 + 512 + lane * 16);
#pragma unroll
      for (int n = 0; n < 4; ++n)
        if (n < ntc) {
          const uint4 bv = *reinterpret_cast<const uint4*>(As + (size_t)(n * 8 + g) * pitch2 + (kh * 4 + q) * 32 + 8 * t);
          mma16816(acc[n], a, bv.x, bv.y);
          mma16816(acc[n], b, bv.z, bv.w);
        }
    }
    // partials -> global [group][source = (K split, K half)][tile][n][lane][4]
    {
      float* dst = p.part2 + (((size_t)rb * 16 + k * 2 + kh) * 4 + t2) * 512 + lane * 4;
#pragma unroll
      for (int n = 0; n < 4; ++n)
        if (n < ntc) __stcg(reinterpret_cast<float4*>(dst + n * 128), make_float4(acc[n][0], acc[n][1], acc[n][2], acc[n][3]));
    }
    HB_TRACE(6);
    hb_group_sync(p.barrier + 33 + rb);
    if (own_cnt) {
      const int nn = tid >> 7, ln = (tid >> 2) & 31, j = tid & 3;
      const int n = 2 * own_np + nn;
      const int m = n * 8 + 2 * (ln & 3) + (j & 1), row = (4 * rb + own_t) * 16 + (ln >> 2) + 8 * (j >> 1);
      if (nn < own_cnt && m < p.M) {
        float s = 0.0f;
#pragma unroll
        for (int src16 = 0; src16 < 16; ++src16) s += __ldcg(p.part2 + (((size_t)rb * 16 + src16) * 4 + own_t) * 512 + n * 128 + ln * 4 + j);
        p.base[(size_t)m * 1024 + row] = s + p.b1[row];
      }
    }
    HB_TRACE(7);
  }
}

}  // namespace tp
