// IEF iterations on ONE thread-block cluster (bf16 mode, <= 32 rows) -- the tail of K3.
//
// Reference: lib/models/spin.py:250-261 -- 3 x { xc = cat[x, pose, shape, cam]; fc1; fc2; pose += decpose(xc); shape += ...; cam += ... }
// (dropout = identity in eval).  The feature part of fc1 (base = W1[:, :2048].x + b1) does not change between iterations and comes
// from the preceding grid-wide kernel; what is left per iteration is a dependent chain of three small layers
// (K = 160 -> 1024 -> 1024 -> 160) that a grid barrier per layer made 3.4 us each.  Here the 2.6 MB of weights of that chain stay
// ON CHIP for all iterations, spread over a cluster of 16 CTAs, and the layers hand their activations over through distributed
// shared memory with transaction barriers -- no grid barrier, no L2 round trip:
//
//   CTA c owns output rows [64c, 64c+64) of fc1 and fc2 (4 weight tiles of 16 rows) and the K slice [64c, 64c+64) of dec.
//   fc1p : warp = (tile, half of the batch), W1p fragments in registers, B operand = bf16 state PS [32,160] (every CTA has a copy);
//          the CTA's [32,64] slice of u1 goes to all 16 CTAs as 16-byte st.async stores (nearest rank first), and each
//          slice completes ITS OWN barrier U[s] in the receiver, so
//   fc2  : (warp = tile x parity of the arrival order) starts on a slice as soon as it has landed: the all-gather of u1 (64 KB into
//          every SM at the 16 B/clk of distributed shared memory = 4K cycles) runs under the MMAs instead of before them.
//          W2: the first 2 slices of a warp's order live in registers, the other 6 in shared memory (96 KB per CTA);
//          B operand = u1 [32,1024] (slice-major, XOR-swizzled: conflict-free LDS.128).  The result slice u2 [32,64] stays local:
//   dec  : split-K over the cluster -- CTA c multiplies ITS u2 slice with Wdec[:, 64c:64c+64] (20 KB) and sends the fp32 partial sums
//          of output tile j (16 columns) to CTA j (st.async, 2 KB per (source, tile), completing barrier P of the owner).
//   owner: CTAs 0..9 add the 16 partials to the fp32 state they keep in registers, and broadcast the bf16 copy of their 16 columns
//          (1 KB of st.async stores per peer, barrier S); the last iteration writes the fp32 state to global memory instead.
//
// Batches below 25 rows move and multiply only the 8-row groups in use.  The barriers are re-armed per iteration (phase = it & 1).
// Ordering argument for buffer re-use: a CTA sends its u1 slice of iteration i+1 only after S(i) completed, i.e. after every
// owner reduced P(i), i.e. after EVERY CTA finished fc2(i) and dec(i) -- so nobody still reads what the slice overwrites, and
// every barrier has completed phase i before any traffic of phase i+1 can reach it.
#pragma once
#include "skinny.cuh"
#include "umma.cuh"

namespace tp {

constexpr int kClCtas = 16, kClMaxIter = 64, kClThreads = 256;
constexpr uint32_t kClW2Chunk = 12 * 1024;                 // shared-memory part of one (tile, K half) of W2: 12 of 16 k-blocks
constexpr uint32_t kClOffW2 = 0;                           // 8 chunks
constexpr uint32_t kClOffWd = 8 * kClW2Chunk;              // Wdec K slice: 10 tiles x 2 k-blocks x 1 KB
constexpr uint32_t kClOffU1 = kClOffWd + 10 * 2048;        // u1: [16 slices][32 rows][128 B]
constexpr uint32_t kClPsTile = 1088;                       // state tile [32 rows][16 cols] bf16 = 1 KB, +64 B so that tiles 2k, 2k+1 sit in different banks
constexpr uint32_t kClOffPs = kClOffU1 + 65536;            // 10 state tiles
constexpr uint32_t kClOffPart = kClOffPs + 10 * kClPsTile; // owner's inbox: [16 sources][4 n][32 lanes][4] fp32
constexpr uint32_t kClOffStage = kClOffPart + 16 * 2048;   // this CTA's u1 slice: transposes the MMA fragments into the 16-byte chunks that are sent
constexpr uint32_t kClOffBar = kClOffStage + 4096;         // mbarriers: weights, U[16 slices], P, S
constexpr uint32_t kClSmemBytes = kClOffBar + 8 * 19;
static_assert(kClSmemBytes <= 232448, "k_ief_cluster: shared memory");

struct IefClParams {
  const float* base;        // [M,1024] fp32: fc1's feature part + b1 (written by the preceding kernel)
  const uint4* w1p;         // tp_pack_mma_a_bf16 images: [1024 x 160], [1024 x 1024], [160 x 1024]
  const uint4* w2;
  const uint4* wdec;
  const float* b2;
  const float* bdec;
  const float* init; int init_rows;
  float* psc;               // [M,160] out
  int M, n_iter;
  int rows_per_cluster;     // the batch is cut into independent clusters of this many rows (a multiple of 8, <= 32): the hand-offs between
                            // the layers are bound by distributed-shared-memory bandwidth (2 KB per row and CTA), so fewer rows per cluster
                            // is faster as long as SMs are free -- 4 clusters of 8 rows at B = 32
  long long* trace;
};

__device__ __forceinline__ uint32_t cl_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cl_map(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cl_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void cl_bulk_g2s(uint32_t dst_cta, const void* src, uint32_t bytes, uint32_t bar_cta) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
                   "r"(dst_cta), "l"(src), "r"(bytes), "r"(bar_cta) : "memory");
}
__device__ __forceinline__ void cl_st_async_f4(uint32_t dst_cluster, const float* v, uint32_t bar_cluster) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];\n" ::
                   "r"(dst_cluster), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void cl_st_async_u4(uint32_t dst_cluster, const uint4& v, uint32_t bar_cluster) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];\n" ::
                   "r"(dst_cluster), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void cl_wait(uint32_t bar_cta, uint32_t parity) {
  uint32_t done = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(bar_cta), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ uint4 lds16(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
#define CL_TRACE(slot) do { if (trace && tid == 0) trace[16384 + (size_t)c * 64 + (slot)] = clock64(); } while (0)

__global__ void __launch_bounds__(kClThreads, 1) k_ief_cluster(const IefClParams pp) {
  extern __shared__ __align__(128) unsigned char cl_smem[];
  // this cluster's rows of the batch
  IefClParams p = pp;
  {
    const int row0 = (int)(blockIdx.x / kClCtas) * pp.rows_per_cluster;
    p.M = min(pp.rows_per_cluster, pp.M - row0);
    p.base += (size_t)row0 * 1024;
    p.psc += (size_t)row0 * 160;
    if (pp.init_rows != 1) p.init += (size_t)row0 * 160;
  }
  long long* const trace = blockIdx.x < kClCtas ? pp.trace : nullptr;      // (stamps of the first cluster only)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int tile = warp & 3, hh = warp >> 2;      // hh: half of the batch (fc1p) / K half (fc2)
  const uint32_t c = cl_rank();
  const uint32_t sm = smem_u32(cl_smem);
  const uint32_t bar_w = sm + kClOffBar, bar_p = sm + kClOffBar + 8 * 17, bar_s = sm + kClOffBar + 8 * 18;
  auto bar_u = [&](int sl) { return sm + kClOffBar + 8u * (1 + sl); };
  const bool owner = c < 10;
  const int n_iter = p.n_iter;
  const int ntc = (p.M + 7) >> 3;                 // 8-row groups in use (1..4)
  auto arm = [&](uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
  };
  auto arm_iteration = [&](bool last_it) {        // one local arrival + the bytes the peers will send this iteration
    for (int sl = 0; sl < 16; ++sl) arm(bar_u(sl), ntc * 1024);
    if (owner) arm(bar_p, 16 * ntc * 512);
    if (!last_it) arm(bar_s, (owner ? 9 : 10) * ntc * 256);
  };
  // this warp's order of u1 slices in fc2: step i takes the slice of rank (c - (2i + hh)) mod 16, which is the (2i+hh)-th to arrive
  auto slice_of = [&](int i) { return (int)((c - (uint32_t)(2 * i + hh)) & 15u); };

  if (tid == 0) {
    for (int i = 0; i < 19; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(sm + kClOffBar + 8u * i));
    arm(bar_w, 8 * kClW2Chunk + 10 * 2048);
    if (n_iter > 0) arm_iteration(n_iter == 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncwarp();
  if (warp == 0) {
    // resident weights (constants: may be fetched before the predecessor kernel has finished).  W2 chunk = (tile, parity), holding
    // steps 2..7 of that warp's slice order, 2 k-blocks (2 KB, contiguous in the packed matrix) per step
    for (int q = lane; q < 48; q += 32) {
      const int ch = q / 6, i = 2 + q % 6, tl = ch >> 1, h = ch & 1;
      const int sl = (int)((c - (uint32_t)(2 * i + h)) & 15u);
      cl_bulk_g2s(sm + kClOffW2 + ch * kClW2Chunk + (i - 2) * 2048, p.w2 + ((size_t)(4 * c + tl) * 32 + 2 * sl) * 64, 2048, bar_w);
    }
    if (lane < 10) cl_bulk_g2s(sm + kClOffWd + lane * 2048, p.wdec + ((size_t)lane * 32 + 2 * c) * 64, 2048, bar_w);
  }
  // register-resident fragments: all of W1p for this warp's tile, steps 0 and 1 of this warp's slice order of W2
  uint4 w1[5][2], w2r[4][2];
  {
    const uint4* s1 = p.w1p + (size_t)(4 * c + tile) * 5 * 64 + lane;
#pragma unroll
    for (int kb = 0; kb < 5; ++kb) { w1[kb][0] = ldg_stream16(s1 + kb * 64); w1[kb][1] = ldg_stream16(s1 + kb * 64 + 32); }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4* s2 = p.w2 + ((size_t)(4 * c + tile) * 32 + 2 * slice_of(q >> 1) + (q & 1)) * 64 + lane;
      w2r[q][0] = ldg_stream16(s2); w2r[q][1] = ldg_stream16(s2 + 32);
    }
  }
  // state: bf16 copy in every CTA, fp32 master in the owner's registers (value f = tid + 256 r of the owner's [4 n][32 lanes][4] tile)
  for (int i = tid; i < 32 * 160; i += kClThreads) {
    const int b = i / 160, col = i - b * 160;
    const float v = p.init_rows == 1 ? p.init[col] : (b < p.M ? p.init[(size_t)b * 160 + col] : 0.0f);
    *reinterpret_cast<__nv_bfloat16*>(cl_smem + kClOffPs + (col >> 4) * kClPsTile + b * 32 + (col & 15) * 2) = __float2bfloat16_rn(v);
  }
  float state[2] = {0.0f, 0.0f};
  int st_b[2], st_col[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int f = tid + 256 * r, n = f >> 7, ln = (f >> 2) & 31, j = f & 3;
    st_b[r] = n * 8 + 2 * (ln & 3) + (j & 1);
    st_col[r] = (ln >> 2) + 8 * (j >> 1);
    if (owner) state[r] = p.init_rows == 1 ? p.init[c * 16 + st_col[r]] : (st_b[r] < p.M ? p.init[(size_t)st_b[r] * 160 + c * 16 + st_col[r]] : 0.0f);
  }
  const float bd0 = owner ? p.bdec[c * 16 + st_col[0]] : 0.0f, bd1 = owner ? p.bdec[c * 16 + st_col[1]] : 0.0f;
  const float b2lo = p.b2[c * 64 + tile * 16 + g], b2hi = p.b2[c * 64 + tile * 16 + g + 8];

  cl_sync();                                    // every CTA's barriers are armed before any remote traffic; PS written
  pdl_wait();                                   // base comes from the preceding kernel
  pdl_launch_dependents();
  float basev[2][4];
#pragma unroll
  for (int nn = 0; nn < 2; ++nn)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int b = (2 * hh + nn) * 8 + 2 * t + (j & 1), u = c * 64 + tile * 16 + g + 8 * (j >> 1);
      basev[nn][j] = b < p.M ? __ldcg(p.base + (size_t)b * 1024 + u) : 0.0f;
    }
  CL_TRACE(0);

  // the bf16 slice layout of u1 / u2: element (row b, column u of 64) -> b*128 + ((u>>3) ^ ((b&1)<<2))*16 + (u&7)*2
  auto slice_off = [](int b, int u) { return (uint32_t)(b * 128 + (((u >> 3) ^ ((b & 1) << 2)) << 4) + (u & 7) * 2); };

  for (int it = 0; it < n_iter; ++it) {
    const uint32_t par = it & 1;
    const bool last = it + 1 == n_iter;
    if (it > 0) {
      cl_wait(bar_s, par ^ 1);                   // the state of the previous iteration, from its 10 owners
      if (tid == 0) arm_iteration(last);         // phase it-1 of every barrier is complete everywhere (header): arm phase it
    }
    CL_TRACE(1 + it * 8);
    // ---- fc1p: u1 = base + W1p . state^T     (warp: tile x half of the batch, K = 160)
    {
      float a1[2][4];
#pragma unroll
      for (int nn = 0; nn < 2; ++nn)
#pragma unroll
        for (int j = 0; j < 4; ++j) a1[nn][j] = basev[nn][j];
#pragma unroll
      for (int kb = 0; kb < 5; ++kb)
#pragma unroll
        for (int nn = 0; nn < 2; ++nn)
          if (2 * hh + nn < ntc) {
            const int row = (2 * hh + nn) * 8 + g;
            const uint4 bv = lds16(sm + kClOffPs + (2 * kb + (t >> 1)) * kClPsTile + row * 32 + (t & 1) * 16);
            mma16816(a1[nn], w1[kb][0], bv.x, bv.y);
            mma16816(a1[nn], w1[kb][1], bv.z, bv.w);
          }
      unsigned char* mine = cl_smem + kClOffStage;
#pragma unroll
      for (int nn = 0; nn < 2; ++nn)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int b = (2 * hh + nn) * 8 + 2 * t + (j & 1), u = tile * 16 + g + 8 * (j >> 1);
          *reinterpret_cast<__nv_bfloat16*>(mine + slice_off(b, u)) = __float2bfloat16_rn(a1[nn][j]);
        }
      __syncthreads();
      // the slice leaves as 16-byte st.async stores straight from registers, one per (thread, CTA of the cluster): the bulk-copy
      // engine of an SM works through its copies one after the other (~280 cycles each, 16 B/clk), the store path does not queue
      if (tid < ntc * 64) {
        const uint4 v = lds16(sm + kClOffStage + tid * 16);
#pragma unroll 4
        for (int kk = 0; kk < kClCtas; ++kk) {
          const uint32_t peer = (c + kk) & 15;   // nearest rank first: the order in which the receivers consume
          cl_st_async_u4(cl_map(sm + kClOffU1 + c * 4096 + tid * 16, peer), v, cl_map(bar_u(c), peer));
        }
      }
    }
    CL_TRACE(2 + it * 8);
    if (it == 0) cl_wait(bar_w, 0);
    // ---- fc2: u2 slice = W2[64c:64c+64, :] . u1^T + b2     (warp: tile x parity of the slice order; a slice is used as it lands)
    {
      float a2[4][4], a2b[4][4];                 // two accumulator sets (the k-blocks of a slice alternate): half the dependent-MMA chain
#pragma unroll
      for (int n = 0; n < 4; ++n) a2[n][0] = a2[n][1] = a2[n][2] = a2[n][3] = a2b[n][0] = a2b[n][1] = a2b[n][2] = a2b[n][3] = 0.0f;
      auto step = [&](int sl, int kk, const uint4& wa, const uint4& wb) {
        const uint32_t rowbase = sm + kClOffU1 + sl * 4096 + g * 128 + (((kk * 4 + t) ^ ((g & 1) << 2)) << 4);
#pragma unroll
        for (int n = 0; n < 4; ++n)
          if (n < ntc) {
            const uint4 bv = lds16(rowbase + n * 1024);
            mma16816(kk ? a2b[n] : a2[n], wa, bv.x, bv.y);
            mma16816(kk ? a2b[n] : a2[n], wb, bv.z, bv.w);
          }
      };
      // the warp's 8 slices in two groups of 4 (nearest ranks first): ONE round of barrier waits per group, then four steps whose
      // shared-memory loads the scheduler may overlap freely (a wait per slice made every step a ~460-cycle dependent round)
      const uint32_t wsm = sm + kClOffW2 + (tile * 2 + hh) * kClW2Chunk + lane * 16;
#pragma unroll
      for (int grp = 0; grp < 2; ++grp) {
#pragma unroll
        for (int i = 4 * grp; i < 4 * grp + 4; ++i) cl_wait(bar_u(slice_of(i)), par);
        if (grp == 0) CL_TRACE(3 + it * 8);
#pragma unroll
        for (int i = 4 * grp; i < 4 * grp + 4; ++i) {
          const int sl = slice_of(i);
          if (i < 2) {
            step(sl, 0, w2r[2 * i][0], w2r[2 * i][1]);
            step(sl, 1, w2r[2 * i + 1][0], w2r[2 * i + 1][1]);
          } else {
            const uint4 wa0 = lds16(wsm + (i - 2) * 2048), wb0 = lds16(wsm + (i - 2) * 2048 + 512);
            const uint4 wa1 = lds16(wsm + (i - 2) * 2048 + 1024), wb1 = lds16(wsm + (i - 2) * 2048 + 1536);
            step(sl, 0, wa0, wb0);
            step(sl, 1, wa1, wb1);
          }
        }
      }
#pragma unroll
      for (int n = 0; n < 4; ++n) { a2[n][0] += a2b[n][0]; a2[n][1] += a2b[n][1]; a2[n][2] += a2b[n][2]; a2[n][3] += a2b[n][3]; }
      __syncthreads();                            // u1 is consumed: its space takes the parity exchange and the u2 slice
      float* red = reinterpret_cast<float*>(cl_smem + kClOffU1);                 // [4 tiles][4 n][32 lanes][4]
      unsigned char* u2s = cl_smem + kClOffU1 + 8192;
      if (hh == 1) {
#pragma unroll
        for (int n = 0; n < 4; ++n)
          *reinterpret_cast<float4*>(red + ((tile * 4 + n) * 32 + lane) * 4) = make_float4(a2[n][0], a2[n][1], a2[n][2], a2[n][3]);
      }
      __syncthreads();
      if (hh == 0) {
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const float4 o = *reinterpret_cast<const float4*>(red + ((tile * 4 + n) * 32 + lane) * 4);
          const float v[4] = {a2[n][0] + o.x + b2lo, a2[n][1] + o.y + b2lo, a2[n][2] + o.z + b2hi, a2[n][3] + o.w + b2hi};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int b = n * 8 + 2 * t + (j & 1), u = tile * 16 + g + 8 * (j >> 1);
            *reinterpret_cast<__nv_bfloat16*>(u2s + slice_off(b, u)) = __float2bfloat16_rn(v[j]);
          }
        }
      }
      __syncthreads();
      CL_TRACE(4 + it * 8);
      // ---- dec, this CTA's K slice: partial[32, 160] = u2 slice . Wdec[:, 64c:64c+64]^T, tile tt goes to CTA tt
      const uint32_t u2a = smem_u32(u2s);
      for (int tt = warp; tt < 10; tt += 8) {
        float a3[4][4];
#pragma unroll
        for (int n = 0; n < 4; ++n) a3[n][0] = a3[n][1] = a3[n][2] = a3[n][3] = 0.0f;
#pragma unroll
        for (int k2 = 0; k2 < 2; ++k2) {
          const uint4 wa = lds16(sm + kClOffWd + tt * 2048 + k2 * 1024 + lane * 16);
          const uint4 wb = lds16(sm + kClOffWd + tt * 2048 + k2 * 1024 + 512 + lane * 16);
          const uint32_t rowbase = u2a + g * 128 + (((k2 * 4 + t) ^ ((g & 1) << 2)) << 4);
#pragma unroll
          for (int n = 0; n < 4; ++n)
            if (n < ntc) {
              const uint4 bv = lds16(rowbase + n * 1024);
              mma16816(a3[n], wa, bv.x, bv.y);
              mma16816(a3[n], wb, bv.z, bv.w);
            }
        }
        const uint32_t dst = cl_map(sm + kClOffPart + c * 2048 + lane * 16, tt), bar = cl_map(bar_p, tt);
#pragma unroll
        for (int n = 0; n < 4; ++n)
          if (n < ntc) cl_st_async_f4(dst + n * 512, a3[n], bar);
      }
    }
    CL_TRACE(5 + it * 8);
    // ---- owners: state += sum of the 16 partials + bdec; broadcast the bf16 copy (or write the result)
    if (owner) {
      cl_wait(bar_p, par);
      CL_TRACE(6 + it * 8);
      const float* part = reinterpret_cast<const float*>(cl_smem + kClOffPart);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (((tid + 256 * r) >> 7) < ntc) {
          float s = 0.0f;
#pragma unroll
          for (int src = 0; src < 16; ++src) s += part[src * 512 + tid + 256 * r];
          state[r] += s + (r == 0 ? bd0 : bd1);
          if (last) {
            if (st_b[r] < p.M) p.psc[(size_t)st_b[r] * 160 + c * 16 + st_col[r]] = state[r];
          } else {
            *reinterpret_cast<__nv_bfloat16*>(cl_smem + kClOffPs + c * kClPsTile + st_b[r] * 32 + st_col[r] * 2) = __float2bfloat16_rn(state[r]);
          }
        }
      }
      if (!last) {
        __syncthreads();
        if (tid < ntc * 16) {                    // the tile's [8 ntc rows][32 B] as 16-byte chunks, to the 15 other CTAs
          const uint4 v = lds16(sm + kClOffPs + c * kClPsTile + tid * 16);
#pragma unroll 5
          for (int kk = 1; kk < kClCtas; ++kk) {
            const uint32_t peer = (c + kk) & 15;
            cl_st_async_u4(cl_map(sm + kClOffPs + c * kClPsTile + tid * 16, peer), v, cl_map(bar_s, peer));
          }
        }
      }
      CL_TRACE(7 + it * 8);
    }
  }
  if (n_iter == 0) {
    cl_wait(bar_w, 0);                           // the weight copies must have landed before the CTA may exit
    if (owner) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
        if (st_b[r] < p.M) p.psc[(size_t)st_b[r] * 160 + c * 16 + st_col[r]] = state[r];
    }
  }
  CL_TRACE(63);
  cl_sync();                                     // no CTA leaves while a peer may still write into its shared memory
}

}  // namespace tp
