// Included by gru.cu (inside namespace tp, after GruParams / gate_fetch / gru_finalize / grid_barrier).
//
// bf16 recurrence, TMA ring variant.  Same decomposition as k_gru_bf16<NT, 2> (U = 32 units per
// CTA = 6 m-tiles, 4 K groups x 2 unit tiles), but the operands arrive through the async proxy:
// a ninth warp issues cp.async.bulk copies that
//   * stream the CTA's fragment-packed W_hh slice through a ring of stages (24 KB = 6 m-tiles x 4 column
//     blocks) guarded by full / empty mbarriers, and
//   * attach to every stage the matching 8 KB slice of h_{t-1} (bf16; kept chunk-tiled and XOR-swizzled in
//     global memory by the gate epilogue, so a slice is ONE contiguous bulk copy and the B-fragment reads
//     are bank-conflict free without padding) -- there is no separate h buffer and no staging phase.
// W_hh does not depend on h, so the producer runs ahead: while the consumers do the gate math and
// wait at the grid barrier, the ring is already refilled with the next step's first chunks -- the
// L2 -> SM stream overlaps the phases that used to leave the L2 idle (measured: the recurrence is
// bound by L2 bandwidth, ~7 TB/s for 128 streaming CTAs, see scripts/micro/l2bw.cu).
constexpr int kTmaThreads = 288;
constexpr int kChunkBlocks = 4;
constexpr int kChunkBytes = 6 * kChunkBlocks * 1024;      // W_hh part of a ring stage: 24 KB
constexpr int kHChunkBytes = 32 * 128 * 2;                // h part: 32 rows x 128 columns bf16 = 8 KB (tiled + swizzled in global)
constexpr int kStageBytes = kChunkBytes + kHChunkBytes;

__device__ __forceinline__ uint32_t sm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(sm_u32(b)), "r"(n));
}
__device__ __forceinline__ void mb_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(sm_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(sm_u32(b)) : "memory");
}
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t addr = sm_u32(b);
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
                   "r"(sm_u32(dst)), "l"(src), "r"(bytes), "r"(sm_u32(bar)) : "memory");
}

template <int NT>
__global__ void __launch_bounds__(kTmaThreads, 1) k_gru_bf16_tma(const GruParams p, const int stages) {
  constexpr int NB = NT * 8, U = 32, KG = 4, RP = U + 4;
  constexpr int GE = (NB * U + 255) / 256;
  extern __shared__ __align__(1024) unsigned char smem_t[];
  const int H = p.H, B = p.B;
  const int nblk = H / 32, nchunks = nblk / kChunkBlocks;
  const size_t region = ((size_t)KG * 3 * NB * RP * 4 + 1023) & ~(size_t)1023;
  float* red = reinterpret_cast<float*>(smem_t);                      // [KG][3][NB][RP]
  unsigned char* ring = smem_t + region;                              // [stages][24 KB W | 8 KB h]
  uint64_t* w_full = reinterpret_cast<uint64_t*>(ring + (size_t)stages * kStageBytes);
  uint64_t* w_empty = w_full + 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3, kg = (warp >> 1) & 3, mg = warp & 1;
  const bool producer = warp == 8;

  if (tid == 0) {
    for (int i = 0; i < stages; ++i) { mb_init(&w_full[i], 2); mb_init(&w_empty[i], 8); }   // full: W arrival + h arrival
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  // job table -> shared memory (dynamic indexing of the kernel-parameter copy costs a constant-cache
  // round trip per field access; measured ~2.5K cycles per use site)
  __shared__ tp_gru_job sjobs[kMaxJobs];
  for (int i = tid; i < (int)(sizeof(tp_gru_job) * kMaxJobs / 4); i += kTmaThreads)
    reinterpret_cast<int*>(sjobs)[i] = reinterpret_cast<const int*>(p.jobs)[i];
  __syncthreads();
  int j, u0;
  locate_item(p, blockIdx.x, j, u0);          // exactly one item per CTA on this path
  const tp_gru_job& jb = sjobs[j];
  const unsigned char* wbase = reinterpret_cast<const unsigned char*>(jb.w_hh);
  uint32_t prod = 0, hprod = 0;               // producer ring positions: W parts issued, h parts issued
  int prefetched = 0;
  int c_stage = 0; uint32_t c_phase = 0;      // consumer ring position (no runtime division in the hot loop)
  // gate-math operands that never change: this thread always owns unit u0 + (tid & 31) and batches
  // (tid >> 5) + 8e, so b_hh is loaded once per kernel and gi / h_prev are one base pointer + strides
  const int uu = tid & 31, bb0 = (tid >> 5) & 7;
  float bh_r = 0.f, bh_z = 0.f, bh_n = 0.f;
  if (!producer) { bh_r = __ldg(jb.b_hh + u0 + uu); bh_z = __ldg(jb.b_hh + H + u0 + uu); bh_n = __ldg(jb.b_hh + 2 * H + u0 + uu); }
  auto ldnc = [](const float* ptr) { float v; asm volatile("ld.global.nc.f32 %0, [%1];\n" : "=f"(v) : "l"(ptr)); return v; };
  auto ldcg = [](const float* ptr) { float v; asm volatile("ld.global.cg.f32 %0, [%1];\n" : "=f"(v) : "l"(ptr)); return v; };

  // Producer warp: lane 0 owns the barrier bookkeeping, lanes 0..5 each issue one of the six W copies of a
  // stage (cp.async.bulk issue is ~50 cycles a piece), lane 6 the h slice.
  auto issue_w = [&](int c) {
    const int st = prod % stages;
    if (lane == 0) {
      mb_wait(&w_empty[st], ((prod / stages) & 1) ^ 1);
      mb_expect_tx(&w_full[st], kChunkBytes);
    }
    __syncwarp();
    if (lane < 6) {
      const int i = lane >> 1, m = lane & 1;
      const size_t src = ((((size_t)i * (H / 16) + (u0 >> 4) + m) * nblk) + (size_t)c * kChunkBlocks) * 1024;
      bulk_g2s(ring + (size_t)st * kStageBytes + (size_t)((i * 2 + m) * kChunkBlocks) * 1024, wbase + src,
               kChunkBlocks * 1024, &w_full[st]);
    }
    ++prod;
  };
  auto issue_h = [&](int c, const __nv_bfloat16* hprev) {     // stage order == W order: hprod trails prod
    const int st = hprod % stages;
    if (lane == 6) {
      mb_expect_tx(&w_full[st], kHChunkBytes);
      bulk_g2s(ring + (size_t)st * kStageBytes + kChunkBytes, hprev + (size_t)c * 32 * 128, kHChunkBytes, &w_full[st]);
    }
    ++hprod;
  };

  // W_hh does not depend on anything the previous kernel writes: when step 0 has no matmul (h0 = 0) the producer fills
  // the ring for step 1 right away -- before the PDL wait, i.e. while the input-projection GEMM is still draining
  if (producer && jb.h0 == nullptr && jb.steps > 1) {
    const int n = stages < nchunks ? stages : nchunks;
    for (int c = 0; c < n; ++c) issue_w(c);
    prefetched = n;
  }
  // PDL: everything above ran while the input-projection GEMM was draining; gi is read from here on
  pdl_wait();
  pdl_launch_dependents();

  // step-0-only jobs without an initial state have no matmul at all: plain gate math, grid-strided
  for (int je = p.n_item_jobs; je < p.njobs; ++je)
    for (int64_t i = blockIdx.x * (int64_t)kTmaThreads + tid; i < (int64_t)B * H; i += (int64_t)gridDim.x * kTmaThreads) {
      const int b = (int)(i / H), u = (int)(i - (int64_t)b * H);
      gru_finalize<true>(p, sjobs[je], je, 0, b, u, gate_fetch(p, sjobs[je], je, 0, b, u), 0.f, 0.f, 0.f);
    }

  unsigned int epoch = 0;
  if (p.any_h0) { seed_h0(p); grid_barrier_sh(p.barrier, ++epoch, p.barrier_shards); }

  for (int s = 0; s < p.max_steps; ++s) {
    TP_TRACE(0);
    const bool active = s < jb.steps;
    const bool have_prev = active && ((s > 0) || (jb.h0 != nullptr));
    if (producer) {
      if (have_prev) {
        asm volatile("fence.proxy.async;\n" ::: "memory");
        const __nv_bfloat16* hprev = p.hbuf_lp + (size_t)(blockIdx.x % kHRep) * p.lp_rep_stride +
                                     (int64_t)(j * 2 + ((s + 1) & 1)) * p.lp_slot;
        // h slices for the stages whose W part was prefetched across the barrier, then the rest in lock-step
        for (int c = 0; c < prefetched; ++c) issue_h(c, hprev);
        for (int c = prefetched; c < nchunks; ++c) { issue_w(c); issue_h(c, hprev); }
        prefetched = 0;
        if (s + 1 < jb.steps) {               // W_hh is step-invariant: refill the ring for the next step now
          const int n = stages < nchunks ? stages : nchunks;
          for (int c = 0; c < n; ++c) issue_w(c);
          prefetched = n;
        }
      }
      __syncwarp();
    } else {
      GateIn gin[GE];
      if (active) {
        // all address arithmetic first (one base + strides), then the loads back to back: no address
        // computation may wait on a register that is the destination of a load still in flight
        const int t_in = jb.t_in0 + s * jb.t_in_step;
        const float* g0 = jb.gi + ((int64_t)t_in * B + bb0) * jb.ldg + (u0 + uu);
        const float* h0 = p.hbuf + ((int64_t)(j * 2 + ((s + 1) & 1)) * B + bb0) * H + (u0 + uu);
        const int64_t gstride = (int64_t)8 * jb.ldg, hstride = (int64_t)8 * H;
#pragma unroll
        for (int e = 0; e < GE; ++e) {
          gin[e].br = bh_r; gin[e].bz = bh_z; gin[e].bn = bh_n; gin[e].hp = 0.0f;
          gin[e].gr = gin[e].gz = gin[e].gn = 0.0f;
        }
#pragma unroll
        for (int e = 0; e < GE; ++e) {
          if (bb0 + 8 * e < B && (NT > 1 || e == 0)) {
            gin[e].gr = ldnc(g0 + e * gstride);
            gin[e].gz = ldnc(g0 + e * gstride + H);
            gin[e].gn = ldnc(g0 + e * gstride + 2 * H);
            if (have_prev) gin[e].hp = ldcg(h0 + e * hstride);
          }
        }
      }
      if (have_prev) {
        float acc[3][NT][4];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[i][n][e] = 0.0f;
        TP_TRACE(1);
        for (int c = 0; c < nchunks; ++c) {
          const int st = c_stage;
          mb_wait(&w_full[st], c_phase);
          const unsigned char* cb = ring + (size_t)st * kStageBytes + (size_t)kg * 1024 + (size_t)lane * 16;
          const unsigned char* hb = ring + (size_t)st * kStageBytes + kChunkBytes;      // [32 rows][128] bf16, swizzled
          uint4 wa[3], wb[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            wa[i] = *reinterpret_cast<const uint4*>(cb + (size_t)((i * 2 + mg) * kChunkBlocks) * 1024);
            wb[i] = *reinterpret_cast<const uint4*>(cb + (size_t)((i * 2 + mg) * kChunkBlocks) * 1024 + 512);
          }
#pragma unroll
          for (int n = 0; n < NT; ++n) {
            const uint4 bv = *reinterpret_cast<const uint4*>(hb + (size_t)(n * 8 + g) * 256 + (size_t)(((kg * 4 + t) ^ ((g & 1) << 2)) << 4));   // rows g, g+1 of a quarter-warp hit different bank halves
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              mma_bf16(acc[i][n], wa[i].x, wa[i].y, wa[i].z, wa[i].w, bv.x, bv.y);
              mma_bf16(acc[i][n], wb[i].x, wb[i].y, wb[i].z, wb[i].w, bv.z, bv.w);
            }
          }
          __syncwarp();
          if (lane == 0) mb_arrive(&w_empty[st]);
          if (++c_stage == stages) { c_stage = 0; c_phase ^= 1; }
        }
        TP_TRACE(2);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int n = 0; n < NT; ++n) {
            float* r0 = red + ((size_t)(kg * 3 + i) * NB + n * 8 + 2 * t) * RP + mg * 16 + g;
            r0[0] = acc[i][n][0];
            r0[RP] = acc[i][n][1];
            r0[8] = acc[i][n][2];
            r0[RP + 8] = acc[i][n][3];
          }
        asm volatile("bar.sync 1, 256;\n" ::: "memory");
        TP_TRACE(3);
      }
      if (active) {
#pragma unroll
        for (int e = 0; e < GE; ++e) {
          const int idx = tid + e * 256;
          const int bb = idx / U, uu = idx - bb * U;
          if (idx >= NB * U || bb >= B) continue;
          float ar = 0.f, az = 0.f, an = 0.f;
          if (have_prev) {
#pragma unroll
            for (int k = 0; k < KG; ++k) {
              ar += red[((size_t)(k * 3 + 0) * NB + bb) * RP + uu];
              az += red[((size_t)(k * 3 + 1) * NB + bb) * RP + uu];
              an += red[((size_t)(k * 3 + 2) * NB + bb) * RP + uu];
            }
          }
          gru_finalize<true>(p, jb, j, s, bb, u0 + uu, gin[e], ar, az, an);
        }
      }
      // h_t in GLOBAL memory is ordered by the grid barrier's release and by the producer's own
      // fence.proxy.async before it issues the bulk reads; red is never touched by the async proxy.
    }
    if (s + 1 < p.max_steps) {
      TP_TRACE(4);
      grid_barrier_sh(p.barrier, ++epoch, p.barrier_shards);
      TP_TRACE(5);
    }
  }
}
