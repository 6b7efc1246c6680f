// Included by gru.cu after gru_tma.inl (same namespace, same helpers).
//
// bf16 recurrence for EXACTLY TWO independent matmul jobs (the two causal directions of TePose's encoder), interleaved.
// k_gru_bf16_tma gives each direction its own 64 CTAs and runs both in lock-step: every step, a CTA streams its
// operands (~13K cycles at the L2 -> SM ceiling) and then sits through ~6.5K cycles of reduction, gate math and grid
// barrier with the L2 idle.  Here every CTA owns 16 hidden units of BOTH directions, each direction is a self-contained
// team of 5 warps (4 K-group consumers + 1 TMA producer) with its own operand ring, reduction buffer and grid-barrier
// counter, and the two teams are never synchronised with each other: while one sits in its gate / barrier phase the
// other one has the L2 stream to itself.  Price: a CTA pulls h_{t-1} of both directions (2 x 128 KB instead of 128 KB
// per step next to the 393 KB of W_hh).
//
// Ring stage of a team: 12 KB W (3 gate tiles x 4 column blocks, fragment-packed) + 8 KB h slice (tiled, swizzled).
// The h parts are empty whenever the team is in its gate phase (h_t does not exist yet), so the cross-warp reduction
// buffer lives there (2048 floats per stage) and all of shared memory goes to ring depth: 5 stages per team.
constexpr int kDualThreads = 320;                         // 2 teams x (4 consumer warps + 1 producer warp)
constexpr int kDualWBytes = 3 * kChunkBlocks * 1024;      // 12 KB
constexpr int kDualStageBytes = kDualWBytes + kHChunkBytes;   // 20 KB
#ifndef TP_DUAL_STAGES
#define TP_DUAL_STAGES 5
#endif
constexpr int kDualStages = TP_DUAL_STAGES;
constexpr int kDualRedFloats = 4 * 3 * 32 * 20;           // [KG][3][NB = 32][RP = 16 + 4]
static_assert(kDualStages >= 4 && 3 * 32 * 20 <= kHChunkBytes / 4, "K group k's reduction slab lives in the h part of ring stage k");
constexpr size_t kDualSmem = (size_t)2 * kDualStages * kDualStageBytes + 512;

__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(count) : "memory"); }

template <int NT>
__global__ void __launch_bounds__(kDualThreads, 1) k_gru_bf16_dual(const GruParams p) {
  constexpr int NB = NT * 8, U = 16, KG = 4, RP = U + 4, GE = NT;        // NB x 16 gate elements per team / 128 threads
  constexpr uint32_t kHB = NB * 256;                                      // bytes of an h slice that are read (NB rows x 128 bf16)
  extern __shared__ __align__(1024) unsigned char smem_d[];
  const int H = p.H, B = p.B;
  const int nblk = p.K / 32, nchunks = nblk / kChunkBlocks;       // K = H, or 3 H for the [hi | lo | hi] operands of TP_PRECISION_BF16X3
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  // team d = direction / job d: consumer warps 4d..4d+3, producer warp 8+d
  const bool producer = warp >= 8;
  const int d = producer ? warp - 8 : warp >> 2;
  const int kg = warp & 3;                                  // consumers: K group
  const int ctid = tid - d * 128;                           // consumers: thread index inside the team (0..127)
  unsigned char* ring = smem_d + (size_t)d * kDualStages * kDualStageBytes;
  // reduction buffer: K group k's slab [3][NB][RP] (1920 floats) lives in the h part (2048 floats) of ring stage k
  auto redk = [&](int k) { return reinterpret_cast<float*>(ring + (size_t)k * kDualStageBytes + kDualWBytes); };
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_d + (size_t)2 * kDualStages * kDualStageBytes);
  uint64_t* full = bars + d * 16;                           // [stages]: W arrival + h arrival
  uint64_t* empty = full + 8;                               // [stages]: 4 consumer warps

  if (tid == 0) {
    for (int q = 0; q < 2; ++q)
      for (int i = 0; i < kDualStages; ++i) { mb_init(&bars[q * 16 + i], 2); mb_init(&bars[q * 16 + 8 + i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __shared__ tp_gru_job sjobs[kMaxJobs];
  for (int i = tid; i < (int)(sizeof(tp_gru_job) * kMaxJobs / 4); i += kDualThreads)
    reinterpret_cast<int*>(sjobs)[i] = reinterpret_cast<const int*>(p.jobs)[i];
  __syncthreads();

  const tp_gru_job& jb = sjobs[d];
  const int u0 = blockIdx.x * U;                            // this CTA's 16 units (of both directions)
  const unsigned char* wbase = reinterpret_cast<const unsigned char*>(jb.w_hh);
  auto wsrc = [&](int i, int c) { return ((((size_t)i * (H / 16) + blockIdx.x) * nblk) + (size_t)c * kChunkBlocks) * 1024; };
  uint32_t prod = 0, hprod = 0;
  int prefetched = 0;
  int c_stage = 0; uint32_t c_phase = 0;
  auto issue_w = [&](int c) {                               // producer: lanes 0..2 copy the three gate tiles of chunk c
    const int st = prod % kDualStages;
    if (lane == 0) {
      mb_wait(&empty[st], ((prod / kDualStages) & 1) ^ 1);
      mb_expect_tx(&full[st], kDualWBytes);
    }
    __syncwarp();
    if (lane < 3)
      bulk_g2s(ring + (size_t)st * kDualStageBytes + (size_t)(lane * kChunkBlocks) * 1024, wbase + wsrc(lane, c), kChunkBlocks * 1024, &full[st]);
    ++prod;
  };
  auto issue_h = [&](int c, const __nv_bfloat16* hprev) {
    const int st = hprod % kDualStages;
    if (lane == 3) {
      mb_expect_tx(&full[st], kHB);
      bulk_g2s(ring + (size_t)st * kDualStageBytes + kDualWBytes, hprev + (size_t)c * 32 * 128, kHB, &full[st]);
    }
    ++hprod;
  };
  // W_hh does not depend on the previous kernel: fill the ring for step 1 before the PDL wait
  if (producer && jb.steps > 1) {
    const int n = kDualStages < nchunks ? kDualStages : nchunks;
    for (int c = 0; c < n; ++c) issue_w(c);
    prefetched = n;
  }
  pdl_wait();
  pdl_launch_dependents();

  // step-0-only jobs without an initial state have no matmul at all: plain gate math, grid-strided
  for (int je = p.n_item_jobs; je < p.njobs; ++je)
    for (int64_t i = blockIdx.x * (int64_t)kDualThreads + tid; i < (int64_t)B * H; i += (int64_t)gridDim.x * kDualThreads) {
      const int b = (int)(i / H), u = (int)(i - (int64_t)b * H);
      if (p.split3) gru_finalize<false>(p, sjobs[je], je, 0, b, u, gate_fetch(p, sjobs[je], je, 0, b, u), 0.f, 0.f, 0.f);
      else gru_finalize<true>(p, sjobs[je], je, 0, b, u, gate_fetch(p, sjobs[je], je, 0, b, u), 0.f, 0.f, 0.f);
    }

  // from here on the two teams never meet: named barriers 1 + d (team, 160 threads) and 3 + d (consumers, 128 threads)
  unsigned int* counter = p.barrier + 32 * d;               // one grid-barrier counter per direction, 128 bytes apart
  unsigned int epoch = 0;
  const int uu = ctid & 15, bb0 = ctid >> 4;                // consumers: gate element (batch bb0 + 8e, unit uu)
  float bh_r = 0.f, bh_z = 0.f, bh_n = 0.f;
  if (!producer) { bh_r = __ldg(jb.b_hh + u0 + uu); bh_z = __ldg(jb.b_hh + H + u0 + uu); bh_n = __ldg(jb.b_hh + 2 * H + u0 + uu); }
  auto ldnc = [](const float* ptr) { float v; asm volatile("ld.global.nc.f32 %0, [%1];\n" : "=f"(v) : "l"(ptr)); return v; };
  auto ldcg = [](const float* ptr) { float v; asm volatile("ld.global.cg.f32 %0, [%1];\n" : "=f"(v) : "l"(ptr)); return v; };

  if (d == 1 && p.dual_skew > 0) {                          // optional phase offset between the teams
    const long long t0 = clock64();
    while (clock64() - t0 < p.dual_skew) { }
  }
  for (int s = 0; s < jb.steps; ++s) {
    TP_TRACE(0);
    const bool have_prev = s > 0;                           // jobs with h0 do not take this kernel
    if (producer) {
      if (have_prev) {
        asm volatile("fence.proxy.async;\n" ::: "memory");
        const __nv_bfloat16* hprev = p.hbuf_lp + (int64_t)(d * 2 + ((s + 1) & 1)) * p.lp_slot;
        for (int c = 0; c < prefetched; ++c) issue_h(c, hprev);
        for (int c = prefetched; c < nchunks; ++c) { issue_w(c); issue_h(c, hprev); }
        prefetched = 0;
        if (s + 1 < jb.steps) {
          const int n = kDualStages < nchunks ? kDualStages : nchunks;
          for (int c = 0; c < n; ++c) issue_w(c);
          prefetched = n;
        }
      }
      __syncwarp();
    } else {
      GateIn gin[GE];
      {
        const int t_in = jb.t_in0 + s * jb.t_in_step;
        const float* g0 = jb.gi + ((int64_t)t_in * B + bb0) * jb.ldg + (u0 + uu);
        const float* h0 = p.hbuf + ((int64_t)(d * 2 + ((s + 1) & 1)) * B + bb0) * H + (u0 + uu);
        const int64_t gstride = (int64_t)8 * jb.ldg, hstride = (int64_t)8 * H;
#pragma unroll
        for (int e = 0; e < GE; ++e) {
          gin[e].br = bh_r; gin[e].bz = bh_z; gin[e].bn = bh_n; gin[e].hp = 0.0f;
          gin[e].gr = gin[e].gz = gin[e].gn = 0.0f;
        }
#pragma unroll
        for (int e = 0; e < GE; ++e) {
          if (bb0 + 8 * e < B) {
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
          mb_wait(&full[st], c_phase);
          const unsigned char* cb = ring + (size_t)st * kDualStageBytes + (size_t)kg * 1024 + (size_t)lane * 16;
          const unsigned char* hb = ring + (size_t)st * kDualStageBytes + kDualWBytes;      // [32 rows][128] bf16, swizzled
          uint4 wa[3], wb[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            wa[i] = *reinterpret_cast<const uint4*>(cb + (size_t)(i * kChunkBlocks) * 1024);
            wb[i] = *reinterpret_cast<const uint4*>(cb + (size_t)(i * kChunkBlocks) * 1024 + 512);
          }
#pragma unroll
          for (int n = 0; n < NT; ++n) {
            const uint4 bv = *reinterpret_cast<const uint4*>(hb + (size_t)(n * 8 + g) * 256 + (size_t)(((kg * 4 + t) ^ ((g & 1) << 2)) << 4));
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              mma_bf16(acc[i][n], wa[i].x, wa[i].y, wa[i].z, wa[i].w, bv.x, bv.y);
              mma_bf16(acc[i][n], wb[i].x, wb[i].y, wb[i].z, wb[i].w, bv.z, bv.w);
            }
          }
          __syncwarp();
          if (lane == 0) mb_arrive(&empty[st]);
          if (++c_stage == kDualStages) { c_stage = 0; c_phase ^= 1; }
        }
        TP_TRACE(2);
        bar_sync(3 + d, 128);                                // every warp of the team is done with the h slices
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int n = 0; n < NT; ++n) {
            float* r0 = redk(kg) + (size_t)(i * NB + n * 8 + 2 * t) * RP + g;
            r0[0] = acc[i][n][0];
            r0[RP] = acc[i][n][1];
            r0[8] = acc[i][n][2];
            r0[RP + 8] = acc[i][n][3];
          }
        bar_sync(3 + d, 128);
        TP_TRACE(3);
      }
#pragma unroll
      for (int e = 0; e < GE; ++e) {
        const int bb = bb0 + 8 * e;
        if (bb >= B) continue;
        float ar = 0.f, az = 0.f, an = 0.f;
        if (have_prev) {
#pragma unroll
          for (int k = 0; k < KG; ++k) {
            const float* rk = redk(k);
            ar += rk[(0 * NB + bb) * RP + uu];
            az += rk[(1 * NB + bb) * RP + uu];
            an += rk[(2 * NB + bb) * RP + uu];
          }
        }
        if (p.split3) gru_finalize<false>(p, jb, d, s, bb, u0 + uu, gin[e], ar, az, an);     // fp32-grade mode: exact expf / tanhf
        else gru_finalize<true>(p, jb, d, s, bb, u0 + uu, gin[e], ar, az, an);
      }
    }
    if (s + 1 < jb.steps) {
      // grid barrier of this direction only: all CTAs' teams d.  Same release / relaxed-poll protocol as grid_barrier().
      TP_TRACE(4);
      bar_sync(1 + d, 160);
      if (!producer && ctid == 0) {
        const unsigned int target = ++epoch * gridDim.x;
        unsigned int seen;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(counter) : "memory");
        do {
          asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];\n" : "=r"(seen) : "l"(counter) : "memory");
        } while (seen < target);
#ifdef TP_BARRIER_ACQUIRE_FENCE
        asm volatile("fence.acq_rel.gpu;\n" ::: "memory");     // the formal acquire (see common.cuh grid_barrier); off by default
#endif
      }
      bar_sync(1 + d, 160);
      TP_TRACE(5);
    }
  }
}
