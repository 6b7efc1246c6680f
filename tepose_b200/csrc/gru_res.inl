// Included by gru.cu after gru_tma.inl (same namespace, same helpers).
//
// bf16 recurrence, RESIDENT-WEIGHT variant of k_gru_bf16_tma.  W_hh does not change between steps, so part of
// the CTA's 393 KB slice stays ON CHIP for the whole kernel and only the rest is re-streamed from L2:
//   * chunks [0, rr)            live in REGISTERS as ready-made A fragments (24 registers per chunk per lane),
//   * chunks [rr, rr + rs)      live in a resident SHARED-MEMORY region, filled once by bulk copies,
//   * chunks [rr + rs, nchunks) stream through a small W ring every step (prefetched across the grid barrier).
// h_{t-1} arrives through its own ring (only the NB rows in use are copied).  The h ring is empty whenever the
// gate phase runs, so the cross-warp reduction buffer `red` ALIASES it -- that is what buys the shared memory
// for the resident chunks.  (chunk = 128 columns of the CTA's 96 rows = 24 KB, fragment-packed by tp_pack_whh_bf16.)
//
// Where it pays (measured on B200, profiles/r02_gru_resident_ab.txt): batch <= 8 (live / windowed streams), where
// `red` and the h slices are small, 10 of 16 chunks stay resident and the step gets 7 % shorter.  At batch 32
// the 9-warp CTA caps registers at 168 (3 warps share one SM sub-partition), `red` needs 54 KB and every resident
// chunk costs a ring stage of latency hiding: the streaming kernel stays faster there and remains the default.
constexpr int kResRegChunks4 = 0;     // register-resident chunks at NB = 32: none -- 9 warps put 3 on one SM sub-partition, 168 registers max
constexpr int kResRegChunks1 = 3;     // ... at NB = 8 (92 + 72 registers)
constexpr int kResHStages = 7;           // 7 x 8 KB = 56 KB >= red for NB = 32 (54 KB)

template <int NT, int RR>
__global__ void __launch_bounds__(kTmaThreads, 1) k_gru_bf16_res(const GruParams p, const int rs_max, const int ws) {
  constexpr int NB = NT * 8, U = 32, KG = 4, RP = U + 4;
  constexpr int GE = (NB * U + 255) / 256;
  constexpr int HS = kResHStages;
  constexpr uint32_t kHB = NB * 256;                                  // bytes of an h slice that are read: NB rows x 128 bf16 (= the stage size)
  static_assert((size_t)KG * 3 * NB * RP * 4 <= (size_t)HS * kHB, "red must fit in the h ring it aliases");
  extern __shared__ __align__(1024) unsigned char smem_t[];
  const int H = p.H, B = p.B;
  const int nblk = H / 32, nchunks = nblk / kChunkBlocks;
  const int rr = RR < nchunks ? RR : nchunks;
  const int rs = rs_max < nchunks - rr ? rs_max : nchunks - rr;
  const int nstream = nchunks - rr - rs;
  unsigned char* hring = smem_t;                                      // [HS][kHB]
  float* red = reinterpret_cast<float*>(smem_t);                      // [KG][3][NB][RP], gate phase only
  unsigned char* resw = smem_t + (((size_t)HS * kHB + 1023) & ~(size_t)1023);   // [rs_max][24 KB]
  unsigned char* wring = resw + (size_t)rs_max * kChunkBytes;         // [ws][24 KB]
  uint64_t* h_full = reinterpret_cast<uint64_t*>(wring + (size_t)ws * kChunkBytes);
  uint64_t* h_empty = h_full + 8;
  uint64_t* w_full = h_full + 16;
  uint64_t* w_empty = h_full + 24;
  uint64_t* res_full = h_full + 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3, kg = (warp >> 1) & 3, mg = warp & 1;
  const bool producer = warp == 8;

  if (tid == 0) {
    for (int i = 0; i < HS; ++i) { mb_init(&h_full[i], 1); mb_init(&h_empty[i], 8); }
    for (int i = 0; i < ws; ++i) { mb_init(&w_full[i], 1); mb_init(&w_empty[i], 8); }
    mb_init(res_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __shared__ tp_gru_job sjobs[kMaxJobs];
  for (int i = tid; i < (int)(sizeof(tp_gru_job) * kMaxJobs / 4); i += kTmaThreads)
    reinterpret_cast<int*>(sjobs)[i] = reinterpret_cast<const int*>(p.jobs)[i];
  __syncthreads();
  int j, u0;
  locate_item(p, blockIdx.x, j, u0);
  const tp_gru_job& jb = sjobs[j];
  const unsigned char* wbase = reinterpret_cast<const unsigned char*>(jb.w_hh);
  // byte offset of (gate i, unit half m, chunk c) inside the packed matrix: 4 consecutive 1 KB column blocks
  auto wsrc = [&](int i, int m, int c) {
    return ((((size_t)i * (H / 16) + (u0 >> 4) + m) * nblk) + (size_t)c * kChunkBlocks) * 1024;
  };

  // ---- one-time fills
  uint4 ra[RR > 0 ? RR : 1][3], rb[RR > 0 ? RR : 1][3];               // register-resident A fragments
  if (producer) {
    if (rs > 0) {
      if (lane == 0) mb_expect_tx(res_full, (uint32_t)rs * kChunkBytes);
      __syncwarp();
      if (lane < 6) {
        const int i = lane >> 1, m = lane & 1;
        for (int r = 0; r < rs; ++r)
          bulk_g2s(resw + (size_t)r * kChunkBytes + (size_t)((i * 2 + m) * kChunkBlocks) * 1024, wbase + wsrc(i, m, rr + r),
                   kChunkBlocks * 1024, res_full);
      }
    }
  } else {
#pragma unroll
    for (int r = 0; r < RR; ++r)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (r < rr) {
          const unsigned char* src = wbase + wsrc(i, mg, r) + (size_t)kg * 1024 + (size_t)lane * 16;
          ra[r][i] = ldg_stream(src);
          rb[r][i] = ldg_stream(src + 512);
        } else {
          ra[r][i] = make_uint4(0, 0, 0, 0); rb[r][i] = make_uint4(0, 0, 0, 0);
        }
      }
  }

  uint32_t wprod = 0, hprod = 0;              // producer ring positions
  int prefetched = 0;
  int w_st = 0, h_st = 0; uint32_t w_ph = 0, h_ph = 0;      // consumer ring positions
  bool res_ready = rs == 0;
  const int uu = tid & 31, bb0 = (tid >> 5) & 7;
  float bh_r = 0.f, bh_z = 0.f, bh_n = 0.f;
  if (!producer) { bh_r = __ldg(jb.b_hh + u0 + uu); bh_z = __ldg(jb.b_hh + H + u0 + uu); bh_n = __ldg(jb.b_hh + 2 * H + u0 + uu); }
  auto ldnc = [](const float* ptr) { float v; asm volatile("ld.global.nc.f32 %0, [%1];\n" : "=f"(v) : "l"(ptr)); return v; };
  auto ldcg = [](const float* ptr) { float v; asm volatile("ld.global.cg.f32 %0, [%1];\n" : "=f"(v) : "l"(ptr)); return v; };

  // Producer warp: lane 0 owns all barrier bookkeeping and the h copy; lanes 0..5 issue the six W copies of a stage.
  // Chunks are issued in the order the consumers use them: streamed, then smem-resident, then register-resident.
  auto issue_w = [&](int c) {
    const int st = wprod % ws;
    if (lane == 0) {
      mb_wait(&w_empty[st], ((wprod / ws) & 1) ^ 1);
      mb_expect_tx(&w_full[st], kChunkBytes);
    }
    __syncwarp();
    if (lane < 6) {
      const int i = lane >> 1, m = lane & 1;
      bulk_g2s(wring + (size_t)st * kChunkBytes + (size_t)((i * 2 + m) * kChunkBlocks) * 1024, wbase + wsrc(i, m, c),
               kChunkBlocks * 1024, &w_full[st]);
    }
    ++wprod;
  };
  auto issue_h = [&](int c, const __nv_bfloat16* hprev) {
    const int st = hprod % HS;
    if (lane == 0) {
      mb_wait(&h_empty[st], ((hprod / HS) & 1) ^ 1);
      mb_expect_tx(&h_full[st], kHB);
      bulk_g2s(hring + (size_t)st * kHB, hprev + (size_t)c * 32 * 128, kHB, &h_full[st]);
    }
    ++hprod;
  };
  // k-th chunk in consumption order -> chunk index
  auto chunk_of = [&](int k) { return k < nstream ? rr + rs + k : (k < nstream + rs ? rr + (k - nstream) : k - nstream - rs); };

  // W_hh does not depend on anything the previous kernel writes: the resident fills above and, when step 0 has no
  // matmul (h0 = 0), the W ring for step 1 are issued before the PDL wait, while the input-projection GEMM drains
  if (producer && jb.h0 == nullptr && jb.steps > 1) {
    const int n = ws < nstream ? ws : nstream;
    for (int k = 0; k < n; ++k) issue_w(chunk_of(k));
    prefetched = n;
  }
  // PDL: everything above ran while the input-projection GEMM was draining; gi is read from here on
  pdl_wait();
  pdl_launch_dependents();

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
        for (int k = 0; k < nchunks; ++k) {
          const int c = chunk_of(k);
          if (k < nstream && k >= prefetched) issue_w(c);
          issue_h(c, hprev);
        }
        prefetched = 0;
        if (s + 1 < jb.steps) {               // W_hh is step-invariant: refill the W ring for the next step now
          const int n = ws < nstream ? ws : nstream;
          for (int k = 0; k < n; ++k) issue_w(chunk_of(k));
          prefetched = n;
        }
      }
      __syncwarp();
    } else {
      GateIn gin[GE];
      if (active) {
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
        // one chunk: 3 gates x NT batch tiles x 2 k-steps of m16n8k16
        auto mma_chunk = [&](const uint4 (&wa)[3], const uint4 (&wb)[3], const unsigned char* hb) {
#pragma unroll
          for (int n = 0; n < NT; ++n) {
            const uint4 bv = *reinterpret_cast<const uint4*>(hb + (size_t)(n * 8 + g) * 256 + (size_t)(((kg * 4 + t) ^ ((g & 1) << 2)) << 4));
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              mma_bf16(acc[i][n], wa[i].x, wa[i].y, wa[i].z, wa[i].w, bv.x, bv.y);
              mma_bf16(acc[i][n], wb[i].x, wb[i].y, wb[i].z, wb[i].w, bv.z, bv.w);
            }
          }
        };
        auto h_next = [&]() { if (++h_st == HS) { h_st = 0; h_ph ^= 1; } };
        // (1) streamed chunks: W from the ring
        for (int k = 0; k < nstream; ++k) {
          mb_wait(&w_full[w_st], w_ph);
          const unsigned char* cb = wring + (size_t)w_st * kChunkBytes + (size_t)kg * 1024 + (size_t)lane * 16;
          uint4 wa[3], wb[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            wa[i] = *reinterpret_cast<const uint4*>(cb + (size_t)((i * 2 + mg) * kChunkBlocks) * 1024);
            wb[i] = *reinterpret_cast<const uint4*>(cb + (size_t)((i * 2 + mg) * kChunkBlocks) * 1024 + 512);
          }
          mb_wait(&h_full[h_st], h_ph);
          mma_chunk(wa, wb, hring + (size_t)h_st * kHB);
          __syncwarp();
          if (lane == 0) { mb_arrive(&w_empty[w_st]); mb_arrive(&h_empty[h_st]); }
          if (++w_st == ws) { w_st = 0; w_ph ^= 1; }
          h_next();
        }
        // (2) shared-memory-resident chunks
        if (!res_ready) { mb_wait(res_full, 0); res_ready = true; }
        for (int r = 0; r < rs; ++r) {
          const unsigned char* cb = resw + (size_t)r * kChunkBytes + (size_t)kg * 1024 + (size_t)lane * 16;
          uint4 wa[3], wb[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            wa[i] = *reinterpret_cast<const uint4*>(cb + (size_t)((i * 2 + mg) * kChunkBlocks) * 1024);
            wb[i] = *reinterpret_cast<const uint4*>(cb + (size_t)((i * 2 + mg) * kChunkBlocks) * 1024 + 512);
          }
          mb_wait(&h_full[h_st], h_ph);
          mma_chunk(wa, wb, hring + (size_t)h_st * kHB);
          __syncwarp();
          if (lane == 0) mb_arrive(&h_empty[h_st]);
          h_next();
        }
        // (3) register-resident chunks
#pragma unroll
        for (int r = 0; r < RR; ++r) {
          if (r < rr) {
            mb_wait(&h_full[h_st], h_ph);
            mma_chunk(ra[r], rb[r], hring + (size_t)h_st * kHB);
            __syncwarp();
            if (lane == 0) mb_arrive(&h_empty[h_st]);
            h_next();
          }
        }
        TP_TRACE(2);
        asm volatile("bar.sync 1, 256;\n" ::: "memory");      // every warp is done with the h ring before red overwrites it
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
      // The next step's h copies (async proxy) into the region red aliases are issued only after the grid
      // barrier below, which every consumer reaches after its last read of red.
    }
    if (s + 1 < p.max_steps) {
      TP_TRACE(4);
      grid_barrier_sh(p.barrier, ++epoch, p.barrier_shards);
      TP_TRACE(5);
    }
  }
}
