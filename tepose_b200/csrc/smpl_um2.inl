// Included by smpl.cu after smpl_um.inl (namespace tp; shares its helpers).  Large-batch SMPL path, second generation:
// BOTH contractions of linear blend skinning run on tcgen05, nothing per (vertex, body) goes through the shared-memory load pipe.
//
//   D1[vertex 128, body 16] x 3 planes = Blend_plane[128, K = 240] . Coef[16, 240]^T                (blended rest-pose vertex p)
//   D2[vertex 128, (body 8, e 12)]     = Wskin[128, 24 joints] . Transform[(body, e), 24 joints]^T  (T_v = sum_j w_vj A_j, 3 x 4)
//
// The first-generation kernel (smpl_um.inl) gathered the four 48-byte joint transforms of every (vertex, body) from shared
// memory: 192 B x 6890 x 65,536 = 87 GB through a 128 B/clk/SM pipe = 2.3 ms before anything else (ncu: 85 shared wavefronts per
// 32 vertex-bodies, LSU 63 % busy, 7.7 ms).  Here the skinning weights are a dense [128 x 24] bf16 operand per vertex tile and the
// joint transforms of 8 bodies a [96 x 24] operand; fp32-grade accuracy comes from the 3-term split W = W_hi + W_lo, A = A_hi + A_lo:
// D2 = W_hi.A_hi + W_lo.A_hi + W_hi.A_lo (6 MMAs of K = 16; the lo.lo term is 2^-18 relative).  The epilogue reads p (3 values)
// and T (12 values) of its (vertex, body) with tcgen05.ld and does 12 FMAs.
//
// TMEM (all 512 columns): [0,48) D1, [48,144) + [144,240) D2 of the two 8-body skinning steps of a group (so neither step waits
// for the epilogue to drain the other: with a single D2 every step cost a full MMA -> epilogue -> MMA round trip, 2.3 ms),
// [240,272) the tile's skinning-weight operand, [272,512) the x and y planes of the tile's blend rows (K = 240 bf16 = 120 columns
// each).  The z plane (64 KB) stays in shared memory as a tcgen05 operand image: TMEM has no room for it next to the second D2.
// Shared memory: ring of {coefficient rows of 16 bodies (8 KB) + transform operands of 2 x 8 bodies (24 KB)} written by
// k_smpl_prepare as tcgen05 operand images and moved by ONE bulk copy each; the z plane; per-warp slabs for the coalesced
// vertex store and the 3xTF32 joint-regressor MMAs (as in the first generation).
// Warp roles: 16 epilogue warps (TMEM lane quadrant = warp % 4; body pair = warp / 4 of every 8-body skinning step), a producer
// warp, three MMA-issuing warps.
constexpr int kU2Threads = 640;                   // 16 epilogue warps, producer warp, three MMA-issuing warps
constexpr int kU2EpiWarps = 16;
constexpr int kU2GB = 16;                          // bodies per blend MMA (N) and per ring stage
constexpr int kU2SB = 8;                           // bodies per skinning MMA (N = 96)
constexpr int kU2Stages = 3;
constexpr int kU2CoefBytes = kU2GB * 256 * 2;       // 8 KB: [K block 4][row 16][128 B]
constexpr int kU2TimgBytes = kU2SB * 12 * 128;      // 12 KB: [row (body, e) 96][128 B] = A_hi (k 0..23) | pad | A_lo (k 32..55) | pad
constexpr int kU2StageBytes = kU2CoefBytes + 2 * kU2TimgBytes;      // 32 KB
constexpr int kU2WBytes = 128 * 128;                // skinning-weight operand of a tile: [row 128][128 B] = W_hi (k 0..23) | pad | W_lo | pad
constexpr int kU2SlabPitch = 100;
constexpr int kU2SlabFloats = 4 * kU2SlabPitch;     // 4 bodies per warp and group
constexpr int kU2ChunkGroups = 256;                 // 4096 bodies per L2-resident chunk
constexpr size_t kU2OffRing = 0;
constexpr size_t kU2ZBytes = 4 * 16384;             // z plane of the tile's blend rows: [K block 4][row 128][128 B]
constexpr size_t kU2OffZ = kU2OffRing + (size_t)kU2Stages * kU2StageBytes;
constexpr size_t kU2OffSlab = kU2OffZ + kU2ZBytes;
constexpr size_t kU2OffFrag = kU2OffSlab + (size_t)kU2EpiWarps * kU2SlabFloats * 4;
constexpr size_t kU2OffBar = kU2OffFrag + kUsFragBytes;
constexpr size_t kU2Smem = kU2OffBar + 256 + 1024;
constexpr uint32_t kU2IdescBlend = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kU2GB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t kU2IdescSkin = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((kU2SB * 12) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t kU2ColD1 = 0, kU2ColD2 = 48, kU2ColW = 240, kU2ColA = 272;
static_assert(kU2Stages * kU2StageBytes >= 4 * 16384, "the TMEM-fill staging area (4 x 16 KB) aliases the ring");

__device__ __forceinline__ void us_tmem_st(uint32_t taddr, const uint32_t* v, int n) {     // n in {32, 16, 8}
  if (n == 32) { us_tmem_st32(taddr, v); return; }
  if (n == 16) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                 "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
    return;
  }
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void us_tmem_ld2(uint32_t taddr, float& a, float& b) {
  uint32_t r0, r1;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(taddr));
  a = __uint_as_float(r0); b = __uint_as_float(r1);
}

// Busy-polling wait for the TMEM hand-offs between the MMA warp and the epilogue warps: mbarrier.try_wait may suspend the thread
// for a system-defined time, and the wake-up showed as the dominant cost of every D2 round trip (two per 16 bodies).
__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}

struct U2Params {
  tp_smpl_model m;
  int n, ngroups, ntiles, nreg;
  const unsigned char* coef_img;   // [ngroups][8 KB]
  const unsigned char* timg;       // [ngroups][2][12 KB]
  const float* jreg;               // [nreg][vp]
  float* verts;                    // [n][n_verts][3] or null
  float* jpart;                    // [n][ntiles][nreg][3]
};

__global__ void __launch_bounds__(kU2Threads, 1) k_smpl_lbs_um2(const U2Params p) {
  extern __shared__ __align__(1024) unsigned char u2_raw[];
  unsigned char* smem = u2_raw + ((1024u - (smem_u32(u2_raw) & 1023u)) & 1023u);
  unsigned char* s_ring = smem + kU2OffRing;           // doubles as the TMEM-fill staging area between tiles
  unsigned char* s_z = smem + kU2OffZ;
  float* s_slab = reinterpret_cast<float*>(smem + kU2OffSlab);
  uint32_t* s_frag = reinterpret_cast<uint32_t*>(smem + kU2OffFrag);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kU2OffBar);
  uint64_t* in_full = bars;                // [3]
  uint64_t* in_empty = bars + 3;           // [3]
  uint64_t* d1_full = bars + 6;
  uint64_t* d1_empty = bars + 7;
  uint64_t* d2_full = bars + 8;            // [2]
  uint64_t* d2_empty = bars + 10;          // [2]
  uint64_t* f_full = bars + 12;            // [4]
  uint64_t* z_full = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < kU2Stages; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], 3); }
    mbar_init(d1_full, 2); mbar_init(d1_empty, kU2EpiWarps);
    for (int i = 0; i < 2; ++i) { mbar_init(&d2_full[i], 1); mbar_init(&d2_empty[i], kU2EpiWarps); }
    for (int i = 0; i < 4; ++i) mbar_init(&f_full[i], 1);
    mbar_init(z_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d1 = tmem_base + kU2ColD1, tmem_d2 = tmem_base + kU2ColD2, tmem_w = tmem_base + kU2ColW, tmem_a = tmem_base + kU2ColA;

  pdl_wait();                       // the operand images come from k_smpl_prepare
  pdl_launch_dependents();

  uint32_t ngrp = 0, nfill = 0;     // groups / tile set-ups this CTA has been through (every role counts in step)
  const int q = warp & 3, sset = warp >> 2;
  for (int chunk0 = 0; chunk0 < p.ngroups; chunk0 += kU2ChunkGroups) {
  const int cgroups = p.ngroups - chunk0 < kU2ChunkGroups ? p.ngroups - chunk0 : kU2ChunkGroups;
  const long long total = (long long)p.ntiles * cgroups;
  long long it = total * blockIdx.x / gridDim.x;
  const long long it_hi = total * (blockIdx.x + 1) / gridDim.x;
  while (it < it_hi) {
    const int tile = (int)(it / cgroups);
    const int gl0 = (int)(it - (long long)tile * cgroups);
    const long long left = it_hi - it;
    const int gl1 = (long long)(cgroups - gl0) < left ? cgroups : gl0 + (int)left;
    const int g0 = chunk0 + gl0, g1 = chunk0 + gl1;
    const int v0 = tile * kUsVT;

    // ------------------------------------------------------------------ tile set-up (everyone; the pipeline is drained)
    __syncthreads();
    {
      const unsigned char* img = reinterpret_cast<const unsigned char*>(p.m.blend_um) + (size_t)tile * kUsTileImage;
      // staged blocks 0..7: x and y planes (plane = blk / 4, K block = blk % 4), block 8: the skinning-weight operand
      const unsigned char* wimg = reinterpret_cast<const unsigned char*>(p.m.skin_um) + (size_t)tile * kU2WBytes;
      auto blk_src = [&](int blk) { return blk < 8 ? img + (size_t)blk * 16384 : wimg; };
      if (warp == 16 && elect_one()) {
        for (int i = 0; i < 4; ++i) { mbar_expect_tx(&f_full[i], 16384); us_bulk_g2s(s_ring + i * 16384, blk_src(i), 16384, &f_full[i]); }
        mbar_expect_tx(z_full, (uint32_t)kU2ZBytes);
        us_bulk_g2s(s_z, img + (size_t)8 * 16384, (uint32_t)kU2ZBytes, z_full);             // z plane: straight into place
      }
      for (int i = tid; i < 4 * 4 * 32 * 4; i += kU2Threads) {          // regressor A fragments (m16n8k8, tf32 hi | lo)
        const int e = i & 3, ln = (i >> 2) & 31, ks = (i >> 7) & 3, qq = i >> 9;
        const int row = (ln >> 2) + ((e & 1) ? 8 : 0), v = qq * 32 + ks * 8 + (ln & 3) + ((e & 2) ? 4 : 0);
        const float val = row < p.nreg ? __ldg(p.jreg + (size_t)row * p.m.vp + v0 + v) : 0.0f;
        const uint32_t hi = __float_as_uint(val) & 0xffffe000u;
        const float lo = val - __uint_as_float(hi);
        uint32_t* dst = s_frag + ((size_t)(qq * 4 + ks) * 32 + ln) * 8;
        dst[e] = hi; dst[4 + e] = __float_as_uint(lo);
      }
      for (int blk = 0; blk < 9; ++blk) {
        const int st = blk & 3;
        if (warp < 4) {
          // stage 0 is used three times per tile set-up (blocks 0, 4, 8), the others twice
          mbar_wait(&f_full[st], ((st == 0 ? 3u : 2u) * nfill + (uint32_t)(blk >> 2)) & 1u);
          const int row = warp * 32 + lane;
          const unsigned char* rowp = s_ring + (size_t)st * 16384 + (size_t)row * 128;
          uint32_t v[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 x = *reinterpret_cast<const uint4*>(rowp + ((c ^ (row & 7)) << 4));
            v[4 * c] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
          }
          const uint32_t lanes = (uint32_t)(warp * 32) << 16;
          if (blk == 8) us_tmem_st(tmem_w + lanes, v, 32);
          else {
            const uint32_t dst = tmem_a + (uint32_t)((blk >> 2) * 120 + (blk & 3) * 32) + lanes;     // K blocks 0..2: 32 columns, block 3: 24 (K = 240)
            if ((blk & 3) < 3) us_tmem_st(dst, v, 32);
            else { us_tmem_st(dst, v, 16); us_tmem_st(dst + 16, v + 16, 8); }
          }
        }
        __syncthreads();
        if (warp == 16 && blk + 4 < 9 && elect_one()) {
          mbar_expect_tx(&f_full[st], 16384);
          us_bulk_g2s(s_ring + st * 16384, blk_src(blk + 4), 16384, &f_full[st]);
        }
      }
      if (warp < 4) asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
      mbar_wait(z_full, nfill & 1u);
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      ++nfill;
    }

    if (warp == 16) {
      // ================================================================ producer: one stage = the operands of 16 bodies
      for (int g = g0; g < g1; ++g, ++ngrp) {
        const uint32_t st = ngrp % kU2Stages, ph = (ngrp / kU2Stages) & 1u;
        mbar_wait(&in_empty[st], ph ^ 1u);
        if (elect_one()) {
          unsigned char* dst = s_ring + (size_t)st * kU2StageBytes;
          mbar_expect_tx(&in_full[st], kU2StageBytes);
          us_bulk_g2s(dst, p.coef_img + (size_t)g * kU2CoefBytes, kU2CoefBytes, &in_full[st]);
          us_bulk_g2s(dst + kU2CoefBytes, p.timg + (size_t)g * 2 * kU2TimgBytes, 2 * kU2TimgBytes, &in_full[st]);
        }
        __syncwarp();
      }
    } else if (warp >= 17) {
      // ================================================================ MMA issuers: three warps, three independent streams.
      // A tcgen05.mma with N = 16 occupies the tensor pipe ~17 cycles but costs its issuing thread ~40 (ncu: the single MMA warp
      // of the first version spent 81 % of its time stalled on UTCHMMA while the pipe was 35 % busy), and the accumulators of
      // the three streams are disjoint: warp 17 = blend planes x, y (A in TMEM), warp 18 = blend plane z (A in shared memory),
      // warp 19 = the two skinning steps.  D1 is complete when both blend warps have committed; a ring stage is free when all
      // three have.
      const uint32_t role = (uint32_t)warp - 17u;
      const uint64_t dz = umma_desc_sw128(smem_u32(s_z));
      for (int g = g0; g < g1; ++g, ++ngrp) {
        const uint32_t st = ngrp % kU2Stages, ph = (ngrp / kU2Stages) & 1u;
        mbar_wait(&in_full[st], ph);
        const uint64_t dc = umma_desc_sw128(smem_u32(s_ring + (size_t)st * kU2StageBytes));
        if (role < 2) {
          mbar_spin(d1_empty, (ngrp & 1u) ^ 1u);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          if (elect_one()) {
            if (role == 0) {
#pragma unroll
              for (int pl = 0; pl < 2; ++pl)
#pragma unroll
                for (int i = 0; i < 15; ++i)
                  us_umma_ts(tmem_d1 + (uint32_t)(pl * 16), tmem_a + (uint32_t)(pl * 120 + i * 8),
                             dc + (uint64_t)((i >> 2) * (2048 >> 4) + (i & 3) * 2), kU2IdescBlend, i == 0 ? 0u : 1u);
            } else {
#pragma unroll
              for (int i = 0; i < 15; ++i)
                umma_f16(tmem_d1 + 32u, dz + (uint64_t)((i >> 2) * (16384 >> 4) + (i & 3) * 2),
                         dc + (uint64_t)((i >> 2) * (2048 >> 4) + (i & 3) * 2), kU2IdescBlend, i == 0 ? 0u : 1u);
            }
            umma_commit(d1_full);
            umma_commit(&in_empty[st]);
          }
          __syncwarp();
        } else {
          const uint64_t dt = dc + (uint64_t)(kU2CoefBytes >> 4);
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            mbar_spin(&d2_empty[h], (ngrp & 1u) ^ 1u);          // buffer h was drained by the epilogue of the PREVIOUS group: no round trip on the critical path
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            if (elect_one()) {
              const uint64_t db = dt + (uint64_t)(h * (kU2TimgBytes >> 4));
              const uint32_t d2 = tmem_d2 + (uint32_t)(h * 96);
              // (W_hi, A_hi), (W_lo, A_hi), (W_hi, A_lo): two K = 16 steps each.  W in TMEM: hi = columns 0..15, lo = 16..31 (8 columns
              // per K step); A rows in shared memory: the lo half sits 64 bytes into the 128-byte rows
              us_umma_ts(d2, tmem_w + 0, db + 0, kU2IdescSkin, 0u);
              us_umma_ts(d2, tmem_w + 8, db + 2, kU2IdescSkin, 1u);
              us_umma_ts(d2, tmem_w + 16, db + 0, kU2IdescSkin, 1u);
              us_umma_ts(d2, tmem_w + 24, db + 2, kU2IdescSkin, 1u);
              us_umma_ts(d2, tmem_w + 0, db + 4, kU2IdescSkin, 1u);
              us_umma_ts(d2, tmem_w + 8, db + 6, kU2IdescSkin, 1u);
              umma_commit(&d2_full[h]);
              if (h == 1) umma_commit(&in_empty[st]);
            }
            __syncwarp();
          }
        }
      }
    } else {
      // ================================================================ epilogue
      const int v = v0 + q * 32 + lane;
      const float tx = __ldg(p.m.template_pad + (size_t)v * 3), ty = __ldg(p.m.template_pad + (size_t)v * 3 + 1),
                  tz = __ldg(p.m.template_pad + (size_t)v * 3 + 2);
      const bool vvalid = v < p.m.n_verts;
      float* slab = s_slab + (size_t)warp * kU2SlabFloats;
      const uint32_t* frag = s_frag + (size_t)q * 4 * 32 * 8 + lane * 8;
      const uint32_t t_lane = (uint32_t)(q * 32) << 16;
      const int64_t nv3 = (int64_t)p.m.n_verts * 3;
      const int gq = lane >> 2, tq = lane & 3;
      const int64_t vbase = (int64_t)(v0 + q * 32) * 3;
      const int64_t vrem = nv3 - vbase;
      const bool st_a2 = lane * 2 + 1 < vrem, st_a1 = lane * 2 < vrem;
      const bool st_b2 = lane < 16 && 64 + lane * 2 + 1 < vrem, st_b1 = lane < 16 && 64 + lane * 2 < vrem;
      const bool tile_full = vrem >= 96;
      // slot bi (0..3) of this warp = body (bi >> 1) * 8 + 2 sset + (bi & 1) of the group
      // partial-sum slots this thread adds up: value i = row * 16 + column, column = slot * 3 + coordinate (12 used of 16)
      int js_off[2]; bool js_ok[2]; int js_body[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int i = q * 32 + lane + 128 * k;
        const int r = i >> 4, col = i & 15, bi = col / 3, c = col - bi * 3;
        const int bl = (bi >> 1) * 8 + 2 * sset + (bi & 1);
        js_ok[k] = r < p.nreg && col < 12;
        js_body[k] = bl;
        js_off[k] = ((bl * p.ntiles + tile) * p.nreg + r) * 3 + c;
      }
      for (int g = g0; g < g1; ++g, ++ngrp) {
        float px[4], py[4], pz[4];
        mbar_spin(d1_full, ngrp & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          us_tmem_ld2(tmem_d1 + t_lane + (uint32_t)(0 + 8 * h + 2 * sset), px[2 * h], px[2 * h + 1]);
          us_tmem_ld2(tmem_d1 + t_lane + (uint32_t)(16 + 8 * h + 2 * sset), py[2 * h], py[2 * h + 1]);
          us_tmem_ld2(tmem_d1 + t_lane + (uint32_t)(32 + 8 * h + 2 * sset), pz[2 * h], pz[2 * h + 1]);
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) us_mb_arrive(d1_empty);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float T[24];
          mbar_spin(&d2_full[h], ngrp & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          us_tmem_ld8(tmem_d2 + t_lane + (uint32_t)(h * 96 + sset * 24), T);
          us_tmem_ld8(tmem_d2 + t_lane + (uint32_t)(h * 96 + sset * 24 + 8), T + 8);
          us_tmem_ld8(tmem_d2 + t_lane + (uint32_t)(h * 96 + sset * 24 + 16), T + 16);
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
          __syncwarp();
          if (lane == 0) us_mb_arrive(&d2_empty[h]);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int bi = 2 * h + i;
            const float* Tb = T + 12 * i;
            const float x = px[bi] + tx, y = py[bi] + ty, z = pz[bi] + tz;
            const float ox = fmaf(Tb[0], x, fmaf(Tb[1], y, fmaf(Tb[2], z, Tb[3])));
            const float oy = fmaf(Tb[4], x, fmaf(Tb[5], y, fmaf(Tb[6], z, Tb[7])));
            const float oz = fmaf(Tb[8], x, fmaf(Tb[9], y, fmaf(Tb[10], z, Tb[11])));
            float* o = slab + bi * kU2SlabPitch + lane * 3;
            o[0] = vvalid ? ox : 0.0f; o[1] = vvalid ? oy : 0.0f; o[2] = vvalid ? oz : 0.0f;
          }
        }
        __syncwarp();
        const int gbody0 = g * kU2GB;
        // coalesced vertex store: 384 contiguous bytes per (warp, body); one 64-bit address product per group, the ragged last
        // vertex tile and the ragged last body group take the predicated path
        if (p.verts) {
          float* d0 = p.verts + (int64_t)(gbody0 + 2 * sset) * nv3 + vbase + lane * 2;
          const bool full_grp = gbody0 + kU2GB <= p.n;
#pragma unroll
          for (int bi = 0; bi < 4; ++bi) {
            float* dstb = d0 + ((bi >> 1) ? nv3 * 8 : 0) + ((bi & 1) ? nv3 : 0);
            const float2* src = reinterpret_cast<const float2*>(slab + bi * kU2SlabPitch);
            if (full_grp && tile_full) {
              *reinterpret_cast<float2*>(dstb) = src[lane];
              if (lane < 16) *reinterpret_cast<float2*>(dstb + 64) = src[32 + lane];
            } else if (gbody0 + (bi >> 1) * 8 + 2 * sset + (bi & 1) < p.n) {
              if (st_a2) *reinterpret_cast<float2*>(dstb) = src[lane];
              else if (st_a1) dstb[0] = src[lane].x;
              if (st_b2) *reinterpret_cast<float2*>(dstb + 64) = src[32 + lane];
              else if (st_b1) dstb[64] = src[32 + lane].x;
            }
          }
        }
        // joint regressors: D[row 16, (slot, coord) 12 of 16] += J[16, 32 v] . V[32 v, 16], 3xTF32 mma.sync
        if (p.nreg > 0) {
          float d[2][4];
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) d[nt][0] = d[nt][1] = d[nt][2] = d[nt][3] = 0.0f;
          int boff[2]; bool bok[2];
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            const int col = nt * 8 + gq;
            bok[nt] = col < 12;
            const int cc = bok[nt] ? col : 0;
            boff[nt] = (cc / 3) * kU2SlabPitch + (cc % 3) + tq * 3;
          }
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint4 fh = *reinterpret_cast<const uint4*>(frag + ks * 32 * 8);
            const uint4 fl = *reinterpret_cast<const uint4*>(frag + ks * 32 * 8 + 4);
            const uint32_t ah[4] = {fh.x, fh.y, fh.z, fh.w}, al[4] = {fl.x, fl.y, fl.z, fl.w};
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
              float b0 = slab[boff[nt] + ks * 24], b1 = slab[boff[nt] + ks * 24 + 12];
              if (!bok[nt]) { b0 = 0.0f; b1 = 0.0f; }
              const uint32_t b0h = __float_as_uint(b0) & 0xffffe000u, b1h = __float_as_uint(b1) & 0xffffe000u;
              const uint32_t b0l = __float_as_uint(b0 - __uint_as_float(b0h)), b1l = __float_as_uint(b1 - __uint_as_float(b1h));
              us_mma_tf32(d[nt], al, b0h, b1h);
              us_mma_tf32(d[nt], ah, b0l, b1l);
              us_mma_tf32(d[nt], ah, b0h, b1h);
            }
          }
          __syncwarp();                                      // the slab is free: it now carries this warp's partial sums [16 rows][16]
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            *reinterpret_cast<float2*>(slab + gq * 16 + nt * 8 + tq * 2) = make_float2(d[nt][0], d[nt][1]);
            *reinterpret_cast<float2*>(slab + (gq + 8) * 16 + nt * 8 + tq * 2) = make_float2(d[nt][2], d[nt][3]);
          }
          asm volatile("bar.sync %0, 128;\n" ::"r"(1 + sset) : "memory");
          {
            const float* s0 = s_slab + (size_t)(sset * 4) * kU2SlabFloats + q * 32 + lane;
            float* jp = p.jpart + (int64_t)gbody0 * p.ntiles * p.nreg * 3;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              if (js_ok[k] && gbody0 + js_body[k] < p.n)
                jp[js_off[k]] = (s0[128 * k] + s0[kU2SlabFloats + 128 * k]) + (s0[2 * kU2SlabFloats + 128 * k] + s0[3 * kU2SlabFloats + 128 * k]);
            }
          }
          asm volatile("bar.sync %0, 128;\n" ::"r"(1 + sset) : "memory");
        } else {
          __syncwarp();
        }
      }
    }
    it += g1 - g0;
  }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 16) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512));
  }
}
