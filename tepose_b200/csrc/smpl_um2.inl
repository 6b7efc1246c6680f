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
// TMEM (512 columns): [0,48) D1, [48,144) D2, [144,504) the vertex tile's blend rows (3 planes x K = 240 bf16 = 120 columns each).
// Shared memory: ring of {coefficient rows of 16 bodies (8 KB) + transform operands of 2 x 8 bodies (24 KB)} written by
// k_smpl_prepare as tcgen05 operand images and moved by ONE bulk copy each; the tile's skinning-weight operand (16 KB); per-warp
// slabs for the coalesced vertex store and the 3xTF32 joint-regressor MMAs (as in the first generation).
// Warp roles: 16 epilogue warps (TMEM lane quadrant = warp % 4; body pair = warp / 4 of every 8-body skinning step), a producer
// warp, an MMA warp.
constexpr int kU2Threads = 576;
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
constexpr size_t kU2OffW = kU2OffRing + (size_t)kU2Stages * kU2StageBytes;
constexpr size_t kU2OffSlab = kU2OffW + kU2WBytes;
constexpr size_t kU2OffFrag = kU2OffSlab + (size_t)kU2EpiWarps * kU2SlabFloats * 4;
constexpr size_t kU2OffBar = kU2OffFrag + kUsFragBytes;
constexpr size_t kU2Smem = kU2OffBar + 256 + 1024;
constexpr uint32_t kU2IdescBlend = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kU2GB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t kU2IdescSkin = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((kU2SB * 12) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t kU2ColD1 = 0, kU2ColD2 = 48, kU2ColA = 144;
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
  unsigned char* s_w = smem + kU2OffW;
  float* s_slab = reinterpret_cast<float*>(smem + kU2OffSlab);
  uint32_t* s_frag = reinterpret_cast<uint32_t*>(smem + kU2OffFrag);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kU2OffBar);
  uint64_t* in_full = bars;                // [3]
  uint64_t* in_empty = bars + 3;           // [3]
  uint64_t* d1_full = bars + 6;
  uint64_t* d1_empty = bars + 7;
  uint64_t* d2_full = bars + 8;
  uint64_t* d2_empty = bars + 9;
  uint64_t* f_full = bars + 10;            // [4]
  uint64_t* w_full = bars + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < kU2Stages; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], 1); }
    mbar_init(d1_full, 1); mbar_init(d1_empty, kU2EpiWarps);
    mbar_init(d2_full, 1); mbar_init(d2_empty, kU2EpiWarps);
    for (int i = 0; i < 4; ++i) mbar_init(&f_full[i], 1);
    mbar_init(w_full, 1);
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
  const uint32_t tmem_d1 = tmem_base + kU2ColD1, tmem_d2 = tmem_base + kU2ColD2, tmem_a = tmem_base + kU2ColA;

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
      if (warp == 16 && elect_one()) {
        for (int i = 0; i < 4; ++i) { mbar_expect_tx(&f_full[i], 16384); us_bulk_g2s(s_ring + i * 16384, img + (size_t)i * 16384, 16384, &f_full[i]); }
        mbar_expect_tx(w_full, kU2WBytes);
        us_bulk_g2s(s_w, reinterpret_cast<const unsigned char*>(p.m.skin_um) + (size_t)tile * kU2WBytes, kU2WBytes, w_full);
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
      for (int blk = 0; blk < 12; ++blk) {                              // block = plane * 4 + K block; K blocks 0..2: 32 columns, block 3: 24 (K = 240)
        const int st = blk & 3;
        if (warp < 4) {
          mbar_wait(&f_full[st], (nfill * 3 + (uint32_t)(blk >> 2)) & 1u);
          const int row = warp * 32 + lane;
          const unsigned char* rowp = s_ring + (size_t)st * 16384 + (size_t)row * 128;
          uint32_t v[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 x = *reinterpret_cast<const uint4*>(rowp + ((c ^ (row & 7)) << 4));
            v[4 * c] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
          }
          const uint32_t dst = tmem_a + (uint32_t)((blk >> 2) * 120 + (blk & 3) * 32) + ((uint32_t)(warp * 32) << 16);
          if ((blk & 3) < 3) us_tmem_st(dst, v, 32);
          else { us_tmem_st(dst, v, 16); us_tmem_st(dst + 16, v + 16, 8); }
        }
        __syncthreads();
        if (warp == 16 && blk + 4 < 12 && elect_one()) {
          mbar_expect_tx(&f_full[st], 16384);
          us_bulk_g2s(s_ring + st * 16384, img + (size_t)(blk + 4) * 16384, 16384, &f_full[st]);
        }
      }
      if (warp < 4) asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
      mbar_wait(w_full, nfill & 1u);
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
    } else if (warp == 17) {
      // ================================================================ MMA issuer
      const uint64_t dw = umma_desc_sw128(smem_u32(s_w));
      for (int g = g0; g < g1; ++g, ++ngrp) {
        const uint32_t st = ngrp % kU2Stages, ph = (ngrp / kU2Stages) & 1u;
        mbar_wait(&in_full[st], ph);
        mbar_wait(d1_empty, (ngrp & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint64_t dc = umma_desc_sw128(smem_u32(s_ring + (size_t)st * kU2StageBytes));
        const uint64_t dt = umma_desc_sw128(smem_u32(s_ring + (size_t)st * kU2StageBytes + kU2CoefBytes));
        if (elect_one()) {
#pragma unroll
          for (int pl = 0; pl < 3; ++pl)
#pragma unroll
            for (int i = 0; i < 15; ++i)
              us_umma_ts(tmem_d1 + (uint32_t)(pl * 16), tmem_a + (uint32_t)(pl * 120 + i * 8),
                         dc + (uint64_t)((i >> 2) * (2048 >> 4) + (i & 3) * 2), kU2IdescBlend, i == 0 ? 0u : 1u);
          umma_commit(d1_full);
        }
        __syncwarp();
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          mbar_wait(d2_empty, ((2u * ngrp + (uint32_t)h) & 1u) ^ 1u);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          if (elect_one()) {
            const uint64_t db = dt + (uint64_t)(h * (kU2TimgBytes >> 4));
            // (W_hi, A_hi), (W_lo, A_hi), (W_hi, A_lo): two K = 16 steps each; the lo halves sit 64 bytes into the 128-byte rows
            umma_f16(tmem_d2, dw + 0, db + 0, kU2IdescSkin, 0u);
            umma_f16(tmem_d2, dw + 2, db + 2, kU2IdescSkin, 1u);
            umma_f16(tmem_d2, dw + 4, db + 0, kU2IdescSkin, 1u);
            umma_f16(tmem_d2, dw + 6, db + 2, kU2IdescSkin, 1u);
            umma_f16(tmem_d2, dw + 0, db + 4, kU2IdescSkin, 1u);
            umma_f16(tmem_d2, dw + 2, db + 6, kU2IdescSkin, 1u);
            umma_commit(d2_full);
            if (h == 1) umma_commit(&in_empty[st]);
          }
          __syncwarp();
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
        mbar_wait(d1_full, ngrp & 1u);
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
          mbar_wait(d2_full, (2u * ngrp + (uint32_t)h) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          us_tmem_ld8(tmem_d2 + t_lane + (uint32_t)(sset * 24), T);
          us_tmem_ld8(tmem_d2 + t_lane + (uint32_t)(sset * 24 + 8), T + 8);
          us_tmem_ld8(tmem_d2 + t_lane + (uint32_t)(sset * 24 + 16), T + 16);
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
          __syncwarp();
          if (lane == 0) us_mb_arrive(d2_empty);
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
        // coalesced vertex store: 384 contiguous bytes per (warp, body)
        if (p.verts) {
#pragma unroll
          for (int bi = 0; bi < 4; ++bi) {
            const int body = gbody0 + (bi >> 1) * 8 + 2 * sset + (bi & 1);
            if (body < p.n) {
              float* dstb = p.verts + (int64_t)body * nv3 + vbase;
              const float2* src = reinterpret_cast<const float2*>(slab + bi * kU2SlabPitch);
              if (st_a2) *reinterpret_cast<float2*>(dstb + lane * 2) = src[lane];
              else if (st_a1) dstb[lane * 2] = src[lane].x;
              if (st_b2) *reinterpret_cast<float2*>(dstb + 64 + lane * 2) = src[32 + lane];
              else if (st_b1) dstb[64 + lane * 2] = src[32 + lane].x;
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
