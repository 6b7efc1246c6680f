// Included by smpl.cu (namespace tp).  Large-batch SMPL path (BASELINE.json configs[3]: 10^4..10^5 bodies): the blend
// contraction on tcgen05 with the skinning as its epilogue -- the blended rest-pose vertices exist only in TMEM/registers.
//
//   D[vertex 128, body 32] (one accumulator per coordinate plane) = Blend_plane[128, K = 256] . Coef[32, 256]^T
//
// * A operand (the vertex tile's blend rows, 3 planes x 128 rows x 256 bf16 = 196 KB) lives in TENSOR MEMORY (384 columns),
//   loaded once per (CTA, tile); B operand (32 bodies' coefficient rows, written by k_smpl_prepare as the 128-byte-swizzle
//   K-major image) arrives with ONE 16 KB cp.async.bulk per group; the bodies' joint transforms (32 x 24 x 12 fp32) with one
//   36 KB bulk copy.  3 x 16 tcgen05.mma (M 128, N 32, K 16) per group, fp32 accumulators in 96 TMEM columns.
// * Epilogue: 16 warps; warp w owns TMEM lane quadrant w % 4 (32 vertices) and the bodies 8 (w / 4) .. +7 of the group.
//   tcgen05.ld puts x,y,z of (lane = vertex) x (8 bodies) in registers and frees the accumulator at once, so the MMAs of the
//   next group run under this group's skinning.  Lane = vertex and the body is warp-uniform: the four 48-byte joint-transform
//   gathers of a vertex are 16-byte shared-memory loads whose addresses coincide across lanes that share a joint (broadcast),
//   and the weighted sums are packed fp32x2 FMAs (FFMA2).  The skinned tile goes through a per-warp slab: coalesced 8-byte
//   vertex stores, and the dense joint regressors as 3xTF32 mma.sync (rows = regressor rows, k = the warp's 32 vertices,
//   n = 8 bodies x 3 coordinates) -- the four quadrant warps of a body set add their partial sums in a fixed order.
// * Work = (vertex tile, body group) pairs in tile-major order, cut into equal contiguous ranges, one per SM (persistent
//   CTAs): a CTA reloads the A operand at most twice.
constexpr int kUsThreads = 576;                 // 16 epilogue warps, producer warp, MMA warp
constexpr int kUsEpiWarps = 16;
constexpr int kUsGB = 32;                       // bodies per group (MMA N)
constexpr int kUsVT = 128;                      // vertices per tile (MMA M)
constexpr int kUsK = 256;
constexpr int kUsBStages = 3;
constexpr int kUsChunkGroups = 128;            // bodies per L2-resident chunk / 32
constexpr int kUsBBytes = kUsGB * kUsK * 2;                 // 16 KB coefficient image of one group
constexpr int kUsTBytes = kUsGB * kJ * 12 * 4;              // 36 KB joint transforms of one group
constexpr int kUsSlabPitch = 100;                           // floats per body in a warp's slab (96 + 4: 16-byte rows, bank shift 4)
constexpr int kUsSlabFloats = 8 * kUsSlabPitch;
constexpr int kUsFragBytes = 4 * 4 * 32 * 8 * 4;            // regressor A fragments [quadrant][k-step][lane][hi 4 | lo 4]
constexpr size_t kUsTileImage = (size_t)3 * 4 * kUsVT * 128;  // blend image of one vertex tile: [plane][K block][row][128 B swizzled]
constexpr size_t kUsOffB = 0;
constexpr size_t kUsOffT = kUsOffB + (size_t)kUsBStages * kUsBBytes;
constexpr size_t kUsOffSlab = kUsOffT + (size_t)2 * kUsTBytes;
constexpr size_t kUsOffFrag = kUsOffSlab + (size_t)kUsEpiWarps * kUsSlabFloats * 4;
constexpr size_t kUsOffBar = kUsOffFrag + kUsFragBytes;
constexpr size_t kUsSmem = kUsOffBar + 256 + 1024;          // + alignment slack
constexpr uint32_t kUsIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kUsGB >> 3) << 17) | ((uint32_t)(kUsVT >> 4) << 24);

__device__ __forceinline__ void us_mb_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void us_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
                   "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void us_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void us_tmem_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
      "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
      "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
      "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void us_tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// packed fp32x2 multiply-add (FFMA2): d = a * b + c on both halves
__device__ __forceinline__ float2 us_fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}\n"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 us_mul2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}\n"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ void us_mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct UsParams {
  tp_smpl_model m;
  int n, ngroups, ntiles, nreg;
  const unsigned char* coef_img;   // [ngroups][16 KB]
  const float* A;                  // [ngroups * 32][24][12]
  const float* jreg;               // [nreg][vp]
  float* verts;                    // [n][n_verts][3] or null
  float* jpart;                    // [n][ntiles][nreg][3]
};

__global__ void __launch_bounds__(kUsThreads, 1) k_smpl_lbs_um(const UsParams p) {
  extern __shared__ __align__(1024) unsigned char us_raw[];
  unsigned char* smem = us_raw + ((1024u - (smem_u32(us_raw) & 1023u)) & 1023u);    // offset form: the pointer stays in the shared window (LDS/STS)
  unsigned char* s_b = smem + kUsOffB;
  unsigned char* s_t = smem + kUsOffT;                 // joint transforms, 2 stages; doubles as the TMEM-fill staging area (4 x 16 KB)
  float* s_slab = reinterpret_cast<float*>(smem + kUsOffSlab);
  uint32_t* s_frag = reinterpret_cast<uint32_t*>(smem + kUsOffFrag);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kUsOffBar);
  uint64_t* b_full = bars;                 // [3]
  uint64_t* b_empty = bars + 3;            // [3]
  uint64_t* t_full = bars + 6;             // [2]
  uint64_t* t_empty = bars + 8;            // [2]
  uint64_t* acc_full = bars + 10;
  uint64_t* acc_empty = bars + 11;
  uint64_t* f_full = bars + 12;            // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 3; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], kUsEpiWarps); }
    mbar_init(acc_full, 1); mbar_init(acc_empty, kUsEpiWarps);
    for (int i = 0; i < 4; ++i) mbar_init(&f_full[i], 1);
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
  const uint32_t tmem_acc = tmem_base, tmem_a = tmem_base + 96;     // 3 x 32 accumulator columns, then 3 planes x 128 columns of A

  // Work order: bodies are taken in CHUNKS of kUsChunkGroups groups (4096 bodies: 6.8 MB of coefficient rows + transforms, which
  // stay in L2 while the 54 vertex tiles re-read them); inside a chunk the (tile, group) items are cut into equal contiguous
  // ranges, one per CTA, in tile-major order.
  pdl_wait();                       // coefficient image / transforms come from k_smpl_prepare
  pdl_launch_dependents();

  uint32_t ngrp = 0;                // groups this CTA has been through (ring stage / parity bookkeeping, all roles in step)
  uint32_t nfill = 0;
  const int q = warp & 3, sset = warp >> 2;                    // epilogue: TMEM lane quadrant, body set
  for (int chunk0 = 0; chunk0 < p.ngroups; chunk0 += kUsChunkGroups) {
  const int cgroups = p.ngroups - chunk0 < kUsChunkGroups ? p.ngroups - chunk0 : kUsChunkGroups;
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
        for (int i = 0; i < 4; ++i) { mbar_expect_tx(&f_full[i], 16384); us_bulk_g2s(s_t + i * 16384, img + (size_t)i * 16384, 16384, &f_full[i]); }
      }
      // regressor A fragments (m16n8k8, tf32 hi | lo) of this tile: [quadrant][k-step][lane][8]
      for (int i = tid; i < 4 * 4 * 32 * 4; i += kUsThreads) {
        const int e = i & 3, ln = (i >> 2) & 31, ks = (i >> 7) & 3, qq = i >> 9;
        const int row = (ln >> 2) + ((e & 1) ? 8 : 0), v = qq * 32 + ks * 8 + (ln & 3) + ((e & 2) ? 4 : 0);
        const float val = row < p.nreg ? __ldg(p.jreg + (size_t)row * p.m.vp + v0 + v) : 0.0f;
        const uint32_t hi = __float_as_uint(val) & 0xffffe000u;
        const float lo = val - __uint_as_float(hi);
        uint32_t* dst = s_frag + ((size_t)(qq * 4 + ks) * 32 + ln) * 8;
        dst[e] = hi; dst[4 + e] = __float_as_uint(lo);
      }
      for (int blk = 0; blk < 12; ++blk) {
        const int st = blk & 3;
        if (warp < 4) {
          mbar_wait(&f_full[st], (nfill * 3 + (uint32_t)(blk >> 2)) & 1u);
          const int row = warp * 32 + lane;
          const unsigned char* rowp = s_t + (size_t)st * 16384 + (size_t)row * 128;
          uint32_t v[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 x = *reinterpret_cast<const uint4*>(rowp + ((c ^ (row & 7)) << 4));
            v[4 * c] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
          }
          us_tmem_st32(tmem_a + (uint32_t)(blk * 32) + ((uint32_t)(warp * 32) << 16), v);   // block = plane * 4 + K block: 32 columns each
        }
        __syncthreads();
        if (warp == 16 && blk + 4 < 12 && elect_one()) {
          mbar_expect_tx(&f_full[st], 16384);
          us_bulk_g2s(s_t + st * 16384, img + (size_t)(blk + 4) * 16384, 16384, &f_full[st]);
        }
      }
      if (warp < 4) asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      ++nfill;
    }

    if (warp == 16) {
      // ================================================================ producer
      for (int g = g0; g < g1; ++g, ++ngrp) {
        const uint32_t sb = ngrp % kUsBStages, pb = (ngrp / kUsBStages) & 1u;
        mbar_wait(&b_empty[sb], pb ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&b_full[sb], kUsBBytes);
          us_bulk_g2s(s_b + (size_t)sb * kUsBBytes, p.coef_img + (size_t)g * kUsBBytes, kUsBBytes, &b_full[sb]);
        }
        __syncwarp();
        const uint32_t stt = ngrp & 1u, pt = (ngrp >> 1) & 1u;
        mbar_wait(&t_empty[stt], pt ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&t_full[stt], kUsTBytes);
          us_bulk_g2s(s_t + (size_t)stt * kUsTBytes, p.A + (size_t)g * (kUsTBytes / 4), kUsTBytes, &t_full[stt]);
        }
        __syncwarp();
      }
    } else if (warp == 17) {
      // ================================================================ MMA issuer
      for (int g = g0; g < g1; ++g, ++ngrp) {
        const uint32_t sb = ngrp % kUsBStages, pb = (ngrp / kUsBStages) & 1u;
        mbar_wait(&b_full[sb], pb);
        mbar_wait(acc_empty, (ngrp & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint64_t db0 = umma_desc_sw128(smem_u32(s_b + (size_t)sb * kUsBBytes));
        if (elect_one()) {
#pragma unroll
          for (int pl = 0; pl < 3; ++pl)
#pragma unroll
            for (int i = 0; i < 16; ++i)
              us_umma_ts(tmem_acc + (uint32_t)(pl * 32), tmem_a + (uint32_t)(pl * 128 + i * 8),
                         db0 + (uint64_t)((i >> 2) * 256 + (i & 3) * 2), kUsIdesc, i == 0 ? 0u : 1u);
          umma_commit(&b_empty[sb]);
          umma_commit(acc_full);
        }
        __syncwarp();
      }
    } else {
      // ================================================================ epilogue: skinning, vertex store, joint regressors
      const int v = v0 + q * 32 + lane;
      float2 sw2[4]; int sj[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float w = k < p.m.ks ? __ldg(p.m.skin_w + (size_t)v * p.m.ks + k) : 0.0f;
        sw2[k] = make_float2(w, w);
        sj[k] = (k < p.m.ks ? __ldg(p.m.skin_idx + (size_t)v * p.m.ks + k) : 0) * 48;
      }
      const float tx = __ldg(p.m.template_pad + (size_t)v * 3), ty = __ldg(p.m.template_pad + (size_t)v * 3 + 1),
                  tz = __ldg(p.m.template_pad + (size_t)v * 3 + 2);
      const bool vvalid = v < p.m.n_verts;
      float* slab = s_slab + (size_t)warp * kUsSlabFloats;
      const uint32_t* frag = s_frag + (size_t)q * 4 * 32 * 8 + lane * 8;
      const uint32_t t_lane = (uint32_t)(q * 32) << 16;
      const int64_t nv3 = (int64_t)p.m.n_verts * 3;
      const int gq = lane >> 2, tq = lane & 3;
      // per-tile store bounds of this lane's two float2 slots (only the last tile is ragged)
      const int64_t vbase = (int64_t)(v0 + q * 32) * 3;
      const int64_t vrem = nv3 - vbase;
      const bool st_a2 = lane * 2 + 1 < vrem, st_a1 = lane * 2 < vrem;
      const bool st_b2 = lane < 16 && 64 + lane * 2 + 1 < vrem, st_b1 = lane < 16 && 64 + lane * 2 < vrem;
      // partial-sum slots this thread adds up after the regressor MMAs: value i = row * 24 + body * 3 + coordinate
      int js_off[3]; bool js_ok[3]; int js_bi[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int i = q * 32 + lane + 128 * k;
        const int r = i / 24, col = i - r * 24, bi = col / 3, c = col - bi * 3;
        js_ok[k] = i < p.nreg * 24;
        js_bi[k] = bi;
        js_off[k] = ((bi * p.ntiles + tile) * p.nreg + r) * 3 + c;
      }
      for (int g = g0; g < g1; ++g, ++ngrp) {
        float px[8], py[8], pz[8];
        mbar_wait(acc_full, ngrp & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        us_tmem_ld8(tmem_acc + t_lane + (uint32_t)(sset * 8), px);
        us_tmem_ld8(tmem_acc + t_lane + (uint32_t)(32 + sset * 8), py);
        us_tmem_ld8(tmem_acc + t_lane + (uint32_t)(64 + sset * 8), pz);
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) us_mb_arrive(acc_empty);
        const uint32_t stt = ngrp & 1u, pt = (ngrp >> 1) & 1u;
        mbar_wait(&t_full[stt], pt);
        const unsigned char* As = s_t + (size_t)stt * kUsTBytes + (size_t)(sset * 8) * (kJ * 48);
#pragma unroll
        for (int bi = 0; bi < 8; ++bi) {
          const unsigned char* Ab = As + bi * (kJ * 48);
          float2 T[6];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 r0 = *reinterpret_cast<const float4*>(Ab + sj[k]);
            const float4 r1 = *reinterpret_cast<const float4*>(Ab + sj[k] + 16);
            const float4 r2 = *reinterpret_cast<const float4*>(Ab + sj[k] + 32);
            if (k == 0) {
              T[0] = us_mul2(sw2[0], make_float2(r0.x, r0.y)); T[1] = us_mul2(sw2[0], make_float2(r0.z, r0.w));
              T[2] = us_mul2(sw2[0], make_float2(r1.x, r1.y)); T[3] = us_mul2(sw2[0], make_float2(r1.z, r1.w));
              T[4] = us_mul2(sw2[0], make_float2(r2.x, r2.y)); T[5] = us_mul2(sw2[0], make_float2(r2.z, r2.w));
            } else {
              T[0] = us_fma2(sw2[k], make_float2(r0.x, r0.y), T[0]); T[1] = us_fma2(sw2[k], make_float2(r0.z, r0.w), T[1]);
              T[2] = us_fma2(sw2[k], make_float2(r1.x, r1.y), T[2]); T[3] = us_fma2(sw2[k], make_float2(r1.z, r1.w), T[3]);
              T[4] = us_fma2(sw2[k], make_float2(r2.x, r2.y), T[4]); T[5] = us_fma2(sw2[k], make_float2(r2.z, r2.w), T[5]);
            }
          }
          const float x = px[bi] + tx, y = py[bi] + ty, z = pz[bi] + tz;
          const float ox = fmaf(T[0].x, x, fmaf(T[0].y, y, fmaf(T[1].x, z, T[1].y)));
          const float oy = fmaf(T[2].x, x, fmaf(T[2].y, y, fmaf(T[3].x, z, T[3].y)));
          const float oz = fmaf(T[4].x, x, fmaf(T[4].y, y, fmaf(T[5].x, z, T[5].y)));
          float* o = slab + bi * kUsSlabPitch + lane * 3;
          o[0] = vvalid ? ox : 0.0f; o[1] = vvalid ? oy : 0.0f; o[2] = vvalid ? oz : 0.0f;
        }
        __syncwarp();
        if (lane == 0) us_mb_arrive(&t_empty[stt]);
        const int body0 = g * kUsGB + sset * 8;
        const int nvalid = p.n - body0;                       // bodies of this set that exist (>= 8 except in the last group)
        // coalesced vertex store: 384 contiguous bytes per (warp, body); the per-lane bounds were settled once per tile
        if (p.verts) {
          float* dstb = p.verts + (int64_t)body0 * nv3 + vbase;
#pragma unroll
          for (int bi = 0; bi < 8; ++bi) {
            if (bi < nvalid) {
              const float2* src = reinterpret_cast<const float2*>(slab + bi * kUsSlabPitch);
              if (st_a2) *reinterpret_cast<float2*>(dstb + lane * 2) = src[lane];
              else if (st_a1) dstb[lane * 2] = src[lane].x;
              if (st_b2) *reinterpret_cast<float2*>(dstb + 64 + lane * 2) = src[32 + lane];
              else if (st_b1) dstb[64 + lane * 2] = src[32 + lane].x;
            }
            dstb += nv3;
          }
        }
        // joint regressors on the tile while it is on chip: D[row 16, (body, coord) 24] += J[16, 32 v] . V[32 v, 24], 3xTF32
        if (p.nreg > 0) {
          float d[3][4];
#pragma unroll
          for (int nt = 0; nt < 3; ++nt) d[nt][0] = d[nt][1] = d[nt][2] = d[nt][3] = 0.0f;
          int boff[3];
#pragma unroll
          for (int nt = 0; nt < 3; ++nt) { const int col = nt * 8 + gq; boff[nt] = (col / 3) * kUsSlabPitch + (col % 3) + tq * 3; }
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint4 fh = *reinterpret_cast<const uint4*>(frag + ks * 32 * 8);
            const uint4 fl = *reinterpret_cast<const uint4*>(frag + ks * 32 * 8 + 4);
            const uint32_t ah[4] = {fh.x, fh.y, fh.z, fh.w}, al[4] = {fl.x, fl.y, fl.z, fl.w};
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) {
              const float b0 = slab[boff[nt] + ks * 24], b1 = slab[boff[nt] + ks * 24 + 12];
              const uint32_t b0h = __float_as_uint(b0) & 0xffffe000u, b1h = __float_as_uint(b1) & 0xffffe000u;
              const uint32_t b0l = __float_as_uint(b0 - __uint_as_float(b0h)), b1l = __float_as_uint(b1 - __uint_as_float(b1h));
              us_mma_tf32(d[nt], al, b0h, b1h);
              us_mma_tf32(d[nt], ah, b0l, b1l);
              us_mma_tf32(d[nt], ah, b0h, b1h);
            }
          }
          __syncwarp();                                      // the slab is free: it now carries this warp's partial sums [16 rows][24]
#pragma unroll
          for (int nt = 0; nt < 3; ++nt) {
            *reinterpret_cast<float2*>(slab + gq * 24 + nt * 8 + tq * 2) = make_float2(d[nt][0], d[nt][1]);
            *reinterpret_cast<float2*>(slab + (gq + 8) * 24 + nt * 8 + tq * 2) = make_float2(d[nt][2], d[nt][3]);
          }
          asm volatile("bar.sync %0, 128;\n" ::"r"(1 + sset) : "memory");
          {
            const float* s0 = s_slab + (size_t)(sset * 4) * kUsSlabFloats + q * 32 + lane;
            float* jp = p.jpart + (int64_t)body0 * p.ntiles * p.nreg * 3;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              if (js_ok[k] && js_bi[k] < nvalid)
                jp[js_off[k]] = (s0[128 * k] + s0[kUsSlabFloats + 128 * k]) + (s0[2 * kUsSlabFloats + 128 * k] + s0[3 * kUsSlabFloats + 128 * k]);
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
