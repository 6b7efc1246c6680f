// Included by gru.cu after gru_dual.inl (same namespace, same helpers; umma.cuh is included by gru.cu).
//
// K2 on the 5th-generation tensor cores with the WHOLE of W_hh resident on chip (bf16 mode, batch <= 32, no h0).
//
// The recurrence is a weight-stationary problem: per step and direction gh^T [3H x B] = W_hh [3H x H] . h_{t-1}^T, the
// same W_hh for every step.  Swapping the operands makes W_hh the M side of tcgen05.mma (M = 128 rows, N = 32 batch
// columns, fp32 accumulator in TMEM), and the A operand of tcgen05.mma may live in TENSOR MEMORY: an SM has 256 KB of
// TMEM next to its 227 KB of shared memory, so 128 SMs hold 128 x (224 + 160) KB = 48 MB -- both directions' W_hh at
// H = 2048 (50.3 MB bf16) fit once the few KB the pipeline needs are taken off.  W_hh is read from HBM ONCE per launch
// (the algorithmic minimum) instead of being re-streamed from L2 every step (1.03 GB of L2 -> SM traffic per launch in
// k_gru_bf16_dual).
//
// Decomposition: a direction is cut into PAIRS of CTAs (a cluster of 2).  A pair owns 64 hidden units = 192 rows of
// W_hh ({r,z,n} x 64) and splits K: CTA k of the pair holds columns [k H/2, (k+1) H/2) of those rows and reads only that
// half of h_{t-1} (64 KB instead of 128 KB per step).  The 192 rows are two MMA tiles:
//   tile 1 (M = 128): r,z,n of units 0..31 and r of units 32..63 -- A in TMEM (up to K = 896; the rest in smem),
//   tile 2 (M = 64) : z,n of units 32..63                        -- A in shared memory (UMMA 128-byte-swizzle image).
// After the K loop each CTA holds partial sums of all 192 rows; CTA k finalises units 32k..32k+31, so the other half
// of the partials crosses to the peer through distributed shared memory: ONE 12 KB cp.async.bulk (shared::cta ->
// shared::cluster) per step that completes a transaction barrier in the peer (per-thread st.shared::cluster stores + a
// cluster-scope release arrive cost 1.7-2K cycles per step, 1.2K of them in the release's memory barrier).  Gate math keeps the fp32 state in registers, writes the next
// step's MMA operand (bf16, stored in global memory as the swizzled smem image so a K block is one 4 KB bulk copy) and
// arrives on the direction's grid barrier.
//
// Warp roles: warps 0-7 epilogue / gate math (warps 0-3 read tile 1's accumulator, warps 4-7 tile 2's; TMEM lane quadrant =
// warp % 4), warp 8 producer (grid-barrier wait + h bulk copies), warps 9 and 10 MMA issuers -- ONE PER TILE: a single thread
// issuing all 128 tcgen05.mma of a step needed 4.4K cycles (35 per MMA: descriptor arithmetic and uniform-register moves of the
// issuing thread, not the tensor pipe, whose 17 + 25 cycles per tile-1 / tile-2 pair add up to 2.7K); the two chains are
// independent (own accumulator, own A operand), so two threads issue them side by side.  Directions never synchronise with
// each other.
constexpr int kUmThreads = 352;                   // 8 epilogue / gate warps, producer warp, two MMA warps (one per tile)
constexpr int kUmEpiThreads = 256;
constexpr int kUmStages = 3;                      // h ring: stages of up to FOUR 64-column K blocks (32 rows x 128 B = 4 KB each)
constexpr int kUmBlockBytes = 32 * 128;
constexpr int kUmStageBytes = 4 * kUmBlockBytes;  // 16 KB = one 128-row W block: the ring doubles as the TMEM-fill staging area
constexpr int kUmStgBytes = 3 * 8 * 512;          // one staging buffer: [gate 3][batch / 4][unit 32][batch % 4] fp32 = 12 KB
// Measured on the B200 (scripts/micro/umma_lat.cu, straight-line issue): tcgen05.mma kind::f16 at N = 32 takes 17 cycles
// with A in TMEM (M = 128), 25 with A in shared memory at M = 64 (3 KB of operands at 128 B/clk) and 40 at M = 128
// (5 KB) -- the shared-memory operand fetch is the floor, which is the second reason to keep A in TMEM.  Back-to-back
// MMAs into the same accumulator are NOT slower than alternating accumulators, so each tile has one accumulator.
struct UmCfg { int max_tmem_k, sync_mode, trace_set; };
inline UmCfg um_cfg() {
  static const UmCfg c = [] {
    UmCfg v;
    const char* e = getenv("TP_UM_TMEMK");       // debug: keep less of tile 1 in TMEM (multiple of 64; 0 = every A operand from shared memory)
    v.max_tmem_k = e ? atoi(e) : 1 << 20;
    // grid-barrier protocol around the h_t exchange (bit flags): 1 = consumer executes fence.acq_rel.gpu after its relaxed
    // polls; 2 = consumer executes fence.proxy.async before its bulk copies; 4 = producers execute fence.proxy.async.global
    // after their h_t stores (generic -> async proxy ordering on the writer side); 8 = consumer polls with ld.acquire.gpu
    // (the formal acquire without a separate fence).  Default 8 | 4.
    e = getenv("TP_UM_SYNC");
    v.sync_mode = e ? atoi(e) : 12;
    e = getenv("TP_UM_TRACESET");                // debug: 1 = trace slots 1..4 hold sub-phases of the epilogue's exchange
    v.trace_set = e ? atoi(e) : 0;
    return v;
  }();
  return c;
}
constexpr uint32_t kUmIdesc128 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t kUmIdesc64 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);

struct UmGeom {
  int KH, KT, nkb, ntm;             // K per CTA, K of tile 1 in TMEM, 64-column blocks per CTA, of which in TMEM
  int bps;                          // K blocks per ring stage: the largest of 4, 2, 1 that divides nkb
  size_t tm_bytes, rest_bytes, t2_bytes, image_bytes;
  uint32_t tmem_cols;
  size_t smem_bytes;
};
__host__ __device__ inline UmGeom um_geom(int H, int max_tmem_k) {
  UmGeom g;
  g.KH = H / 2;
  int cap = 896;                                          // 448 columns of A beside 2 x 32 accumulator columns
  if (max_tmem_k < cap) cap = max_tmem_k < 0 ? 0 : max_tmem_k / 64 * 64;
  g.KT = g.KH < cap ? g.KH : cap;
  g.nkb = g.KH / 64;
  g.ntm = g.KT / 64;
  g.bps = g.nkb % 4 == 0 ? 4 : g.nkb % 2 == 0 ? 2 : 1;
  g.tm_bytes = (size_t)g.ntm * 128 * 128;                 // [block][128 rows][128 B], swizzled (staged through smem into TMEM)
  g.rest_bytes = (size_t)(g.nkb - g.ntm) * 128 * 128;     // [block][128 rows][128 B], swizzled
  g.t2_bytes = (size_t)g.nkb * 64 * 128;                  // [block][64 rows][128 B], swizzled
  g.image_bytes = g.tm_bytes + g.rest_bytes + g.t2_bytes; // = 384 * KH
  const uint32_t need = 64 + (uint32_t)g.KT / 2;
  g.tmem_cols = need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
  // shared memory: resident W, h ring (the CTA's own partial sums alias its first 12 KB between the last MMA of a step
  // and the next grid barrier, when the ring is idle), the partial sums the peer sends, barriers
  g.smem_bytes = 1024 + g.rest_bytes + g.t2_bytes + (size_t)kUmStages * kUmStageBytes + (size_t)kUmStgBytes + 256;
  return g;
}

// (gate, unit within the pair's 64) of local row r of tile 1 (r < 128) / tile 2 (r < 64)
__host__ __device__ inline void um_row_t1(int r, int& gate, int& unit) {
  if (r < 96) { gate = r >> 5; unit = r & 31; } else { gate = 0; unit = 32 + (r - 96); }
}
__host__ __device__ inline void um_row_t2(int r, int& gate, int& unit) { gate = 1 + (r >> 5); unit = 32 + (r & 31); }

// weight_hh [3H,H] fp32 -> per-CTA images (see UmGeom); one thread per 16-byte chunk
__global__ void k_pack_whh_umma(const float* __restrict__ w, uint4* __restrict__ dst, int H, int max_tmem_k) {
  const UmGeom g = um_geom(H, max_tmem_k);
  const size_t chunks_per_cta = g.image_bytes / 16;
  const size_t total = chunks_per_cta * (size_t)(H / 32);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cta = (int)(i / chunks_per_cta);
    const size_t off = (i - (size_t)cta * chunks_per_cta) * 16;
    const int pair = cta >> 1, khalf = cta & 1;
    int gate, unit, k;
    if (off < g.tm_bytes + g.rest_bytes) {          // tile 1: blocks [0, ntm) end up in TMEM, the others stay in shared memory
      const int blk = (int)(off / 16384), rem = (int)(off % 16384), row = rem >> 7, ch = ((rem & 127) >> 4) ^ (row & 7);
      um_row_t1(row, gate, unit);
      k = blk * 64 + ch * 8;
    } else {
      const size_t o = off - g.tm_bytes - g.rest_bytes;
      const int blk = (int)(o / 8192), rem = (int)(o % 8192), row = rem >> 7, ch = ((rem & 127) >> 4) ^ (row & 7);
      um_row_t2(row, gate, unit);
      k = blk * 64 + ch * 8;
    }
    const float* s = w + ((size_t)gate * H + (size_t)pair * 64 + unit) * H + (size_t)khalf * g.KH + k;
    const float4 a = *reinterpret_cast<const float4*>(s), b = *reinterpret_cast<const float4*>(s + 4);
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
    uint4 o;
    o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
    o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
    dst[i] = o;
  }
}

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
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
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void mb_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mb_wait_cluster(uint64_t* b, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t addr = sm_u32(b);
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ float4 ldnc_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

#define UM_TRACE(step, slot)                                                                          \
  do {                                                                                                \
    if (p.trace) p.trace[((size_t)blockIdx.x * p.max_steps + (step)) * 8 + (slot)] = clock64();        \
  } while (0)

__device__ __forceinline__ void bulk_s2peer(uint32_t dst_cluster, const void* src, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
                   "r"(dst_cluster), "r"(sm_u32(src)), "r"(bytes), "r"(bar_cluster) : "memory");
}
// SFU sigmoid / tanh: ex2.approx + rcp.approx (2^-22 relative each); saturate cleanly (ex2 -> inf -> rcp -> 0).  The generic
// __expf / __fdividef pair costs ~15 instructions per gate with range fix-ups; the gate phase is a serial section of every step.
__device__ __forceinline__ float um_sigmoid(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(e) : "f"(-1.4426950408889634f * x));
  asm("rcp.approx.ftz.f32 %0, %1;\n" : "=f"(r) : "f"(1.0f + e));
  return r;
}
__device__ __forceinline__ float um_tanh(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(e) : "f"(-2.8853900817779268f * x));
  asm("rcp.approx.ftz.f32 %0, %1;\n" : "=f"(r) : "f"(1.0f + e));
  return fmaf(2.0f, r, -1.0f);
}

struct UmParams {
  GruParams g;
  const unsigned char* w_img[kMaxJobs];   // tp_pack_whh_umma images of the matmul jobs
  int max_tmem_k, sync_mode, trace_set;
};

__device__ __forceinline__ void um_load_block(uint32_t (&v)[32], const uint4* src) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const uint4 x = ldg_stream(src + q);
    v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
  }
}
__device__ __forceinline__ float ldnc_f1(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];\n" : "=f"(v) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(kUmThreads, 1) k_gru_umma(const UmParams up) {
  const GruParams& p = up.g;
  const long long t_start = clock64();
  extern __shared__ __align__(1024) unsigned char smem_um_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_um_raw) + 1023) & ~(uintptr_t)1023);
  const int H = p.H, B = p.B;
  const UmGeom geo = um_geom(H, up.max_tmem_k);
  unsigned char* s_rest = smem;                                   // tile 1, K >= KT: [block][128 rows][128 B]
  unsigned char* s_t2 = s_rest + geo.rest_bytes;                  // tile 2: [block][64 rows][128 B]
  unsigned char* s_ring = s_t2 + geo.t2_bytes;                    // h ring
  float* s_own = reinterpret_cast<float*>(s_ring);                // partial sums of this CTA for its own units (aliases the idle ring)
  float* s_send = reinterpret_cast<float*>(s_ring + kUmStgBytes); // partial sums for the peer's units, bulk-copied into its s_peer (idle ring too)
  float* s_peer = reinterpret_cast<float*>(s_ring + (size_t)kUmStages * kUmStageBytes);  // partial sums the peer CTA sends
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(s_peer) + kUmStgBytes);
  uint64_t* full = bars;                  // [stages]
  uint64_t* empty = bars + kUmStages;     // [stages]
  uint64_t* tfull = bars + 2 * kUmStages;     // [stages]: TMEM-fill staging (prologue only)
  uint64_t* tempty = bars + 3 * kUmStages;
  uint64_t* wres = bars + 4 * kUmStages;
  uint64_t* acc_full = wres + 1;
  uint64_t* xbar = wres + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wres + 3);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank();                           // = K half of this CTA
  const int ctas_per_dir = H / 32;
  const int d = blockIdx.x / ctas_per_dir;                        // direction / job
  const int cta_in_dir = blockIdx.x - d * ctas_per_dir;
  const int pair = cta_in_dir >> 1;

  __shared__ tp_gru_job sjobs[kMaxJobs];
  for (int i = tid; i < (int)(sizeof(tp_gru_job) * kMaxJobs / 4); i += kUmThreads)
    reinterpret_cast<int*>(sjobs)[i] = reinterpret_cast<const int*>(p.jobs)[i];
  if (tid == 0) {
    for (int i = 0; i < kUmStages; ++i) { mb_init(&full[i], 1); mb_init(&empty[i], 2); mb_init(&tfull[i], 1); mb_init(&tempty[i], 4); }   // empty / acc_full: one commit per MMA warp
    mb_init(wres, 1); mb_init(acc_full, 2); mb_init(xbar, 1);      // xbar: one local arrive.expect_tx + the peer's 12 KB bulk copy
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(sm_u32(tmem_slot)), "r"(geo.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d1 = tmem_base, tmem_d2 = tmem_base + 32, tmem_a = tmem_base + 64;
  const tp_gru_job& jb = sjobs[d];
  const unsigned char* img = up.w_img[d] + (size_t)cta_in_dir * geo.image_bytes;
  const int steps = jb.steps;
  if (tid == 0 && p.trace) { p.trace[(size_t)blockIdx.x * p.max_steps * 8 + 1] = t_start; UM_TRACE(0, 2); }

  // ---- resident weights: W_hh does not depend on the previous kernel, so all of this runs before the PDL wait.
  // Shared-memory part: bulk copies straight into place.  TMEM part: 16 KB blocks are bulk-copied into the (still idle)
  // h ring, three in flight, and each epilogue thread moves its row (128 swizzled bytes, conflict-free 16-byte reads) into
  // its TMEM lane with one tcgen05.st per block.  (Reading the rows straight from global memory -- 32 different lines per
  // warp load -- took 50K cycles; this takes what HBM takes.)
  if (warp == 8 && steps > 1) {
    if (elect_one()) {
      mb_expect_tx(wres, (uint32_t)(geo.rest_bytes + geo.t2_bytes));   // tile-2 blocks are 8 KB, rest blocks 16 KB: 8 KB copies cover both
      const unsigned char* src = img + geo.tm_bytes;
      for (size_t o = 0; o < geo.rest_bytes + geo.t2_bytes; o += 8192) bulk_g2s(s_rest + o, src + o, 8192, wres);
    }
    __syncwarp();
    for (int blk = 0; blk < geo.ntm; ++blk) {
      const int st = blk % kUmStages;
      mb_wait(&tempty[st], ((blk / kUmStages) & 1) ^ 1);               // four epilogue warps have read the previous block out
      if (elect_one()) {
        mb_expect_tx(&tfull[st], 16384);
        bulk_g2s(s_ring + (size_t)st * kUmStageBytes, img + (size_t)blk * 16384, 16384, &tfull[st]);
      }
      __syncwarp();
    }
  }
  if (warp < 4 && steps > 1) {
    const int row = warp * 32 + lane;
    const uint32_t dst = tmem_a + ((uint32_t)(warp * 32) << 16);
    for (int blk = 0; blk < geo.ntm; ++blk) {
      const int st = blk % kUmStages;
      mb_wait(&tfull[st], (blk / kUmStages) & 1);
      const unsigned char* rowp = s_ring + (size_t)st * kUmStageBytes + (size_t)row * 128;
      uint32_t v[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint4 x = *reinterpret_cast<const uint4*>(rowp + ((q ^ (row & 7)) << 4));
        v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
      }
      tmem_st32(dst + (uint32_t)(blk * 32), v);
      __syncwarp();
      if (lane == 0) mb_arrive(&tempty[st]);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) UM_TRACE(0, 3);
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  cluster_sync_all();                       // barriers of both CTAs initialised before any remote arrive; TMEM A visible to the MMA thread
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (tid == 0) UM_TRACE(0, 4);

  pdl_wait();
  pdl_launch_dependents();
  if (tid == 0) UM_TRACE(0, 5);

  unsigned int* counter = p.barrier + 16 * d;                             // 64 bytes apart: four directions fit the 256 zeroed bytes of the workspace header
  __nv_bfloat16* hlp = p.hbuf_lp + (int64_t)(d * 2) * p.lp_slot;          // two parity slots of [H/64 blocks][32 rows][64] swizzled
  const int bps = geo.bps;

  if (warp == 8) {
    // ===================== producer: grid-barrier wait, then the h_{t-1} half of this CTA, `bps` K blocks per stage.
    // The whole warp runs the loop so that every address / stage index is warp-uniform (UBLKCP and UTCHMMA take uniform
    // registers; inside a lane-0-only region the compiler has to broadcast each operand through an ELECT / R2UR loop,
    // measured at ~115 cycles per MMA); only the asynchronous instructions themselves are issued by one elected lane.
    uint32_t st = 0, ph = 1;
    const uint32_t stage_tx = (uint32_t)bps * kUmBlockBytes;
    for (int s = 1; s < steps; ++s) {
      const unsigned int target = (unsigned int)s * (unsigned int)ctas_per_dir;
      unsigned int seen;
      if (up.sync_mode & 8) {
        // acquire loads: the formal synchronises-with for the h_{t-1} words the other CTAs released.  Only this warp polls
        // and only the async proxy reads that data, so the L1 invalidate an acquire implies costs nobody anything.
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(seen) : "l"(counter) : "memory");
        } while (seen < target);
      } else {
        do {
          asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];\n" : "=r"(seen) : "l"(counter) : "memory");
        } while (seen < target);
      }
      if (lane == 0 && !up.trace_set) UM_TRACE(s, 1);
      if (up.sync_mode & 1) asm volatile("fence.acq_rel.gpu;\n" ::: "memory");
      if (up.sync_mode & 2) asm volatile("fence.proxy.async;\n" ::: "memory");
      if (lane == 0 && !up.trace_set) UM_TRACE(s, 2);
      const unsigned char* hsrc = reinterpret_cast<const unsigned char*>(hlp + (int64_t)((s + 1) & 1) * p.lp_slot) +
                                  (size_t)rank * geo.nkb * kUmBlockBytes;
      for (int kb = 0; kb < geo.nkb; kb += bps) {
        mb_wait(&empty[st], ph);
        if (elect_one()) {
          mb_expect_tx(&full[st], stage_tx);
          bulk_g2s(s_ring + (size_t)st * kUmStageBytes, hsrc + (size_t)kb * kUmBlockBytes, stage_tx, &full[st]);
        }
        __syncwarp();
        if (++st == kUmStages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 9 || warp == 10) {
    // ===================== MMA issuers (warp-uniform loops, one elected lane issues): warp 9 tile 1 (A in TMEM / rest in shared
    // memory), warp 10 tile 2 (A in shared memory).  Both wait on the same full barriers and both commit to empty / acc_full.
    const bool t1 = warp == 9;
    if (steps > 1) {
      mb_wait(wres, 0);
      if (lane == 0 && t1) UM_TRACE(0, 6);
      uint32_t st = 0, ph = 0;
      const uint64_t d_ring = umma_desc_sw128(sm_u32(s_ring)), d_t2 = umma_desc_sw128(sm_u32(s_t2)), d_rest = umma_desc_sw128(sm_u32(s_rest));
      const int total_stages = (steps - 1) * (geo.nkb / bps);
      int issued = 0;
      bool landed = false;                   // elected lane only: the current stage's full barrier has already been waited on
      for (int s = 1; s < steps; ++s) {
        for (int kb0 = 0; kb0 < geo.nkb; kb0 += bps, ++issued) {
          const uint32_t nst = st + 1 == kUmStages ? 0u : st + 1, nph = st + 1 == kUmStages ? ph ^ 1u : ph;
          const bool has_next = issued + 1 < total_stages && kb0 + bps < geo.nkb;   // look ahead inside a step only (the next step's h does not exist yet)
          const uint64_t db0 = d_ring + (uint64_t)(st * (kUmStageBytes >> 4));
          const uint64_t da20 = d_t2 + (uint64_t)(kb0 * 512);
          const uint32_t ta0 = tmem_a + (uint32_t)(kb0 * 32);
          const uint32_t acc0 = kb0 != 0 ? 1u : 0u;
          if (elect_one()) {
            if (!landed) mb_wait(&full[st], ph);
            landed = false;
            if (t1 && up.trace_set == 2 && kb0 / bps < 4) UM_TRACE(s, 2 * (kb0 / bps));
            if (t1 && kb0 == 0 && !up.trace_set) UM_TRACE(s, 3);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            // descriptor start addresses are in 16-byte units: a 4 KB h block = 256, an 8 KB tile-2 block = 512, a 16 KB rest block = 1024
            if (bps == 4 && kb0 + 4 <= geo.ntm) {
              // the common case as straight-line code: 4 K blocks x 4 k-steps.  The wait for the NEXT stage sits before the
              // last quarter, while the tensor pipe still has queued work.
              if (t1) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  if (i == 12 && has_next) { mb_wait(&full[nst], nph); landed = true; }
                  umma_f16_ts(tmem_d1, ta0 + (uint32_t)(i * 8), db0 + (uint64_t)((i >> 2) * 256 + (i & 3) * 2), kUmIdesc128, i == 0 ? acc0 : 1u);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  if (i == 12 && has_next) { mb_wait(&full[nst], nph); landed = true; }
                  umma_f16(tmem_d2, da20 + (uint64_t)((i >> 2) * 512 + (i & 3) * 2), db0 + (uint64_t)((i >> 2) * 256 + (i & 3) * 2), kUmIdesc64, i == 0 ? acc0 : 1u);
                }
              }
            } else {
              for (int j = 0; j < bps; ++j) {
                const int kb = kb0 + j;
                const uint64_t db = db0 + (uint64_t)(j * 256);
                if (!t1) {
                  const uint64_t da2 = da20 + (uint64_t)(j * 512);
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks) umma_f16(tmem_d2, da2 + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), kUmIdesc64, (kb | ks) != 0 ? 1u : 0u);
                } else if (kb < geo.ntm) {
                  const uint32_t ta = ta0 + (uint32_t)(j * 32);
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks) umma_f16_ts(tmem_d1, ta + (uint32_t)(ks * 8), db + (uint64_t)(ks * 2), kUmIdesc128, (kb | ks) != 0 ? 1u : 0u);
                } else {
                  const uint64_t da1 = d_rest + (uint64_t)((kb - geo.ntm) * 1024);
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks) umma_f16(tmem_d1, da1 + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), kUmIdesc128, (kb | ks) != 0 ? 1u : 0u);
                }
              }
            }
            umma_commit(&empty[st]);
            if (kb0 + bps >= geo.nkb) umma_commit(acc_full);
            if (t1 && up.trace_set == 2 && kb0 / bps < 4) UM_TRACE(s, 2 * (kb0 / bps) + 1);
          }
          __syncwarp();
          st = nst; ph = nph;
        }
        if (lane == 0 && t1 && !up.trace_set) UM_TRACE(s, 4);
      }
    }
  } else if (warp < 8) {
    // ===================== epilogue / gate math (8 warps)
    // single-step jobs from a zero state have no matmul: plain gate math, grid-strided over the epilogue threads
    for (int je = p.n_item_jobs; je < p.njobs; ++je) {
      const tp_gru_job& js = sjobs[je];
      for (int64_t i = blockIdx.x * (int64_t)kUmEpiThreads + tid; i < (int64_t)B * H; i += (int64_t)gridDim.x * kUmEpiThreads) {
        const int b = (int)(i / H), u = (int)(i - (int64_t)b * H);
        const float* gi = js.gi + ((int64_t)js.t_in0 * B + b) * js.ldg;
        const float r = um_sigmoid(gi[u] + js.b_hh[u]);
        const float z = um_sigmoid(gi[H + u] + js.b_hh[H + u]);
        const float n = um_tanh(gi[2 * H + u] + r * js.b_hh[2 * H + u]);
        const float h = (1.0f - z) * n;
        if (js.gates) {
          float* gt = js.gates + ((int64_t)js.t_out0 * B + b) * 4 * H;
          gt[u] = r; gt[H + u] = z; gt[2 * H + u] = n; gt[3 * H + u] = js.b_hh[2 * H + u];
        }
        if (js.y) js.y[((int64_t)js.t_out0 * B + b) * js.ldy + u] = h;
        if (js.y_lp) reinterpret_cast<__nv_bfloat16*>(js.y_lp)[((int64_t)js.t_out0 * B + b) * js.ldy_lp + u] = __float2bfloat16_rn(h);
        if (js.h_final) js.h_final[(int64_t)b * js.ld_hf + u] = h;
      }
    }
    // Gate thread (warp g, lane l) owns unit U = 64 pair + 32 rank + l for the batch rows 4g .. 4g+3: global accesses are
    // coalesced over units, and the staged partial sums [gate][batch / 4][unit][batch % 4] are read as conflict-free float4.
    const int U = pair * 64 + (int)rank * 32 + lane;
    const int b0 = warp * 4;
    float bh[3], hp[4];
#pragma unroll
    for (int g = 0; g < 3; ++g) bh[g] = __ldg(jb.b_hh + g * H + U);
#pragma unroll
    for (int j = 0; j < 4; ++j) hp[j] = 0.0f;
    // Accumulator rows: warps 0-3 read tile-1 row 32 warp + lane, warps 4-7 (lanes 0..15) tile-2 row 16 (warp - 4) + lane, 32
    // batch columns each.  Rows of the CTA's own units go to s_own, the peer's to s_send; float4 q of a row lands at
    // [gate][q][unit][4] (512 contiguous bytes per warp store instruction, conflict-free).
    const uint32_t peer = rank ^ 1u;
    uint32_t dst;
    {
      int gate, unit;
      if (warp < 4) um_row_t1(warp * 32 + lane, gate, unit);
      else um_row_t2((warp - 4) * 16 + (lane & 15), gate, unit);  // tile 2 holds units 32..63 only: CTA 1's
      const bool own = (uint32_t)(unit >> 5) == rank;
      dst = sm_u32(own ? s_own : s_send) + (uint32_t)(gate * 8 * 512 + (unit & 31) * 16);
    }
    const bool writes = warp < 4 || lane < 16;
    const uint32_t t_src = (warp < 4 ? tmem_d1 : tmem_d2) + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t peer_buf = map_to_cta(sm_u32(s_peer), peer), peer_bar = map_to_cta(sm_u32(xbar), peer);
    // operand slot of (unit U, row b): [block = U / 64][row b][16-byte group ((U % 64) / 8) ^ (b & 7)][U % 8]
    auto h_slot = [&](int b) { return ((size_t)(U >> 6) * 32 + b) * 64 + (size_t)(((((U & 63) >> 3)) ^ (b & 7)) << 3) + (U & 7); };

    for (int s = 0; s < steps; ++s) {
      if (tid == 0 && up.trace_set != 2) UM_TRACE(s, 0);
      // input pre-activations of this step: independent of the matmul, in flight while it runs
      float gin[3][4];
      {
        const int t_in = jb.t_in0 + s * jb.t_in_step;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int b = b0 + j < B ? b0 + j : 0;
          const float* g0 = jb.gi + ((int64_t)t_in * B + b) * jb.ldg + U;
#pragma unroll
          for (int g = 0; g < 3; ++g) gin[g][j] = ldnc_f1(g0 + g * H);
        }
      }
      float acc[3][4];
#pragma unroll
      for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[g][j] = 0.0f;
      if (s > 0) {
        // every thread waits itself: one polling lane per warp + __syncwarp measured 1.7K cycles per step SLOWER (wake-up latency)
        mb_wait(acc_full, (uint32_t)(s - 1) & 1u);
        if (tid == 0 && up.trace_set != 2) UM_TRACE(s, 5);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        uint32_t v[32];
        tmem_ld32(t_src, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        if (tid == 0 && up.trace_set == 1) UM_TRACE(s, 1);
        if (writes) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(dst + q * 512), "r"(v[4 * q]), "r"(v[4 * q + 1]), "r"(v[4 * q + 2]), "r"(v[4 * q + 3]) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");      // s_send: generic stores -> the bulk copy's async-proxy reads
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        bar_sync(2, kUmEpiThreads);
        if (tid == 0 && up.trace_set == 1) UM_TRACE(s, 2);
        if (warp == 0) {
          if (elect_one()) {
            mb_expect_tx(xbar, kUmStgBytes);                                   // what the peer sends me this step
            bulk_s2peer(peer_buf, s_send, kUmStgBytes, peer_bar);              // what I send the peer
          }
          __syncwarp();
        }
        if (tid == 0 && up.trace_set == 1) UM_TRACE(s, 3);
        mb_wait(xbar, (uint32_t)(s - 1) & 1u);                                 // the peer's partial sums have landed in s_peer
        if (tid == 0 && up.trace_set != 2) UM_TRACE(s, 6);
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          const int o = ((g * 8 + warp) * 32 + lane) * 4;
          const float4 a = *reinterpret_cast<const float4*>(s_own + o), c = *reinterpret_cast<const float4*>(s_peer + o);
          acc[g][0] = a.x + c.x; acc[g][1] = a.y + c.y; acc[g][2] = a.z + c.z; acc[g][3] = a.w + c.w;
        }
      }
      float hn[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float r = um_sigmoid(gin[0][j] + (acc[0][j] + bh[0]));
        const float z = um_sigmoid(gin[1][j] + (acc[1][j] + bh[1]));
        const float n = um_tanh(gin[2][j] + r * (acc[2][j] + bh[2]));
        hn[j] = b0 + j < B ? (1.0f - z) * n + z * hp[j] : 0.0f;
        hp[j] = hn[j];
        if (jb.gates && b0 + j < B) {       // saved for tp_gru_cell_backward
          float* gt = jb.gates + ((int64_t)(jb.t_out0 + s * jb.t_out_step) * B + b0 + j) * 4 * H;
          gt[U] = r; gt[H + U] = z; gt[2 * H + U] = n; gt[3 * H + U] = acc[2][j] + bh[2];
        }
      }
      if (s + 1 < steps) {
        __nv_bfloat16* hw = hlp + (int64_t)(s & 1) * p.lp_slot;
#pragma unroll
        for (int j = 0; j < 4; ++j) hw[h_slot(b0 + j)] = __float2bfloat16_rn(hn[j]);
        // h_t of this CTA's units is written: arrive on the direction's grid barrier (release, cumulative over the
        // stores of the epilogue threads ordered by the named barrier); the caller-visible outputs follow AFTER the
        // arrival, off the critical path of the other CTAs
        if (tid == 0 && up.trace_set == 1) UM_TRACE(s, 4);
        if (up.sync_mode & 4) {
          asm volatile("fence.proxy.async.global;\n" ::: "memory");        // h_t words: generic stores -> bulk-copy (async proxy) reads of other CTAs
          asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // s_own / s_send live in the ring the next step's bulk copies overwrite
        }
        bar_sync(1, kUmEpiThreads);
        if (tid == 0) {
          if (up.trace_set != 2) UM_TRACE(s, 7);
          asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(counter) : "memory");
        }
      }
      const int t_out = jb.t_out0 + s * jb.t_out_step;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int b = b0 + j;
        if (b < B) {
          if (jb.y) jb.y[((int64_t)t_out * B + b) * jb.ldy + U] = hn[j];
          if (jb.y_lp) reinterpret_cast<__nv_bfloat16*>(jb.y_lp)[((int64_t)t_out * B + b) * jb.ldy_lp + U] = __float2bfloat16_rn(hn[j]);
          if (jb.h_final && s == steps - 1) jb.h_final[(int64_t)b * jb.ld_hf + U] = hn[j];
        }
      }
    }
    if (tid == 0 && steps > 0) UM_TRACE(steps - 1, 7);
  }
  // no CTA of the pair may exit while the other can still write into its shared memory
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  cluster_sync_all();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(geo.tmem_cols));
  }
}
