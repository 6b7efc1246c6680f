// K1 -- GRU input projection on the 5th-gen tensor cores:  out = A . W^T + bias
// (replaces the input GEMM inside cuDNN's GRU / the x.W_ih^T half of torch.nn.GRU,
// reference lib/models/tepose.py:53-64,73,76).
//
//   * operands bf16, K-major, moved by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a
//     6-stage shared-memory ring guarded by mbarriers;
//   * one elected thread issues tcgen05.mma (cta_group::1, M=128, N=128, K=16) with the fp32
//     accumulator in TMEM (128 lanes x 128 columns);
//   * tcgen05.commit releases ring slots / signals the epilogue; four warps read the
//     accumulator back with tcgen05.ld (32 lanes x 32 columns per instruction), add the bias
//     and store fp32.
// One 128x128 output tile per CTA; tiles that share a W tile are adjacent in launch order so
// the W tile is fetched from HBM once and re-served from L2.
#include "common.cuh"
#include "umma.cuh"
#include <cuda.h>

namespace tp {

// Tile 128 x BN, BN = 192 or 128 (chosen per launch): the kernel is bound by the L2 -> SM operand stream (every k-step a
// CTA pulls (128 + BN) x 64 bf16), so a wider tile raises the MACs per byte; BN = 192 also turns the 2.92 waves of the
// B=32,T=16 input projection (432 tiles of 128 x 128) into 1.95 (288 tiles).  Skinny launches (one wave or less of
// 128-wide tiles, e.g. the B = 1 live window) stay at BN = 128: they stream weights and want more CTAs.
constexpr int TC_BM = 128, TC_BK = 64;
template <int BN, int BMH> struct TcCfg {                             // BMH: 128-row accumulators per CTA (1 or 2)
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256 && (BMH == 1 || BMH == 2), "UMMA N / epilogue chunking");
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;                   // one 128-row A block of a k-step: 16 KB
  static constexpr int STAGE_BYTES = BMH * A_BYTES + BN * TC_BK * 2;  // 56 KB at (192, 2), 32 KB at (128, 1)
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES;           // 3 / 6
  static constexpr int TMEM_COLS = BMH * BN <= 32 ? 32 : BMH * BN <= 64 ? 64 : BMH * BN <= 128 ? 128 : BMH * BN <= 256 ? 256 : 512;
  static constexpr int EPI_PITCH = BN + 4;                            // floats per staged row (= 4 mod 32)
  static_assert((size_t)TC_BM * EPI_PITCH * 4 <= (size_t)STAGES * STAGE_BYTES, "epilogue staging must fit in the ring");
  static_assert(BMH * BN <= 512, "TMEM has 512 columns");
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
};
constexpr int TC_MAX_SEGS = 8;
constexpr int TC_THREADS = 128;

struct TcParams {
  tp_gemm_seg seg[TC_MAX_SEGS];
  int tile_begin[TC_MAX_SEGS + 1];
  int halves[TC_MAX_SEGS];          // 128-row accumulators per tile of this segment (2: 256-row tiles, every W tile loaded half as often)
  int nseg, kblocks;
};

template <int BN, int BMH>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_gemm_bf16_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const TcParams p) {
  using Cfg = TcCfg<BN, BMH>;
  constexpr int TC_BN = BN, TC_STAGES = Cfg::STAGES, TC_STAGE_BYTES = Cfg::STAGE_BYTES, TC_TMEM_COLS = Cfg::TMEM_COLS, TC_EPI_PITCH = Cfg::EPI_PITCH;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // carve: [stages][A 16KB | W 16KB] (1024-aligned), then barriers
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + TC_STAGES;
  uint64_t* tmem_full_bar = empty_bar + TC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile decode
  int s = 0;
#pragma unroll
  for (int q = 1; q < TC_MAX_SEGS; ++q)
    if (q < p.nseg && (int)blockIdx.x >= p.tile_begin[q]) s = q;
  // static-index select: a dynamic index into the kernel parameters costs a constant-cache miss per field
  tp_gemm_seg sg = p.seg[0];
  int tile0 = p.tile_begin[0], halves = p.halves[0];
#pragma unroll
  for (int q = 1; q < TC_MAX_SEGS; ++q)
    if (q == s) { sg = p.seg[q]; tile0 = p.tile_begin[q]; halves = p.halves[q]; }
  if (BMH == 1) halves = 1;
  const int local = blockIdx.x - tile0;
  const int bm = TC_BM * halves;                        // rows of this tile: 128 or 256
  const int m_tiles = (sg.m_rows + bm - 1) / bm;
  const int n_tile = local / m_tiles, m_tile = local - n_tile * m_tiles;
  const int m0 = m_tile * bm, n0 = n_tile * TC_BN;      // segment-local

  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
  }
  // bias of this tile's 128 columns -> shared memory now, so the epilogue does not wait on global loads
  __shared__ float s_bias[TC_BN];
  for (int c = threadIdx.x; c < TC_BN; c += TC_THREADS) s_bias[c] = (sg.bias && n0 + c < sg.n_cols) ? sg.bias[n0 + c] : 0.0f;
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TC_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  // PDL: barriers, TMEM and the tensor maps were set up while the producer of A (the pack kernel) was still running
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (elect_one()) {  // ===== TMA producer =====
      for (int kb = 0; kb < p.kblocks; ++kb) {
        const int st = kb % TC_STAGES;
        const uint32_t ph = (kb / TC_STAGES) & 1;
        mbar_wait(&empty_bar[st], ph ^ 1);
        unsigned char* a_dst = smem + st * TC_STAGE_BYTES;
        unsigned char* w_dst = a_dst + BMH * Cfg::A_BYTES;
        mbar_expect_tx(&full_bar[st], (uint32_t)(halves * Cfg::A_BYTES + TC_BN * TC_BK * 2));
        tma_load_2d(a_dst, &map_a, &full_bar[st], kb * TC_BK, sg.m_start + m0);
        if (BMH == 2 && halves == 2) tma_load_2d(a_dst + Cfg::A_BYTES, &map_a, &full_bar[st], kb * TC_BK, sg.m_start + m0 + TC_BM);
        tma_load_2d(w_dst, &map_w, &full_bar[st], kb * TC_BK, sg.n_start + n0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {  // ===== MMA issuer =====
      for (int kb = 0; kb < p.kblocks; ++kb) {
        const int st = kb % TC_STAGES;
        const uint32_t ph = (kb / TC_STAGES) & 1;
        mbar_wait(&full_bar[st], ph);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint32_t a_addr = smem_u32(smem + st * TC_STAGE_BYTES);
        const uint32_t w_addr = a_addr + BMH * Cfg::A_BYTES;
        const uint64_t db = umma_desc_sw128(w_addr);
#pragma unroll
        for (int h = 0; h < BMH; ++h) {
          if (h < halves) {
            const uint64_t da = umma_desc_sw128(a_addr + h * Cfg::A_BYTES);
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {
              // advance 16 elements (32 B) along K inside the 128-byte swizzle row: +2 in 16-byte units;
              // accumulator h lives TC_BN TMEM columns further
              umma_f16(tmem_base + (uint32_t)(h * TC_BN), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), Cfg::IDESC, (kb | k) != 0 ? 1u : 0u);
            }
          }
        }
        umma_commit(&empty_bar[st]);   // slot reusable once these MMAs have read it
      }
      umma_commit(tmem_full_bar);      // accumulator complete
    }
    __syncwarp();
  }

  // ===== epilogue: all four warps, warp w owns TMEM lanes 32w..32w+31 =====
  mbar_wait(tmem_full_bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  // Accumulator rows live one per thread (TMEM lane), but a row-per-thread global store touches 32 different
  // cache lines per instruction.  The ring stages are idle now (every MMA has completed), so the tile is
  // transposed through them: thread = row writes its 128 values (pitch 132 floats: conflict-free 16-byte
  // accesses), then each warp streams ITS 32 rows out with one fully coalesced 512-byte store per row.
  float* stage = reinterpret_cast<float*>(smem) + (size_t)(warp * 32) * TC_EPI_PITCH;
  const bool vec_ok = ((sg.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(sg.out) & 15) == 0);
#pragma unroll 1
  for (int h = 0; h < halves; ++h) {
#pragma unroll 1
  for (int c0 = 0; c0 < TC_BN; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(h * TC_BN + c0), v);
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 bb = *reinterpret_cast<const float4*>(&s_bias[c0 + q * 4]);
      float4 o;
      o.x = __uint_as_float(v[q * 4 + 0]) + bb.x; o.y = __uint_as_float(v[q * 4 + 1]) + bb.y;
      o.z = __uint_as_float(v[q * 4 + 2]) + bb.z; o.w = __uint_as_float(v[q * 4 + 3]) + bb.w;
      *reinterpret_cast<float4*>(stage + (size_t)lane * TC_EPI_PITCH + c0 + q * 4) = o;
    }
  }
  __syncwarp();
#pragma unroll 2
  for (int r = 0; r < 32; ++r) {
    const int row = m0 + h * TC_BM + warp * 32 + r;      // segment-local output row
    if (row >= sg.m_rows) break;
#pragma unroll
    for (int cc = lane * 4; cc < TC_BN; cc += 128) {     // 512 contiguous bytes per warp store
      const int n = n0 + cc;
      const float4 o = *reinterpret_cast<const float4*>(stage + (size_t)r * TC_EPI_PITCH + cc);
      float* dst = sg.out + (int64_t)row * sg.ldc + n;
      if (n + 3 < sg.n_cols && vec_ok) {
        *reinterpret_cast<float4*>(dst) = o;
      } else {
        const float e[4] = {o.x, o.y, o.z, o.w};
        for (int i = 0; i < 4; ++i)
          if (n + i < sg.n_cols) dst[i] = e[i];
      }
    }
  }
  __syncwarp();                                          // the staging rows are rewritten by the next half
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"((uint32_t)TC_TMEM_COLS));
  }
}

// ---- host side: tensor maps through the driver entry point (no libcuda link dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

static int make_map(CUtensorMap* map, const void* base, int rows, int kp, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(TP_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)kp, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kp * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(TP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return TP_OK;
}

}  // namespace tp

using namespace tp;

extern "C" int tp_gemm_bf16_tc(const void* A, int a_rows, const void* W, int w_rows, int kp,
                               const tp_gemm_seg* segs, int nseg, void* stream) {
  TP_CHECK_ARG(A && W && segs, "tp_gemm_bf16_tc: null pointer");
  TP_CHECK_ARG(nseg >= 1 && nseg <= TC_MAX_SEGS, "tp_gemm_bf16_tc: nseg=%d out of range", nseg);
  TP_CHECK_ARG(kp > 0 && kp % TC_BK == 0, "tp_gemm_bf16_tc: kp=%d must be a positive multiple of %d", kp, TC_BK);
  TP_CHECK_ARG(a_rows > 0 && w_rows > 0, "tp_gemm_bf16_tc: empty operand");
  TP_CHECK_ARG(aligned16(A) && aligned16(W), "tp_gemm_bf16_tc: operands must be 16-byte aligned");
  int dev = 0, major = 0;
  TP_CUDA(cudaGetDevice(&dev));
  TP_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return fail(TP_ERR_UNSUPPORTED, "tp_gemm_bf16_tc needs an sm_100 device (found sm_%d)", major);
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.nseg = nseg;
  p.kblocks = kp / TC_BK;
  // tile shape: once the launch has more than one wave of 128 x 128 tiles it is bound by the L2 -> SM operand stream, so
  // tiles get wider (192) and, for segments whose row count is a multiple of 256, twice as tall (two TMEM accumulators
  // share every W tile); skinny launches stay at 128 x 128: they stream weights and want CTAs.
  int tiles128 = 0;
  for (int i = 0; i < nseg; ++i) tiles128 += (int)(ceil_div(segs[i].m_rows, TC_BM) * ceil_div(segs[i].n_cols, 128));
  static const int bn_env = getenv("TP_TC_BN") ? atoi(getenv("TP_TC_BN")) : 0;
  // 256-row tiles (two TMEM accumulators sharing each W tile) cost ring depth -- 3 stages of 56 KB instead of 5 of 40 KB --
  // and measured slower on the B=32,T=16 input projection (55 us vs 47 us): opt-in only (TP_TC_TALL=1)
  static const bool tall = getenv("TP_TC_TALL") != nullptr;
  const int bn = bn_env == 128 || bn_env == 192 ? bn_env : (tiles128 > sm_count() ? 192 : 128);
  int tiles = 0;
  for (int i = 0; i < nseg; ++i) {
    const tp_gemm_seg& sg = segs[i];
    TP_CHECK_ARG(sg.out && sg.m_rows > 0 && sg.n_cols > 0, "tp_gemm_bf16_tc: segment %d is empty / has no output", i);
    TP_CHECK_ARG(sg.m_start >= 0 && sg.m_start + sg.m_rows <= a_rows, "tp_gemm_bf16_tc: segment %d rows out of range", i);
    TP_CHECK_ARG(sg.n_start >= 0 && sg.n_start + sg.n_cols <= w_rows, "tp_gemm_bf16_tc: segment %d cols out of range", i);
    p.seg[i] = sg;
    p.halves[i] = (bn == 192 && tall && sg.m_rows % (2 * TC_BM) == 0) ? 2 : 1;
    p.tile_begin[i] = tiles;
    tiles += (int)(ceil_div(sg.m_rows, TC_BM * p.halves[i]) * ceil_div(sg.n_cols, bn));
  }
  for (int i = nseg; i <= TC_MAX_SEGS; ++i) p.tile_begin[i] = tiles;
  CUtensorMap map_a, map_w;
  int rc = make_map(&map_a, A, a_rows, kp, TC_BM);
  if (rc != TP_OK) return rc;
  rc = make_map(&map_w, W, w_rows, kp, bn);
  if (rc != TP_OK) return rc;
  if (bn == 192 && tall) {
    using Cfg = TcCfg<192, 2>;
    const size_t smem = (size_t)Cfg::STAGES * Cfg::STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    TP_CUDA(cudaFuncSetAttribute(k_gemm_bf16_tc<192, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PdlConfig lc(dim3((unsigned)tiles), dim3(TC_THREADS), smem, (cudaStream_t)stream);
    TP_CUDA(cudaLaunchKernelEx(&lc.cfg, k_gemm_bf16_tc<192, 2>, map_a, map_w, p));
  } else if (bn == 192) {
    using Cfg = TcCfg<192, 1>;
    const size_t smem = (size_t)Cfg::STAGES * Cfg::STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    TP_CUDA(cudaFuncSetAttribute(k_gemm_bf16_tc<192, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PdlConfig lc(dim3((unsigned)tiles), dim3(TC_THREADS), smem, (cudaStream_t)stream);
    TP_CUDA(cudaLaunchKernelEx(&lc.cfg, k_gemm_bf16_tc<192, 1>, map_a, map_w, p));
  } else {
    using Cfg = TcCfg<128, 1>;
    const size_t smem = (size_t)Cfg::STAGES * Cfg::STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    TP_CUDA(cudaFuncSetAttribute(k_gemm_bf16_tc<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PdlConfig lc(dim3((unsigned)tiles), dim3(TC_THREADS), smem, (cudaStream_t)stream);
    TP_CUDA(cudaLaunchKernelEx(&lc.cfg, k_gemm_bf16_tc<128, 1>, map_a, map_w, p));
  }
  TP_LAUNCH_CHECK();
  return TP_OK;
}
