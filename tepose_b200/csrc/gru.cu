// K2 -- GRU recurrence as ONE persistent cooperative kernel (replaces the cuDNN / ATen
// GRU time loop behind torch.nn.GRU, reference lib/models/tepose.py:53-64,73,76).
//
// Work decomposition: every job (= one direction of one layer) is cut into items of U
// hidden units; an item owns the 3U rows {r,z,n} x U of W_hh, computes
// gh = W_hh[rows,:] . h_prev^T for the whole batch tile, applies the gate math and publishes
// its U columns of h_t.  All items of all jobs advance one timestep, then the grid
// synchronises (h_t must be complete before anybody starts t+1).
//
//   fp32 : FFMA, W_hh / h_prev staged K-major through shared memory with cp.async.
//   bf16 : mma.sync m16n8k16 (batch <= 32 makes this a weight-streaming problem, not a
//          tensor-throughput one); W_hh fragments are loaded straight from global/L2 with
//          128-bit loads in a permuted-K order, h_prev (bf16) is staged once per step in
//          shared memory, 8 warps split K (x M) and reduce through shared memory.
//          State, gates and accumulation stay fp32.
#include "common.cuh"
#include "umma.cuh"
#include <stdlib.h>

namespace tp {

constexpr int kMaxJobs = 4;
constexpr int kHRep = 1;          // replicas of the bf16 state; >1 spreads readers over copies (measured: no gain, the
                                  // h broadcast is not L2-hot-spot bound), kept as a knob
constexpr int kGruThreads = 256;

struct GruParams {
  tp_gru_job jobs[kMaxJobs];
  int item_begin[kMaxJobs + 1];
  int njobs, B, H, U, max_steps, total_items, any_h0;
  int K;                    // reduction length of the recurrent matmul: H, or 3 H in TP_PRECISION_BF16X3 (operands [hi | lo | hi])
  int split3;               // TP_PRECISION_BF16X3: the bf16 operand copy of h_t is written as hi at column u, lo at H + u, hi at 2 H + u
  int n_item_jobs;          // jobs [0, n_item_jobs) are cut into items; the rest are step-0-only elementwise jobs
  float* hbuf;              // [njobs][2][B][H]  fp32 state ping-pong
  __nv_bfloat16* hbuf_lp;   // [kHRep][njobs][2][B][H]  bf16 copies (MMA operand of the next step)
  size_t lp_rep_stride;     // elements between replicas
  int64_t lp_slot;          // elements per (job, parity) bf16 state slot
  int lp_tiled;             // 1: bf16 state is stored [chunk = H/128][32 rows][128] with 16-byte group index XOR ((row & 1) << 2)
                            //    (one contiguous 8 KB block per ring stage of k_gru_bf16_tma); 0: row-major [B][H]
  unsigned int* barrier;    // monotonic grid-barrier counters (zeroed before launch), barrier_shards of them 128 B apart
  int barrier_shards;
  int dual_skew;            // k_gru_bf16_dual: SM cycles team 1 waits before its first step (de-phases the two teams)
  long long* trace;         // debug: [gridDim][max_steps][8] SM-clock stamps (tp_gru_set_trace), or null
};

#define TP_TRACE(slot)                                                                       \
  do {                                                                                       \
    if (p.trace && threadIdx.x == 0) p.trace[((size_t)blockIdx.x * p.max_steps + s) * 8 + (slot)] = clock64(); \
  } while (0)

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
// bf16 mode: the state is re-quantised to bf16 every step (2^-9 relative), so SFU-based
// exp / divide (~1e-6 absolute) are far below the mode's own rounding.
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 2.0f * sigmoid_fast(2.0f * x) - 1.0f; }

__device__ __forceinline__ void cp_async16_z(void* smem, const void* gmem, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ void mma_bf16(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// element index of (batch b, unit u) inside one [B,H] bf16 state slot
__device__ __forceinline__ int64_t lp_index(const GruParams& p, int b, int u) {
  if (!p.lp_tiled) return (int64_t)b * p.K + u;
  const int chunk = u >> 7, c = u & 127;
  return ((int64_t)chunk * 32 + b) * 128 + ((((c >> 3) ^ ((b & 1) << 2)) << 3) | (c & 7));
}

// Operands of the gate math that do not depend on this step's matmul; fetched early so their
// latency hides behind the W_hh stream.
struct GateIn { float gr, gz, gn, br, bz, bn, hp; };

__device__ __forceinline__ GateIn gate_fetch(const GruParams& p, const tp_gru_job& jb, int j, int s, int b, int u) {
  const int H = p.H, B = p.B;
  const int t_in = jb.t_in0 + s * jb.t_in_step;
  const float* gi = jb.gi + ((int64_t)t_in * B + b) * jb.ldg;
  GateIn g;
  // volatile asm loads: the compiler must not sink them below the (asm volatile) MMA loop -- the whole
  // point is that they are in flight while W_hh streams
  auto ldnc = [](const float* ptr) { float v; asm volatile("ld.global.nc.f32 %0, [%1];\n" : "=f"(v) : "l"(ptr)); return v; };
  g.gr = ldnc(gi + u); g.gz = ldnc(gi + H + u); g.gn = ldnc(gi + 2 * H + u);
  g.br = ldnc(jb.b_hh + u); g.bz = ldnc(jb.b_hh + H + u); g.bn = ldnc(jb.b_hh + 2 * H + u);
  g.hp = 0.0f;
  const bool have_prev = (s > 0) || (jb.h0 != nullptr);
  const float* hprev = p.hbuf + ((int64_t)(j * 2 + ((s + 1) & 1)) * B) * H;
  if (have_prev) asm volatile("ld.global.cg.f32 %0, [%1];\n" : "=f"(g.hp) : "l"(hprev + (int64_t)b * H + u));
  return g;
}

// Gate math of torch.nn.GRU for one (batch b, hidden unit u): acc_* are W_h* . h_prev.
template <bool FAST = false>
__device__ __forceinline__ void gru_finalize(const GruParams& p, const tp_gru_job& jb, int j, int s, int b, int u,
                                             const GateIn& g, float acc_r, float acc_z, float acc_n) {
  const int H = p.H, B = p.B;
  float r, z, n;
  if (FAST) {
    r = sigmoid_fast(g.gr + (acc_r + g.br));
    z = sigmoid_fast(g.gz + (acc_z + g.bz));
    n = tanh_fast(g.gn + r * (acc_n + g.bn));
  } else {
    r = sigmoidf_(g.gr + (acc_r + g.br));
    z = sigmoidf_(g.gz + (acc_z + g.bz));
    n = tanhf(g.gn + r * (acc_n + g.bn));
  }
  float h = (1.0f - z) * n + z * g.hp;
  const int64_t slot = ((int64_t)(j * 2 + (s & 1)) * B + b) * H + u;
  p.hbuf[slot] = h;
  {
    const __nv_bfloat16 hb = __float2bfloat16_rn(h);
    const int64_t base = (int64_t)(j * 2 + (s & 1)) * p.lp_slot;
    const int64_t lp = base + lp_index(p, b, u);
#pragma unroll
    for (int r = 0; r < kHRep; ++r) p.hbuf_lp[(size_t)r * p.lp_rep_stride + lp] = hb;
    if (p.split3) {
      const __nv_bfloat16 lo = __float2bfloat16_rn(h - __bfloat162float(hb));
      p.hbuf_lp[base + lp_index(p, b, H + u)] = lo;
      p.hbuf_lp[base + lp_index(p, b, 2 * H + u)] = hb;
    }
  }
  const int t_out = jb.t_out0 + s * jb.t_out_step;
  if (jb.y) jb.y[((int64_t)t_out * B + b) * jb.ldy + u] = h;
  if (jb.y_lp) reinterpret_cast<__nv_bfloat16*>(jb.y_lp)[((int64_t)t_out * B + b) * jb.ldy_lp + u] = __float2bfloat16_rn(h);
  if (jb.h_final && s == jb.steps - 1) jb.h_final[(int64_t)b * jb.ld_hf + u] = h;
  if (jb.gates) {                 // saved for tp_gru_cell_backward
    float* gt = jb.gates + ((int64_t)t_out * B + b) * 4 * H;
    gt[u] = r; gt[H + u] = z; gt[2 * H + u] = n; gt[3 * H + u] = acc_n + g.bn;
  }
}

__device__ __forceinline__ void locate_item(const GruParams& p, int item, int& j, int& u0) {
  j = 0;
#pragma unroll
  for (int q = 1; q < kMaxJobs; ++q)
    if (q < p.njobs && item >= p.item_begin[q]) j = q;
  u0 = (item - p.item_begin[j]) * p.U;
}

// h0 -> state slot 1 (the "previous" slot of step 0), fp32 and bf16.
__device__ void seed_h0(const GruParams& p) {
  const int64_t per = (int64_t)p.B * p.H;
  for (int j = 0; j < p.njobs; ++j) {
    const float* h0 = p.jobs[j].h0;
    if (!h0) continue;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < per; i += (int64_t)gridDim.x * blockDim.x) {
      float v = h0[i];
      p.hbuf[(int64_t)(j * 2 + 1) * per + i] = v;
      const int64_t lp = (int64_t)(j * 2 + 1) * p.lp_slot + lp_index(p, (int)(i / p.H), (int)(i % p.H));
#pragma unroll
      for (int r = 0; r < kHRep; ++r) p.hbuf_lp[(size_t)r * p.lp_rep_stride + lp] = __float2bfloat16_rn(v);
    }
  }
}

// ------------------------------------------------------------------------------------ fp32
// U = 32 units per item, batch tile 32: thread (u = tid/8, bq = tid%8) owns gates r,z,n of
// unit u for batches bq + 8i.  Rows are staged K-major (pitch 36 floats) so the float4 reads
// along K are conflict free.
__global__ void __launch_bounds__(kGruThreads, 1) k_gru_f32(const GruParams p) {
  constexpr int U = 32, BT = 32, BK = 32, LD = BK + 4, STAGES = 3;
  extern __shared__ __align__(16) float smem_f[];
  float* Ws = smem_f;                          // [STAGES][3U][LD]
  float* Hs = smem_f + STAGES * 3 * U * LD;    // [STAGES][BT][LD]
  const int tid = threadIdx.x, u = tid >> 3, bq = tid & 7;
  const int H = p.H, B = p.B;
  unsigned int epoch = 0;

  if (p.any_h0) { seed_h0(p); grid_barrier_sh(p.barrier, ++epoch, p.barrier_shards); }

  for (int s = 0; s < p.max_steps; ++s) {
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      int j, u0;
      locate_item(p, item, j, u0);
      const tp_gru_job& jb = p.jobs[j];
      if (s >= jb.steps) continue;
      const bool have_prev = (s > 0) || (jb.h0 != nullptr);
      const float* W = reinterpret_cast<const float*>(jb.w_hh);
      const float* hprev = p.hbuf + ((int64_t)(j * 2 + ((s + 1) & 1)) * B) * H;
      for (int b0 = 0; b0 < B; b0 += BT) {
        float acc[3][4];
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[g][i] = 0.0f;
        if (have_prev) {
          const int nkt = H / BK;
          auto load_stage = [&](int stage, int kt) {
            const int k0 = kt * BK;
            for (int c = tid; c < (3 * U + BT) * (BK / 4); c += kGruThreads) {
              int r = c >> 3, q = c & 7;
              if (r < 3 * U) {
                int g = r >> 5, uu = r & 31;
                cp_async16_z(&Ws[(stage * 3 * U + r) * LD + q * 4],
                             W + ((int64_t)g * H + u0 + uu) * H + k0 + q * 4, true);
              } else {
                int bb = r - 3 * U;
                bool ok = (b0 + bb) < B;
                cp_async16_z(&Hs[(stage * BT + bb) * LD + q * 4],
                             ok ? (hprev + (int64_t)(b0 + bb) * H + k0 + q * 4) : hprev, ok);
              }
            }
          };
#pragma unroll
          for (int st = 0; st < STAGES - 1; ++st) {
            if (st < nkt) load_stage(st, st);
            cp_commit();
          }
          for (int kt = 0; kt < nkt; ++kt) {
            cp_wait<STAGES - 2>();
            __syncthreads();
            int nk = kt + STAGES - 1;
            if (nk < nkt) load_stage(nk % STAGES, nk);
            cp_commit();
            const float* ws = Ws + (kt % STAGES) * 3 * U * LD;
            const float* hs = Hs + (kt % STAGES) * BT * LD;
#pragma unroll
            for (int q = 0; q < BK / 4; ++q) {
              float4 w[3], hv[4];
#pragma unroll
              for (int g = 0; g < 3; ++g) w[g] = *reinterpret_cast<const float4*>(&ws[(g * U + u) * LD + q * 4]);
#pragma unroll
              for (int i = 0; i < 4; ++i) hv[i] = *reinterpret_cast<const float4*>(&hs[(bq + 8 * i) * LD + q * 4]);
#pragma unroll
              for (int g = 0; g < 3; ++g)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  acc[g][i] = fmaf(w[g].x, hv[i].x, acc[g][i]);
                  acc[g][i] = fmaf(w[g].y, hv[i].y, acc[g][i]);
                  acc[g][i] = fmaf(w[g].z, hv[i].z, acc[g][i]);
                  acc[g][i] = fmaf(w[g].w, hv[i].w, acc[g][i]);
                }
            }
          }
          cp_wait<0>();
          __syncthreads();  // smem ring is reused by the next batch tile / item
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int b = b0 + bq + 8 * i;
          if (b < B) gru_finalize(p, jb, j, s, b, u0 + u, gate_fetch(p, jb, j, s, b, u0 + u), acc[0][i], acc[1][i], acc[2][i]);
        }
      }
    }
    if (s + 1 < p.max_steps) grid_barrier_sh(p.barrier, ++epoch, p.barrier_shards);
  }
}

// ------------------------------------------------------------------------------------ bf16
// NT = batch-tile / 8 (1 or 4), MG = U / 16 (1 or 2).  Warp w -> (kg = w / MG, mg = w % MG):
// K is split over KG = 8/MG warp groups, each warp owns 3 m-tiles (one per gate, 16 units).
template <int NT, int MG>
__global__ void __launch_bounds__(kGruThreads, 1) k_gru_bf16(const GruParams p) {
  constexpr int KG = 8 / MG, U = 16 * MG, NB = NT * 8, RP = U + 4;
  constexpr int PF = 4;                                       // W_hh blocks (32 columns) in flight per warp
  constexpr int GE = (NB * U + kGruThreads - 1) / kGruThreads; // gate elements per thread
  extern __shared__ __align__(16) unsigned char smem_b[];
  const int H = p.H, B = p.B;
  const int HP = H + 32;                                     // bf16 row pitch of the staged h
  __nv_bfloat16* hs = reinterpret_cast<__nv_bfloat16*>(smem_b);          // [NB][HP]
  float* red = reinterpret_cast<float*>(smem_b + (size_t)NB * HP * 2);   // [KG][3][NB][RP]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int kg = warp / MG, mg = warp % MG;
  const int nblk = H / 32;
  const int blk_lo = (kg * nblk) / KG, blk_hi = ((kg + 1) * nblk) / KG;
  unsigned int epoch = 0;

  if (p.any_h0) { seed_h0(p); grid_barrier_sh(p.barrier, ++epoch, p.barrier_shards); }

  // W_hh ring: buf[q][i] holds rows (g, g+8) of gate i for block (blk_lo + q mod PF)
  // fragment-packed W_hh (tp_pack_whh_bf16): w_a[q][i] / w_b[q][i] are the two k-subtiles of block q, gate i
  uint4 w_a[PF][3], w_b[PF][3];
  auto w_rows = [&](int j, int u0, const uint4* (&wrow)[3]) {
    const uint4* W = reinterpret_cast<const uint4*>(p.jobs[j].w_hh);
    const int64_t ut = (u0 >> 4) + mg;
#pragma unroll
    for (int i = 0; i < 3; ++i) wrow[i] = W + (((int64_t)i * (H / 16) + ut) * nblk) * 64 + lane;
  };
  auto w_prologue = [&](const uint4* (&wrow)[3]) {
#pragma unroll
    for (int q = 0; q < PF; ++q)
      if (blk_lo + q < blk_hi) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          w_a[q][i] = ldg_stream(wrow[i] + (int64_t)(blk_lo + q) * 64);
          w_b[q][i] = ldg_stream(wrow[i] + (int64_t)(blk_lo + q) * 64 + 32);
        }
      }
  };
  // single-item CTAs (the common case) start streaming the next step's W_hh before the barrier
  const bool single_item = p.total_items <= (int)gridDim.x;
  bool ring_primed = false;

  for (int s = 0; s < p.max_steps; ++s) {
    TP_TRACE(0);
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      int j, u0;
      locate_item(p, item, j, u0);
      const tp_gru_job& jb = p.jobs[j];
      if (s >= jb.steps) continue;
      const bool have_prev = (s > 0) || (jb.h0 != nullptr);
      const __nv_bfloat16* hprev = p.hbuf_lp + (int64_t)(j * 2 + ((s + 1) & 1)) * p.lp_slot;
      const uint4* wrow[3];
      w_rows(j, u0, wrow);
      for (int b0 = 0; b0 < B; b0 += NB) {
        if (have_prev) {
          // stage h_prev[b0 : b0+NB, K-slice of this warp group] (bf16): each K group copies and
          // consumes its own columns, so groups never wait for each other
          {
            const int c_lo = blk_lo * 4, c_n = (blk_hi - blk_lo) * 4;      // 16-byte chunks per row
            const int gt = mg * 32 + lane;                                  // thread index inside the K group
#pragma unroll 4
            for (int bb = 0; bb < NB; ++bb) {
              const bool ok = (b0 + bb) < B;
              for (int c = gt; c < c_n; c += MG * 32) {
                const int q = c_lo + c;
                cp_async16_z(hs + (size_t)bb * HP + q * 8, ok ? (hprev + (int64_t)(b0 + bb) * H + q * 8) : hprev, ok);
              }
            }
          }
          cp_commit();
          if (!ring_primed) w_prologue(wrow);
          ring_primed = false;
        }
        // gate operands of this thread's (batch, unit) elements: independent of the matmul
        GateIn gin[GE];
#pragma unroll
        for (int e = 0; e < GE; ++e) {
          const int idx = tid + e * kGruThreads;
          const int bb = idx / U, uu = idx - bb * U;
          if (idx < NB * U && b0 + bb < B) gin[e] = gate_fetch(p, jb, j, s, b0 + bb, u0 + uu);
        }
        if (have_prev) {
          float acc[3][NT][4];
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int n = 0; n < NT; ++n)
#pragma unroll
              for (int e = 0; e < 4; ++e) acc[i][n][e] = 0.0f;
          cp_wait<0>();
          if (MG == 1) __syncwarp();
          else asm volatile("bar.sync %0, %1;\n" ::"r"(1 + kg), "r"(MG * 32) : "memory");
          TP_TRACE(1);
          for (int blk = blk_lo; blk < blk_hi; blk += PF) {
#pragma unroll
            for (int q = 0; q < PF; ++q) {
              const int cur = blk + q;
              if (cur < blk_hi) {
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                  const uint4 bv = *reinterpret_cast<const uint4*>(hs + (size_t)(n * 8 + g) * HP + cur * 32 + 8 * t);
#pragma unroll
                  for (int i = 0; i < 3; ++i) {
                    mma_bf16(acc[i][n], w_a[q][i].x, w_a[q][i].y, w_a[q][i].z, w_a[q][i].w, bv.x, bv.y);
                    mma_bf16(acc[i][n], w_b[q][i].x, w_b[q][i].y, w_b[q][i].z, w_b[q][i].w, bv.z, bv.w);
                  }
                }
                if (cur + PF < blk_hi) {
#pragma unroll
                  for (int i = 0; i < 3; ++i) {
                    w_a[q][i] = ldg_stream(wrow[i] + (int64_t)(cur + PF) * 64);
                    w_b[q][i] = ldg_stream(wrow[i] + (int64_t)(cur + PF) * 64 + 32);
                  }
                }
              }
            }
          }
          TP_TRACE(2);
          // partial sums -> red[kg][gate][n][unit]
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
          __syncthreads();
          TP_TRACE(3);
        }
#pragma unroll
        for (int e = 0; e < GE; ++e) {
          const int idx = tid + e * kGruThreads;
          const int bb = idx / U, uu = idx - bb * U;
          if (idx >= NB * U || b0 + bb >= B) continue;
          float ar = 0.f, az = 0.f, an = 0.f;
          if (have_prev) {
#pragma unroll
            for (int k = 0; k < KG; ++k) {
              ar += red[((size_t)(k * 3 + 0) * NB + bb) * RP + uu];
              az += red[((size_t)(k * 3 + 1) * NB + bb) * RP + uu];
              an += red[((size_t)(k * 3 + 2) * NB + bb) * RP + uu];
            }
          }
          gru_finalize(p, jb, j, s, b0 + bb, u0 + uu, gin[e], ar, az, an);
        }
        __syncthreads();  // hs / red are reused by the next batch tile / item
      }
    }
    if (s + 1 < p.max_steps) {
      if (single_item && (int)blockIdx.x < p.total_items && B <= NB) {
        int j, u0;
        locate_item(p, blockIdx.x, j, u0);
        if (s + 1 < p.jobs[j].steps) {      // W_hh does not depend on h: fetch across the barrier
          const uint4* wrow[3];
          w_rows(j, u0, wrow);
          w_prologue(wrow);
          ring_primed = true;
        }
      }
      TP_TRACE(4);
      grid_barrier_sh(p.barrier, ++epoch, p.barrier_shards);
      TP_TRACE(5);
    }
  }
}

#include "gru_tma.inl"
#include "gru_res.inl"
#include "gru_dual.inl"
#include "gru_umma.inl"

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace tp

using namespace tp;

extern "C" int tp_pack_whh_bf16(const float* w_hh, void* dst, int H, void* stream) {
  TP_CHECK_ARG(w_hh && dst && H >= 32 && H % 32 == 0, "tp_pack_whh_bf16: need non-null pointers and H %% 32 == 0 (H=%d)", H);
  // [3H,H] row-major: the generic fragment packing with 16-row tiles ordered gate-major is exactly
  // the [gate][unit_tile][k block][q][lane][word] order k_gru_bf16 streams.
  return tp_pack_mma_a_bf16(w_hh, H, 3 * H, H, dst, stream);
}

extern "C" size_t tp_whh_umma_bytes(int H) { return H >= 128 && H % 128 == 0 ? (size_t)6 * H * H : 0; }

extern "C" int tp_pack_whh_umma(const float* w_hh, void* dst, int H, void* stream) {
  TP_CHECK_ARG(w_hh && dst && H >= 128 && H % 128 == 0, "tp_pack_whh_umma: need non-null pointers and H %% 128 == 0 (H=%d)", H);
  TP_CHECK_ARG(aligned16(w_hh) && aligned16(dst), "tp_pack_whh_umma: pointers must be 16-byte aligned");
  const size_t chunks = (size_t)6 * H * H / 16;
  const unsigned grid = (unsigned)(chunks / 256 < 4096 ? (chunks + 255) / 256 : 4096);
  k_pack_whh_umma<<<grid, 256, 0, (cudaStream_t)stream>>>(w_hh, reinterpret_cast<uint4*>(dst), H, um_cfg().max_tmem_k);
  TP_LAUNCH_CHECK();
  return TP_OK;
}

extern "C" void tp_gru_set_trace(void* device_buffer) { tp::set_trace_ptr(reinterpret_cast<long long*>(device_buffer)); }

extern "C" size_t tp_gru_workspace_bytes(int njobs, int B, int H) {
  size_t per = (size_t)njobs * 2 * B * H;
  size_t per_lp = (size_t)njobs * 2 * (B < 32 ? 32 : B) * 3 * H;  // the tiled layout pads a slot to 32 rows; x3: the [hi | lo | hi] operand of TP_PRECISION_BF16X3
  return 256 + align_up(per * sizeof(float), 256) + kHRep * align_up(per_lp * sizeof(__nv_bfloat16), 256);
}

template <typename KernelT>
static int launch_coop(KernelT kfn, const GruParams& p, size_t smem, cudaStream_t st) {
  TP_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  TP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, kGruThreads, smem));
  if (per_sm < 1) return fail(TP_ERR_UNSUPPORTED, "tp_gru_recurrence: kernel does not fit on an SM (smem=%zu)", smem);
  int grid = p.total_items < sm_count() ? p.total_items : sm_count();
  void* args[] = {(void*)&p};
  TP_CUDA(cudaLaunchCooperativeKernel((const void*)kfn, dim3(grid), dim3(kGruThreads), args, smem, st));
  count_launch();
  return TP_OK;
}

template <typename KernelT>
static int launch_coop_n(KernelT kfn, const GruParams& p, int stages, int threads, int grid, size_t smem, cudaStream_t st) {
  TP_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  TP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, threads, smem));
  if (per_sm < 1) return fail(TP_ERR_UNSUPPORTED, "tp_gru_recurrence: kernel does not fit on an SM (smem=%zu)", smem);
  PdlConfig lc(dim3(grid), dim3(threads), smem, st, /*cooperative=*/true);
  TP_CUDA(cudaLaunchKernelEx(&lc.cfg, kfn, p, stages));
  count_launch();
  return TP_OK;
}

template <typename KernelT>
static int launch_coop_n2(KernelT kfn, const GruParams& p, int a0, int a1, int threads, int grid, size_t smem, cudaStream_t st) {
  TP_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  TP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, threads, smem));
  if (per_sm < 1) return fail(TP_ERR_UNSUPPORTED, "tp_gru_recurrence: kernel does not fit on an SM (smem=%zu)", smem);
  PdlConfig lc(dim3(grid), dim3(threads), smem, st, /*cooperative=*/true);
  TP_CUDA(cudaLaunchKernelEx(&lc.cfg, kfn, p, a0, a1));
  count_launch();
  return TP_OK;
}

static int gru_recurrence(const tp_gru_job* jobs_in, int njobs, int B, int H, int precision, void* workspace,
                          size_t workspace_bytes, void* barrier, void* stream);

extern "C" int tp_gru_recurrence(const tp_gru_job* jobs_in, int njobs, int B, int H, int precision,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  return gru_recurrence(jobs_in, njobs, B, H, precision, workspace, workspace_bytes, nullptr, stream);
}

extern "C" int tp_gru_recurrence_ex(const tp_gru_job* jobs_in, int njobs, int B, int H, int precision,
                                    void* workspace, size_t workspace_bytes, void* barrier, void* stream) {
  TP_CHECK_ARG(!barrier || (reinterpret_cast<uintptr_t>(barrier) & 127) == 0, "tp_gru_recurrence_ex: barrier must be 128-byte aligned");
  return gru_recurrence(jobs_in, njobs, B, H, precision, workspace, workspace_bytes, barrier, stream);
}

static int gru_recurrence(const tp_gru_job* jobs_in, int njobs, int B, int H, int precision, void* workspace,
                          size_t workspace_bytes, void* barrier, void* stream) {
  TP_CHECK_ARG(jobs_in && njobs >= 1 && njobs <= kMaxJobs, "tp_gru_recurrence: njobs=%d out of range [1,%d]", njobs, kMaxJobs);
  TP_CHECK_ARG(B >= 1 && H >= 32 && H % 32 == 0, "tp_gru_recurrence: need B>=1 and H a multiple of 32 (B=%d H=%d)", B, H);
  TP_CHECK_ARG(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "tp_gru_recurrence: workspace must be 256-byte aligned");
  TP_CHECK_ARG(workspace_bytes >= tp_gru_workspace_bytes(njobs, B, H), "tp_gru_recurrence: workspace too small");
  TP_CHECK_ARG(precision == TP_PRECISION_FP32 || precision == TP_PRECISION_BF16 || precision == TP_PRECISION_BF16X3, "tp_gru_recurrence: bad precision");
  // jobs with a recurrent matmul first; single-step jobs from a zero state are pure gate math
  tp_gru_job jobs[kMaxJobs];
  int n_mat = 0;
  for (int j = 0; j < njobs; ++j)
    if (!(jobs_in[j].steps == 1 && jobs_in[j].h0 == nullptr)) jobs[n_mat++] = jobs_in[j];
  int n_all = n_mat;
  for (int j = 0; j < njobs; ++j)
    if (jobs_in[j].steps == 1 && jobs_in[j].h0 == nullptr) jobs[n_all++] = jobs_in[j];
  GruParams p;
  memset(&p, 0, sizeof(p));
  p.njobs = njobs; p.B = B; p.H = H; p.trace = trace_ptr();
  p.K = precision == TP_PRECISION_BF16X3 ? 3 * H : H;
  p.split3 = precision == TP_PRECISION_BF16X3 ? 1 : 0;
  for (int j = 0; j < njobs; ++j) {
    const tp_gru_job& jb = jobs[j];
    TP_CHECK_ARG(jb.gi && jb.w_hh && jb.b_hh && jb.steps >= 1, "tp_gru_recurrence: a job has null gi/w_hh/b_hh or steps<1");
    TP_CHECK_ARG(aligned16(jb.w_hh), "tp_gru_recurrence: w_hh must be 16-byte aligned");
    p.jobs[j] = jb;
    if (jb.steps > p.max_steps) p.max_steps = jb.steps;
    if (jb.h0) p.any_h0 = 1;
  }
  size_t per = (size_t)njobs * 2 * B * H;
  p.barrier = reinterpret_cast<unsigned int*>(barrier ? barrier : workspace);
  static const int shards_env = getenv("TP_BARRIER_SHARDS") ? atoi(getenv("TP_BARRIER_SHARDS")) : 1;   // measured: 8 shards 0.3020 ms vs 1 shard 0.2997 ms per step (same box) -- the arrivals are not the bottleneck
  p.barrier_shards = barrier ? (shards_env >= 1 && shards_env <= kBarrierShards ? shards_env : 1) : 1;   // a caller-provided slot is kBarrierShards x 128 zeroed bytes
  p.hbuf = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + 256);
  p.hbuf_lp = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<unsigned char*>(workspace) + 256 + align_up(per * sizeof(float), 256));
  const size_t per_lp = (size_t)njobs * 2 * (B < 32 ? 32 : B) * 3 * H;
  p.lp_rep_stride = align_up(per_lp * sizeof(__nv_bfloat16), 256) / sizeof(__nv_bfloat16);
  p.lp_slot = (int64_t)B * H; p.lp_tiled = 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (!barrier) TP_CUDA(cudaMemsetAsync(workspace, 0, 256, st));      // a caller-provided barrier word is already zero (and keeps
                                                                      // the memset node from sitting between this kernel and its PDL predecessor)
  const int sms = sm_count();

  // ---- bf16, W_hh resident in TMEM + shared memory, tcgen05 (gru_umma.inl): every matmul job has a tp_pack_whh_umma image
  static const bool no_umma = getenv("TP_GRU_NO_UMMA") != nullptr;
  if (precision == TP_PRECISION_BF16 && !no_umma && n_mat >= 1 && !p.any_h0 && H % 128 == 0 && B <= 32 && n_mat * (H / 32) <= sms) {
    bool have_img = true;
    for (int j = 0; j < n_mat; ++j) have_img = have_img && jobs[j].w_hh_umma != nullptr;
    const UmGeom geo = um_geom(H, um_cfg().max_tmem_k);
    int dev = 0, optin = 0;
    TP_CUDA(cudaGetDevice(&dev));
    TP_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    cudaFuncAttributes fa;
    TP_CUDA(cudaFuncGetAttributes(&fa, k_gru_umma));
    if (have_img && geo.smem_bytes + fa.sharedSizeBytes <= (size_t)optin) {
      UmParams up;
      memset(&up, 0, sizeof(up));
      p.n_item_jobs = n_mat;
      p.lp_tiled = 2; p.lp_slot = (int64_t)32 * H;
      up.g = p;
      up.max_tmem_k = um_cfg().max_tmem_k; up.sync_mode = um_cfg().sync_mode; up.trace_set = um_cfg().trace_set;
      for (int j = 0; j < n_mat; ++j) {
        TP_CHECK_ARG(aligned16(jobs[j].w_hh_umma), "tp_gru_recurrence: w_hh_umma must be 16-byte aligned");
        TP_CHECK_ARG(aligned16(jobs[j].gi) && jobs[j].ldg % 4 == 0, "tp_gru_recurrence: gi must be 16-byte aligned with ldg %% 4 == 0");
        up.w_img[j] = reinterpret_cast<const unsigned char*>(jobs[j].w_hh_umma);
      }
      TP_CUDA(cudaFuncSetAttribute(k_gru_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)geo.smem_bytes));
      cudaLaunchConfig_t cfg = cudaLaunchConfig_t{};
      cudaLaunchAttribute attr[3];
      cfg.gridDim = dim3((unsigned)(n_mat * (H / 32))); cfg.blockDim = dim3(kUmThreads);
      cfg.dynamicSmemBytes = geo.smem_bytes; cfg.stream = st;
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
      attr[2].id = cudaLaunchAttributeCooperative;
      attr[2].val.cooperative = 1;
      cfg.attrs = attr;
      // co-residency: every cluster of the grid must be resident at once (the kernel spins on grid barriers)
      cfg.numAttrs = 1;
      int max_clusters = 0;
      TP_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, k_gru_umma, &cfg));
      if (max_clusters * 2 >= (int)cfg.gridDim.x) {
        static int coop_ok = getenv("TP_UM_NOCOOP") ? 0 : -1;   // cooperative + cluster launches: probed once, plain cluster launch otherwise (TP_UM_NOCOOP: profilers)
        cfg.numAttrs = coop_ok == 0 ? 2 : 3;
        cudaError_t e = cudaLaunchKernelEx(&cfg, k_gru_umma, up);
        if (e != cudaSuccess && cfg.numAttrs == 3) {
          (void)cudaGetLastError();
          coop_ok = 0;
          cfg.numAttrs = 2;
          e = cudaLaunchKernelEx(&cfg, k_gru_umma, up);
        } else if (e == cudaSuccess && coop_ok < 0) {
          coop_ok = 1;
        }
        if (e != cudaSuccess) return fail(TP_ERR_CUDA, "k_gru_umma launch failed: %s", cudaGetErrorString(e));
        count_launch();
        return TP_OK;
      }
    }
  }

  // ---- bf16 fast path: TMA-fed ring, one item per CTA
  static const bool no_tma = getenv("TP_GRU_NO_TMA") != nullptr;
  // ---- two matmul jobs at batch 9..32 (the two causal directions of the encoder): interleaved teams, one CTA per 16 units
  static const bool no_dual = getenv("TP_GRU_NO_DUAL") != nullptr;
  if ((precision == TP_PRECISION_BF16 || precision == TP_PRECISION_BF16X3) && !no_tma && !no_dual && n_mat == 2 && !p.any_h0 && H % 128 == 0 &&
      B <= 32 && H / 16 <= sms) {
    static const int skew_env = getenv("TP_GRU_DUAL_SKEW") ? atoi(getenv("TP_GRU_DUAL_SKEW")) : 0;
    p.dual_skew = skew_env;
    p.U = 16; p.n_item_jobs = n_mat;
    p.total_items = H / 16;
    p.lp_tiled = 1; p.lp_slot = (int64_t)32 * p.K;
    auto launch = [&](auto kfn) -> int {
      TP_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDualSmem));
      int per_sm = 0;
      TP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, kDualThreads, kDualSmem));
      if (per_sm < 1) return fail(TP_ERR_UNSUPPORTED, "tp_gru_recurrence: dual kernel does not fit on an SM");
      PdlConfig lc(dim3(H / 16), dim3(kDualThreads), kDualSmem, st, /*cooperative=*/true);
      TP_CUDA(cudaLaunchKernelEx(&lc.cfg, kfn, p));
      count_launch();
      return TP_OK;
    };
    return B <= 8 ? launch(k_gru_bf16_dual<1>) : launch(k_gru_bf16_dual<4>);
  }
  if (precision == TP_PRECISION_BF16X3)
    return fail(TP_ERR_UNSUPPORTED, "tp_gru_recurrence(BF16X3): only exactly two matmul jobs without h0 at B <= 32, H %% 128 == 0 (the encoder's "
                                    "two full directions) are built; use TP_PRECISION_FP32 for this job set");
  if (precision == TP_PRECISION_BF16 && !no_tma && n_mat >= 1 && H % 128 == 0 && B <= 32 && n_mat * (H / 32) <= sms) {
    const int NB = B <= 8 ? 8 : 32;
    // resident-weight variant: part of each CTA's W_hh slice stays in registers / shared memory for all steps
    // Measured (profiles/r02_gru_resident_ab.txt): at NB = 8 it cuts the windowed live step by 7 %; at NB = 32 the
    // shared memory the resident chunks take is worth more as ring depth (streaming kernel 175 us, resident 189-203 us),
    // so NB = 32 stays on k_gru_bf16_tma unless TP_GRU_RES=1 forces the resident kernel.
    static const bool no_res = getenv("TP_GRU_NO_RES") != nullptr;
    static const bool force_res = getenv("TP_GRU_RES") != nullptr;
    if (!no_res && (NB == 8 || force_res)) {
      static const int ws_env = getenv("TP_GRU_WS") ? atoi(getenv("TP_GRU_WS")) : 0;
      static const int rs_env = getenv("TP_GRU_RS") ? atoi(getenv("TP_GRU_RS")) : -1;
      const int nchunks = H / 128;
      const int RR = NB == 8 ? kResRegChunks1 : kResRegChunks4;
      const int rr = RR < nchunks ? RR : nchunks;
      int ws = ws_env > 0 ? (ws_env > 8 ? 8 : ws_env) : 2;
      cudaFuncAttributes fa;
      if (NB == 8) TP_CUDA(cudaFuncGetAttributes(&fa, k_gru_bf16_res<1, kResRegChunks1>));
      else TP_CUDA(cudaFuncGetAttributes(&fa, k_gru_bf16_res<4, kResRegChunks4>));
      int dev = 0, optin = 0;
      TP_CUDA(cudaGetDevice(&dev));
      TP_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
      const long hring = (((long)kResHStages * NB * 256) + 1023) & ~1023L;
      const long avail = (long)optin - (long)fa.sharedSizeBytes - 512 - hring - (long)ws * kChunkBytes;
      int rs = avail > 0 ? (int)(avail / kChunkBytes) : 0;
      if (rs > nchunks - rr) rs = nchunks - rr;
      if (rs_env >= 0 && rs > rs_env) rs = rs_env;
      if (avail >= 0) {
        p.U = 32; p.n_item_jobs = n_mat;
        int items = 0;
        for (int j = 0; j < n_mat; ++j) { p.item_begin[j] = items; items += H / 32; }
        for (int j = n_mat; j <= kMaxJobs; ++j) p.item_begin[j] = items;
        p.total_items = items;
        p.lp_tiled = 1; p.lp_slot = (int64_t)32 * H;
        const size_t smem = (size_t)hring + (size_t)(rs + ws) * kChunkBytes + 512;
        if (NB == 8) return launch_coop_n2(k_gru_bf16_res<1, kResRegChunks1>, p, rs, ws, kTmaThreads, items, smem, st);
        return launch_coop_n2(k_gru_bf16_res<4, kResRegChunks4>, p, rs, ws, kTmaThreads, items, smem, st);
      }
    }
    const size_t region = ((size_t)4 * 3 * NB * 36 * 4 + 1023) & ~(size_t)1023;     // red only: h rides in the ring
    const size_t budget = 225 * 1024;
    int stages = (int)((budget - region - 256) / kStageBytes);
    if (stages > 8) stages = 8;
    if (stages > H / 128) stages = H / 128;
    if (stages >= 2 || (stages == 1 && H == 128)) {
      p.U = 32; p.n_item_jobs = n_mat;
      int items = 0;
      for (int j = 0; j < n_mat; ++j) { p.item_begin[j] = items; items += H / 32; }
      for (int j = n_mat; j <= kMaxJobs; ++j) p.item_begin[j] = items;
      p.total_items = items;
      const size_t smem = region + (size_t)stages * kStageBytes + 256;
      p.lp_tiled = 1; p.lp_slot = (int64_t)32 * H;
      if (NB == 8) return launch_coop_n(k_gru_bf16_tma<1>, p, stages, kTmaThreads, items, smem, st);
      return launch_coop_n(k_gru_bf16_tma<4>, p, stages, kTmaThreads, items, smem, st);
    }
  }

  // ---- generic paths: every job is cut into items
  int units_total = njobs * H;
  int U = 32;
  if (precision == TP_PRECISION_BF16 && units_total / 16 <= sms) U = 16;
  p.U = U; p.n_item_jobs = njobs;
  int items = 0;
  for (int j = 0; j < njobs; ++j) { p.item_begin[j] = items; items += H / U; }
  for (int j = njobs; j <= kMaxJobs; ++j) p.item_begin[j] = items;
  p.total_items = items;

  if (precision == TP_PRECISION_FP32) {
    size_t smem = (size_t)3 * (3 * 32 + 32) * 36 * sizeof(float);
    return launch_coop(k_gru_f32, p, smem, st);
  }
  const int NB = (B <= 8) ? 8 : 32;
  const int KG = (U == 16) ? 8 : 4;
  size_t smem = (size_t)NB * (H + 32) * 2 + (size_t)KG * 3 * NB * (U + 4) * sizeof(float);
  if (smem > 227 * 1024) return fail(TP_ERR_UNSUPPORTED, "tp_gru_recurrence(bf16): H=%d needs %zu B of shared memory", H, smem);
  if (NB == 8) {
    if (U == 16) return launch_coop(k_gru_bf16<1, 1>, p, smem, st);
    return launch_coop(k_gru_bf16<1, 2>, p, smem, st);
  }
  if (U == 16) return launch_coop(k_gru_bf16<4, 1>, p, smem, st);
  return launch_coop(k_gru_bf16<4, 2>, p, smem, st);
}
