// Micro-benchmark: cost of small-N tcgen05.mma (kind::f16, bf16 in / fp32 out) on one SM as a function of the A operand
// source (TMEM or shared memory), M (64 / 128), N, and the number of independent accumulators consecutive MMAs rotate over.
// Straight-line code: 32 MMAs per block, fully unrolled with compile-time operands (no index arithmetic between them),
// issued by one elected lane of a converged warp; 16 blocks per measurement.  Operands are zeros (timing only).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I tepose_b200/csrc -o scripts/micro/umma_lat scripts/micro/umma_lat.cu
#include "umma.cuh"
#include <cstdio>
#include <cstdlib>
using namespace tp;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

// MODE 0: TS M=128; 1: SS M=128; 2: SS M=64; 3: alternate TS M=128 / SS M=64 (k_gru_umma pattern, NACC chains each)
template <int MODE, int N, int NACC, int COMMIT = 0>
__global__ void __launch_bounds__(128, 1) k_lat(long long* out) {
  extern __shared__ __align__(1024) unsigned char raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[16];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(&bar2[i], 1); mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tm = slot;
  constexpr int M = MODE == 2 ? 64 : 128;
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  constexpr uint32_t idesc64 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
  if (threadIdx.x < 32) {
    const uint64_t da = umma_desc_sw128(smem_u32(smem));                 // A: 128 rows x 64 (16 KB)
    const uint64_t db = umma_desc_sw128(smem_u32(smem + 32768));         // B: up to 256 rows x 64 (32 KB)
    const uint32_t a_tm = tm + 384;                                      // A in TMEM: columns 384..
    long long t_issue = 0, t_done = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int blk = 0; blk < 16; ++blk) {
        if (elect_one()) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const uint32_t d = tm + (uint32_t)((i % NACC) * N);
            const uint32_t acc = 1u;
            const int ks = i & 3;
            if (MODE == 0) mma_ts(d, a_tm + (uint32_t)((i & 15) * 8), db + (uint64_t)(ks * 2), idesc, acc);
            else if (MODE == 1 || MODE == 2) umma_f16(d, da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), idesc, acc);
            else {
              if ((i & 1) == 0) mma_ts(tm + (uint32_t)(((i >> 1) % NACC) * N), a_tm + (uint32_t)((i & 15) * 8), db + (uint64_t)(ks * 2), idesc, acc);
              else umma_f16(tm + (uint32_t)((NACC + (i >> 1) % NACC) * N), da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), idesc64, acc);
            }
          }
          if (COMMIT) umma_commit(&bar2[blk]);
        }
        __syncwarp();
      }
      const long long t1 = clock64();
      if (elect_one()) umma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, (uint32_t)rep & 1u);
      const long long t2 = clock64();
      t_issue = t1 - t0; t_done = t2 - t0;
    }
    if (threadIdx.x == 0) { out[0] = t_issue; out[1] = t_done; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512u));
}

static long long* d_out;
template <int MODE, int N, int NACC, int COMMIT = 0>
void run(const char* name) {
  const int smem = 100 * 1024;
  cudaFuncSetAttribute(k_lat<MODE, N, NACC, COMMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_lat<MODE, N, NACC, COMMIT><<<1, 128, smem>>>(d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
  printf("%-12s N=%3d chains=%d: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA  (%s)\n", name, N, NACC, (double)h[0] / 512, (double)h[1] / 512, cudaGetErrorString(e));
}

int main() {
  cudaMalloc(&d_out, 64);
  run<0, 32, 1>("TS M128"); run<0, 32, 2>("TS M128"); run<0, 32, 4>("TS M128"); run<0, 32, 8>("TS M128");
  run<0, 16, 1>("TS M128"); run<0, 64, 1>("TS M128"); run<0, 64, 2>("TS M128"); run<0, 128, 1>("TS M128"); run<0, 128, 2>("TS M128"); run<0, 256, 1>("TS M128");
  run<1, 32, 1>("SS M128"); run<1, 32, 2>("SS M128"); run<1, 32, 4>("SS M128"); run<1, 128, 1>("SS M128"); run<1, 256, 1>("SS M128");
  run<2, 32, 1>("SS M64"); run<2, 32, 2>("SS M64"); run<2, 32, 4>("SS M64"); run<2, 8, 1>("SS M64"); run<2, 128, 1>("SS M64");
  run<3, 32, 1, 1>("mix+commit/32"); run<0, 32, 1, 1>("TS+commit/32");
  run<3, 32, 1>("TS128+SS64"); run<3, 32, 2>("TS128+SS64"); run<3, 32, 4>("TS128+SS64");
  return 0;
}
