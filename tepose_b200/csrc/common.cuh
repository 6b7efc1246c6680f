// Shared host/device helpers for the tepose_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/tepose_b200.h"

namespace tp {

// thread-local error text behind tp_last_error()
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define TP_CHECK_ARG(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) return ::tp::fail(TP_ERR_INVALID, __VA_ARGS__);  \
  } while (0)

#define TP_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (call);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return ::tp::fail(TP_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                               \
  } while (0)

#define TP_LAUNCH_CHECK()                                                                   \
  do {                                                                                      \
    ::tp::count_launch();                                                                   \
    cudaError_t _e = cudaPeekAtLastError();                                                 \
    if (_e != cudaSuccess)                                                                  \
      return ::tp::fail(TP_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                                \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

int sm_count();  // of the current device (cached)

bool pdl_enabled();            // tp_set_pdl(); default on
// launch configuration with programmatic stream serialization (and, optionally, a cooperative grid)
struct PdlConfig {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[2];
  PdlConfig(dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool cooperative = false) {
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.numAttrs = 1;
    if (cooperative) {
      attr[1].id = cudaLaunchAttributeCooperative;
      attr[1].val.cooperative = 1;
      cfg.numAttrs = 2;
    }
    cfg.attrs = attr;
  }
};
void count_launch();
long long* trace_ptr();          // debug stamp buffer (tp_gru_set_trace), or null
void set_trace_ptr(long long* p);  // bumps the counter behind tp_launch_count()


#ifdef __CUDACC__
// Programmatic dependent launch (PDL).  A kernel launched with pdl_config() may be scheduled as soon as every CTA of
// its predecessor in the stream has called pdl_launch_dependents() (or exited); it must call pdl_wait() before it
// touches anything the predecessor wrote (the wait returns when the predecessor grid has completed and flushed).
// Both are no-ops for kernels launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }

// Grid barrier for a co-resident (cooperatively launched) grid: one arrival per CTA on a
// monotonic counter, release/acquire at gpu scope (the release is cumulative over the CTA's
// writes that thread 0 observed through bar.sync).  CONTRACT: after this barrier, data written by
// other CTAs must be read with L2-coherent loads (__ldcg / cp.async.bulk), never through L1.  State that crosses the barrier is read with
// L2-coherent loads (cp.async.cg / ld.global.cg), so no L1 invalidation is needed.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int seen;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(counter) : "memory");
    // relaxed polls (ld.acquire in the loop would invalidate the whole L1 -- CCTL.IVALL -- on every
    // iteration), then ONE acquire fence once the last arrival has been observed
    do {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];\n" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
#ifdef TP_BARRIER_ACQUIRE_FENCE
    // Formal acquire.  It compiles to MEMBAR + CCTL.IVALL, and the L1 invalidate stalls every load the
    // CTA issues next for ~4-5K cycles (measured, profiles/).  It is only needed when data produced by
    // other CTAs is read through L1; the kernels using this barrier read such data exclusively with
    // L2-coherent accesses (ld.global.cg -> LDG.STRONG.GPU, cp.async.bulk), so it is off by default.
    asm volatile("fence.acq_rel.gpu;\n" ::: "memory");
#endif
  }
  __syncthreads();
}

// Sharded variant: `shards` counters 128 bytes apart (base + 32 k); CTA c arrives on counter c % shards and lanes
// 0..shards-1 of warp 0 each poll one counter, so the arrivals are `shards` short chains of same-address atomics instead
// of one long one.  shards == 1 is grid_barrier().  epoch counts the barriers of this launch from 1.
constexpr int kBarrierShards = 8;
__device__ __forceinline__ void grid_barrier_sh(unsigned int* base, unsigned int epoch, int shards) {
  __syncthreads();
  if ((int)threadIdx.x < shards) {
    const unsigned int k = threadIdx.x;
    unsigned int* ctr = base + 32 * k;
    if (k == blockIdx.x % (unsigned)shards) asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(ctr) : "memory");
    const unsigned int pop = gridDim.x / (unsigned)shards + (k < gridDim.x % (unsigned)shards ? 1u : 0u);
    const unsigned int target = epoch * pop;
    unsigned int seen;
    do {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];\n" : "=r"(seen) : "l"(ctr) : "memory");
    } while (seen < target);
#ifdef TP_BARRIER_ACQUIRE_FENCE
    asm volatile("fence.acq_rel.gpu;\n" ::: "memory");
#endif
  }
  __syncthreads();
}

#endif

}  // namespace tp
