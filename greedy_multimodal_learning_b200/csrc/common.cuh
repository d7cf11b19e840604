// Shared device/host helpers for libgml_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gml_b200.h"

namespace gml {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// ---- error plumbing ---------------------------------------------------------------------
void set_last_cuda_error(cudaError_t e, const char* what);

#define GML_CUDA_TRY(expr)                          \
  do {                                              \
    cudaError_t _e = (expr);                        \
    if (_e != cudaSuccess) {                        \
      ::gml::set_last_cuda_error(_e, #expr);        \
      return GML_E_CUDA;                            \
    }                                               \
  } while (0)

#define GML_LAUNCH_CHECK() GML_CUDA_TRY(cudaGetLastError())

// ---- launch accounting / optional per-kernel timing (capi.cu) ----------------------------
enum KernelTag {
  kTagMean = 0,   // plane means (forward pass 1)
  kTagDGate,      // <grad_out, input> plane dots (backward pass 1)
  kTagScaleFwd,   // gating pass
  kTagScaleBwd,   // gradient apply pass
  kTagGemm,       // squeeze/excitation FCs and their gradients
  kTagSmall,      // column sums, fills, running-mean update
  kTagFusedFwd,   // cluster/shared-memory-resident forward
  kTagFusedBwd,   // cluster/shared-memory-resident backward
  kTagSqnorm,     // multi-tensor sum of squares
  kTagStats,      // squeeze accumulation, accuracy counts
  kTagCount
};
// RAII around one kernel launch: counts it, and when profiling is on brackets it with events.
struct LaunchScope {
  LaunchScope(int tag, cudaStream_t st);
  ~LaunchScope();
  int tag_;
  cudaStream_t st_;
  void* rec_;
};

#define GML_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r != GML_OK) return _r; \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- device helpers -------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int WIDTH>
__device__ __forceinline__ float group_sum(float v) {  // reduce inside aligned groups of WIDTH lanes
#pragma unroll
  for (int o = WIDTH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming (read-once) 128-bit load: no L1 allocation, default L2 policy
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
// L2 eviction-priority policies.  On sm_100 the inline `.L2::evict_*` qualifiers exist only for
// 256-bit loads; 128-bit loads take a policy register through `.L2::cache_hint`.
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ldg_hint(const float4* p, uint64_t policy) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p), "l"(policy));
  return r;
}
__device__ __forceinline__ float ldg_hint_f32(const float* p, uint64_t policy) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(policy));
  return r;
}
// streaming store: written once, not re-read by this kernel
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// streaming store that also tells L2 to drop the line first (keeps evict_last lines of other streams resident)
__device__ __forceinline__ void stg_hint(float4* p, const float4& v, uint64_t policy) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w), "l"(policy)
               : "memory");
}

__device__ __forceinline__ float sigmoidf_ref(float x) { return 1.0f / (1.0f + expf(-x)); }
// MUFU.EX2 + MUFU.RCP: ~2 ulp of exp and 1 ulp of the reciprocal, i.e. a few 1e-7 relative on the gate (the parity
// tolerance is 1e-5); used where the sigmoid sits on a latency chain (GEMM epilogue of the tile pipeline)
__device__ __forceinline__ float sigmoidf_fast(float x) { return __frcp_rn(1.0f + __expf(-x)); }

}  // namespace gml
