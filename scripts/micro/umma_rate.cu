// Microbenchmark: issue rate / execution time of tcgen05.mma (SS mode, K-major SWIZZLE_128B operands in shared
// memory) for kind::tf32 and kind::f16 (bf16) at 128xN tiles.  One CTA per SM; one thread issues `iters` MMAs back to
// back into one TMEM accumulator, commits, and waits.   nvcc -gencode arch=compute_100a,code=sm_100a umma_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t sa(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory"); return t; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred) : "r"(0xffffffffu));
  return pred != 0;
}

template <int KIND, int MODE>  // KIND 0 tf32, 1 bf16; MODE 0 one divergent thread, 1 warp-uniform loop + elect
__global__ void __launch_bounds__(128, 1) rate_kernel(int n, int iters, int distinct, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (sa(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x;
  for (int i = tid; i < 48 * 1024; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sa(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sa(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  if ((MODE != 1 && MODE != 3) ? tid == 0 : tid < 32) {
    const bool leader = (MODE != 1 && MODE != 3) ? true : (MODE == 3 ? tid == 0 : elect_one());
    // idesc: D fp32 (bit 4), A/B format at [7,10)/[10,13): tf32 = 2, bf16 = 1; N >> 3 at [17,23), M >> 4 at [24,29)
    const uint32_t fmt = KIND == 0 ? 2u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = sa(smem), b0 = sa(smem + 64 * 1024);
    const long long t0 = clk();
    if (MODE == 3) {
      const uint32_t ua0 = __shfl_sync(0xffffffffu, a0, 0), ub0 = __shfl_sync(0xffffffffu, b0, 0);
      const uint32_t utm = __shfl_sync(0xffffffffu, tmem, 0);
      for (int i = 0; i < iters; i += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t da = desc(ua0 + 32u * j), db = desc(ub0 + 32u * j);
          if (KIND == 0)
            asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(utm), "l"(da), "l"(db), "r"(idesc), "r"(1u));
          else
            asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(utm), "l"(da), "l"(db), "r"(idesc), "r"(1u));
        }
      }
    } else if (MODE == 2) {
      uint64_t dav[4], dbv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { dav[j] = desc(a0 + 32u * j); dbv[j] = desc(b0 + 32u * j); }
      for (int i = 0; i < iters; i += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (KIND == 0)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(dav[j]), "l"(dbv[j]), "r"(idesc), "r"(1u));
          else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(dav[j]), "l"(dbv[j]), "r"(idesc), "r"(1u));
        }
      }
    } else
    for (int i = 0; i < iters; ++i) {
      const uint32_t o = (uint32_t)(i % distinct) * 32u;  // k-step inside the 128-byte rows
      const uint64_t da = desc(a0 + (o & 96u) + ((o >> 7) * 16384u) % 49152u), db = desc(b0 + (o & 96u));
      if (!leader) continue;
      if (KIND == 0) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(i > 0 ? 1u : 0u) : "memory");
      } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(i > 0 ? 1u : 0u) : "memory");
      }
    }
    const long long t1 = clk();
    if (leader)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sa(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(sa(&bar)), "r"(0u) : "memory");
    }
    const long long t2 = clk();
    if (blockIdx.x == 0 && leader) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(rate_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(rate_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(rate_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(rate_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(rate_kernel<0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(rate_kernel<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(rate_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(rate_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2048;
  for (int mode = 0; mode < 4; ++mode) {
    const int grid = 148;
    for (int kind = 0; kind < 2; ++kind) {
      for (int n : {64, 128, 256}) {
        for (int distinct : {4}) {
          long long h[2];
          for (int rep = 0; rep < 2; ++rep) {
            if (kind == 0 && mode == 0) rate_kernel<0, 0><<<grid, 128, smem>>>(n, iters, distinct, d);
            else if (kind == 1 && mode == 0) rate_kernel<1, 0><<<grid, 128, smem>>>(n, iters, distinct, d);
            else if (kind == 0 && mode == 1) rate_kernel<0, 1><<<grid, 128, smem>>>(n, iters, distinct, d);
            else if (kind == 1 && mode == 1) rate_kernel<1, 1><<<grid, 128, smem>>>(n, iters, distinct, d);
            else if (kind == 0 && mode == 2) rate_kernel<0, 2><<<grid, 128, smem>>>(n, iters, distinct, d);
            else if (kind == 1 && mode == 2) rate_kernel<1, 2><<<grid, 128, smem>>>(n, iters, distinct, d);
            else if (kind == 0) rate_kernel<0, 3><<<grid, 128, smem>>>(n, iters, distinct, d);
            else rate_kernel<1, 3><<<grid, 128, smem>>>(n, iters, distinct, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          }
          cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          const double k = kind == 0 ? 8 : 16;
          printf("mode %d grid %3d %s 128x%3dx%2.0f distinct-k-steps %d: issue %.1f cyc/MMA, complete %.1f cyc/MMA -> %.0f FLOP/clk/SM\n",
                 mode, grid, kind == 0 ? "tf32" : "bf16", n, k, distinct, (double)h[0] / iters, (double)h[1] / iters,
                 2.0 * 128 * n * k * iters / (double)h[1]);
        }
      }
    }
  }
  return 0;
}
