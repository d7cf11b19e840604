// tcgen05 / TMEM building blocks shared by the batched FC GEMM kernel (gemm_kernels.cu) and the tile-pipeline
// kernels (tile_kernels.cu): cp.async staging into the canonical K-major SWIZZLE_128B operand layout, the 3xTF32
// operand split, shared-memory descriptors, MMA issue, commit and TMEM loads.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gml {
namespace {

__device__ __forceinline__ void cp_async16(float* dst_smem, const float* src, int src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int UM = 128, UN = 128, UK = 32, USTAGES = 3, UGROUP = 4;
constexpr int U_PRODUCERS = 256, UTHREADS = U_PRODUCERS + 32;
constexpr int U_TILE_BYTES = UM * UK * 4;        // 16 KB
constexpr int U_STAGE_BYTES = 4 * U_TILE_BYTES;  // A big | A small | B big | B small
constexpr uint32_t U_TMEM_COLS = 256;            // two accumulators of 128 lanes x 128 fp32 columns

__device__ __forceinline__ uint32_t u_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void u_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(u_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void u_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(u_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void u_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spins = 0; !ok; ++spins) {  // bounded: a lost completion must trap, never hang the GPU
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(u_smem_addr(bar)), "r"(parity)
        : "memory");
    if (!ok && spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void u_commit(uint64_t* bar) {
  // arrives on the barrier once every MMA this thread has issued so far is complete
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(u_smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ uint64_t u_desc(uint32_t addr) {
  // start address [0,14) >> 4; leading byte offset [16,30) unused for swizzled K-major (1); stride byte offset
  // [32,46) = 1024 >> 4; descriptor version 1 at [46,48); layout type SWIZZLE_128B (2) at [61,64)
  return (uint64_t)((addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void u_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- warp-converged issue ------------------------------------------------------------------------------------------
// tcgen05.mma / commit / TMA take their operands from UNIFORM registers.  Measured on B200 (scripts/micro/umma_rate.cu):
// when the issuing code sits in a divergent region (`if (lane == 0) { ... }`) the compiler wraps every MMA in an
// ELECT / R2UR.BROADCAST / branch loop and one MMA issues per ~170 cycles whatever its shape; when the WHOLE warp
// runs the issue sequence converged and one lane is elected INSIDE the instruction's own asm block, the operand moves
// are software-pipelined and a 128x128x8 TF32 MMA issues every 64 cycles -- the tensor pipe's rate.
// All functions below must be called by all 32 lanes of a converged warp.
__device__ __forceinline__ void u_mma_tf32_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate));
}
__device__ __forceinline__ void u_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(u_smem_addr(bar))
      : "memory");
}
// the 12 MMAs (4 k-steps x {small*big, big*small, big*big}) of one 32-k operand stage at byte offset `stage_off`
__device__ __forceinline__ void u_mma_stage_elect(uint32_t smem_base, uint32_t stage_off, uint32_t acc, bool first,
                                                  uint32_t idesc) {
  const uint32_t a_big = smem_base + stage_off, a_small = a_big + U_TILE_BYTES;
  const uint32_t b_big = a_big + 2 * U_TILE_BYTES, b_small = a_big + 3 * U_TILE_BYTES;
#pragma unroll
  for (int j = 0; j < UK / 8; ++j) {
    const uint32_t o = (uint32_t)j * 32u;  // 8 k further inside the 128-byte rows
    u_mma_tf32_elect(acc, u_desc(a_small + o), u_desc(b_big + o), idesc, (j != 0 || !first) ? 1u : 0u);
    u_mma_tf32_elect(acc, u_desc(a_big + o), u_desc(b_small + o), idesc, 1u);
    u_mma_tf32_elect(acc, u_desc(a_big + o), u_desc(b_big + o), idesc, 1u);
  }
}

__device__ __forceinline__ void u_tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// byte offset of chunk kc of row t in the swizzled K-major tile
__device__ __forceinline__ uint32_t u_kmajor_off(int t, int kc) {
  return (uint32_t)((t >> 3) * 1024 + (t & 7) * 128 + ((kc ^ (t & 7)) << 4));
}
constexpr int U_MN_PITCH = UM + 4;  // floats per k row of an MN-major landing zone (16-byte aligned rows; a column
                                    // walk of 4 k then touches banks 4 apart -> conflict-free with the lane map below)
// K-major source: chunk i of a producer thread is row t, k-chunk kc; it lands at its final (swizzled) offset.
template <bool KC>
__device__ __forceinline__ void u_chunk(int i, int warp, int lane, int& t, int& kc) {
  const int wc = warp + 8 * i;
  if (KC) {  // 8 rows x 4 k-chunks per warp: 64-byte global runs, one swizzle atom row group per 8 lanes
    kc = (wc >> 4) * 4 + (lane >> 3);
    t = (wc & 15) * 8 + (lane & 7);
  } else {   // final chunks of an MN-major source: 16 rows x 2 k-chunks per warp (see u_read_chunks)
    kc = (wc >> 3) * 2 + (lane >> 4);
    t = (wc & 7) * 16 + (lane & 15);
  }
}
template <bool KC>
__device__ __forceinline__ void u_load_tile(unsigned char* tile, const float* __restrict__ src, int ld, int t0, int tmax,
                                            int k0, int kmax, int warp, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int bytes;
    const float* g;
    uint32_t land;
    if (KC) {
      int t, kc;
      u_chunk<true>(i, warp, lane, t, kc);
      land = u_kmajor_off(t, kc);
      t += t0;
      const int k = k0 + kc * 4;
      bytes = (t < tmax) ? (kmax - k) * 4 : 0;
      g = src + (size_t)t * ld + k;
    } else {  // k row `kr`, 4 consecutive t: one warp covers a whole 512-byte row of the tile
      const int kr = warp + 8 * i, tc = lane;
      land = (uint32_t)((kr * U_MN_PITCH + tc * 4) * 4);
      const int k = k0 + kr, t = t0 + tc * 4;
      bytes = (k < kmax) ? (tmax - t) * 4 : 0;
      g = src + (size_t)k * ld + t;
    }
    bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
    cp_async16(reinterpret_cast<float*>(tile + land), bytes > 0 ? g : src, bytes);
  }
}
// K-major: my own chunks back from where cp.async put them.  MN-major: the landing zone holds [k][t]; chunk
// (t, kc) is the column walk k = 4 kc .. 4 kc + 3 at fixed t -- 16 lanes on 16 consecutive t, the two lane halves
// 4 k rows (16 banks) apart: conflict-free.  (Needs a barrier first: other threads loaded those rows.)
template <bool KC>
__device__ __forceinline__ void u_read_chunks(const unsigned char* tile, int warp, int lane, float4 (&x)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int t, kc;
    u_chunk<KC>(i, warp, lane, t, kc);
    if (KC) {
      x[i] = *reinterpret_cast<const float4*>(tile + u_kmajor_off(t, kc));
    } else {
      const float* col = reinterpret_cast<const float*>(tile) + (kc * 4) * U_MN_PITCH + t;
      x[i] = make_float4(col[0], col[U_MN_PITCH], col[2 * U_MN_PITCH], col[3 * U_MN_PITCH]);
    }
  }
}
// big (in place for K-major sources, where the raw chunk already sits at its final offset and the tensor core
// reads only the top 19 bits of each word, i.e. sees exactly x & ~0x1fff) and small = x - big (twin tile)
template <bool KC>
__device__ __forceinline__ void u_write_split(unsigned char* big_tile, unsigned char* small_tile, int warp, int lane,
                                              const float4 (&x)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int t, kc;
    u_chunk<KC>(i, warp, lane, t, kc);
    const uint32_t fin = u_kmajor_off(t, kc);
    const float4 v = x[i];
    float4 b, s;
    b.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); s.x = v.x - b.x;
    b.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); s.y = v.y - b.y;
    b.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); s.z = v.z - b.z;
    b.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); s.w = v.w - b.w;
    if (!KC) *reinterpret_cast<float4*>(big_tile + fin) = b;
    *reinterpret_cast<float4*>(small_tile + fin) = s;
  }
}

}  // namespace
}  // namespace gml
