// Pipelined fp32 SGEMM for the batched squeeze/excitation FCs and their gradients.
//
// The FC problems of an MMTM block are small (M = batch <= ~1k, N, K in {128..1024}): with one
// 64x64 tile per CTA there are far fewer tiles than SMs and every tile walks a long, latency-bound
// K loop.  This kernel therefore
//   * splits K across CTAs (split-K) so that ~2 CTAs per SM are busy, and
//   * feeds each CTA through a 3-stage cp.async (LDGSTS) pipeline, operands kept in their natural
//     layout in shared memory (no register transpose), read back conflict-free as 128-bit vectors.
// Split-K partials go to a caller-provided workspace; the LAST CTA to finish a tile (ticket
// counter) adds the partials in split order -- a fixed order, so results are bit-reproducible --
// and applies the epilogue (bias / ReLU / sigmoid / ReLU-mask / accumulate).
// CUDA cores on purpose: 1e-5 fp32 parity with the reference's Linear layers (see fc_kernels.cu).
#include <cooperative_groups.h>

#include <type_traits>

#include "common.cuh"
#include "kernels.h"
#include "umma.cuh"

namespace cg = cooperative_groups;

namespace gml {

int g_gemm_big_tiles = 0;  // tunable "gemm_big_tiles"
int g_gemm_tf32x3 = 1;     // tunable "gemm_tf32x3": tensor-core 3xTF32 inner product (0 = CUDA-core FFMA)
int g_gemm_umma = 1;       // tunable "gemm_umma": tcgen05 128x128 kernel for the large problems (0 = never)
long long* g_gemm_trace = nullptr;  // debug: globaltimer stamps of CTA (0,0,0) of the tcgen05 kernel (device buffer, 64 slots)

namespace {

// Two tile shapes share one kernel template: 64x64 (4x4 outputs per thread, static shared memory) for the
// small problems and 128x128 (8x8 per thread, 60 KB dynamic shared memory) for batch >= 512.
constexpr int BK = 16, STAGES = 3, THREADS = 256;  // (5 stages measured: no gain)
constexpr int PITCH_KC = BK + 4;   // [row][k] rows of 16 floats, pitch 20 -> conflict-free 128-bit reads
template <int BT> struct TileGeom {
  static constexpr int kPitchMN = BT + 4;  // [k][row]
  static constexpr int kFloats = (BT * PITCH_KC > BK * (BT + 4)) ? BT * PITCH_KC : BK * (BT + 4);
};

struct PipeBatch {
  GemmDesc d[3];         // up to three independent problems per launch (blockIdx.z / splits)
  float* part[3];        // split-K partials per problem: [splits][m][n]
  unsigned int* tickets; // [count][tiles_m * tiles_n]
  int splits;
  int k_per_split;       // multiple of BK
  long long* trace;      // debug: phase stamps of CTA (0,0,0) of the tcgen05 kernel, or nullptr
};

// one operand tile for k in [k0, k0 + BK): KC = k-contiguous source (src[t * ld + k]), else src[k * ld + t]
template <bool KC, int BT, int PITCH_MN = TileGeom<BT>::kPitchMN>
__device__ __forceinline__ void load_tile(float* s, const float* __restrict__ src, int ld, int t0, int tmax, int k0,
                                          int kmax, int tid) {
  constexpr int kChunks = BT * BK / 4;  // 16-byte chunks per tile
#pragma unroll
  for (int c = tid; c < kChunks; c += THREADS) {
    if (KC) {
      const int row = c >> 2, kc = (c & 3) * 4;
      const int t = t0 + row, k = k0 + kc;
      int bytes = (t < tmax) ? (kmax - k) * 4 : 0;
      bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
      const float* g = bytes > 0 ? src + (size_t)t * ld + k : src;
      cp_async16(s + row * PITCH_KC + kc, g, bytes);
    } else {
      const int kk = c / (BT / 4), tq = (c % (BT / 4)) * 4;
      const int k = k0 + kk, t = t0 + tq;
      int bytes = (k < kmax) ? (tmax - t) * 4 : 0;
      bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
      const float* g = bytes > 0 ? src + (size_t)k * ld + t : src;
      cp_async16(s + kk * PITCH_MN + tq, g, bytes);
    }
  }
}

// row of the tile that output slot i of thread-coordinate t_idx maps to
template <bool KC, int TM>
__device__ __forceinline__ int tile_row(int t_idx, int i) {
  return KC ? t_idx + 16 * i : (i >> 2) * 64 + t_idx * 4 + (i & 3);
}
// TM (rows of this thread) x 4 (k) block of an operand tile
template <bool KC, int TM>
__device__ __forceinline__ void frag(const float* s, int t_idx, int kk, float (&v)[TM][4]) {
  if (KC) {  // rows t_idx + 16 i, vector along k
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const float4 t = *reinterpret_cast<const float4*>(s + (t_idx + 16 * i) * PITCH_KC + kk);
      v[i][0] = t.x; v[i][1] = t.y; v[i][2] = t.z; v[i][3] = t.w;
    }
  } else {   // rows (h * 64 + 4 t_idx + i), vector along rows
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int h = 0; h < TM / 4; ++h) {
        const float4 t =
            *reinterpret_cast<const float4*>(s + (kk + q) * TileGeom<16 * TM>::kPitchMN + h * 64 + t_idx * 4);
        v[h * 4 + 0][q] = t.x; v[h * 4 + 1][q] = t.y; v[h * 4 + 2][q] = t.z; v[h * 4 + 3][q] = t.w;
      }
  }
}

// operand pointers / bounds for the k-tile starting at absolute k0 (second K segment: see GemmDesc)
struct KSeg { const float* a; const float* b; int lda, ldb, k0, kmax; };
__device__ __forceinline__ KSeg k_segment(const GemmDesc& d, int k0, int k_end) {
  if (d.k_split && k0 >= d.k_split) return KSeg{d.a2, d.b2, d.lda2, d.ldb2, k0 - d.k_split, k_end - d.k_split};
  return KSeg{d.a, d.b, d.lda, d.ldb, k0, d.k_split ? min(k_end, d.k_split) : k_end};
}

__device__ __forceinline__ bool aligned16_dev(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ float epilogue(const GemmDesc& d, float v, int m, int n, const float* cp) {
  if (d.beta) v += *cp;
  if (d.bias) v += d.bias[n];
  if (d.act == kActRelu) v = fmaxf(v, 0.f);
  else if (d.act == kActSigmoid) v = sigmoidf_ref(v);
  else if (d.act == kActReluMask) v = d.mask[(size_t)m * d.ldmask + n] > 0.f ? v : 0.f;
  return v;
}

template <bool A_KC, bool B_KC, int TM>
__global__ void __launch_bounds__(THREADS, TM == 4 ? 3 : 1) gemm_pipe_kernel(const PipeBatch pb) {
  constexpr int BM = 16 * TM, BN = 16 * TM;
  constexpr int TILE_FLOATS = TileGeom<BM>::kFloats;
  extern __shared__ __align__(16) float gemm_smem[];
  float (*As)[TILE_FLOATS] = reinterpret_cast<float (*)[TILE_FLOATS]>(gemm_smem);
  float (*Bs)[TILE_FLOATS] = reinterpret_cast<float (*)[TILE_FLOATS]>(gemm_smem + STAGES * TILE_FLOATS);
  __shared__ unsigned int s_ticket;
  const int prob = blockIdx.z / pb.splits, split = blockIdx.z - prob * pb.splits;
  const GemmDesc d = prob == 0 ? pb.d[0] : (prob == 1 ? pb.d[1] : pb.d[2]);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (m0 >= d.m || n0 >= d.n) return;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int k_begin = split * pb.k_per_split;
  const int k_end = min(d.k, k_begin + pb.k_per_split);
  const int nk = k_end > k_begin ? (k_end - k_begin + BK - 1) / BK : 0;

  float acc[TM][TM];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TM; ++j) acc[i][j] = 0.f;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) {
      const KSeg ks = k_segment(d, k_begin + s * BK, k_end);
      load_tile<A_KC, BM>(As[s], ks.a, ks.lda, m0, d.m, ks.k0, ks.kmax, tid);
      load_tile<B_KC, BN>(Bs[s], ks.b, ks.ldb, n0, d.n, ks.k0, ks.kmax, tid);
    }
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    const int nxt = kt + STAGES - 1;
    if (nxt < nk) {
      const KSeg ks = k_segment(d, k_begin + nxt * BK, k_end);
      load_tile<A_KC, BM>(As[nxt % STAGES], ks.a, ks.lda, m0, d.m, ks.k0, ks.kmax, tid);
      load_tile<B_KC, BN>(Bs[nxt % STAGES], ks.b, ks.ldb, n0, d.n, ks.k0, ks.kmax, tid);
    }
    cp_async_commit();
    const float* as = As[kt % STAGES];
    const float* bs = Bs[kt % STAGES];
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      float av[TM][4], bv[TM][4];
      frag<A_KC, TM>(as, ty, kk, av);
      frag<B_KC, TM>(bs, tx, kk, bv);
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TM; ++j) acc[i][j] = fmaf(av[i][q], bv[j][q], acc[i][j]);
    }
  }
  cp_async_wait<0>();

  auto row_of = [&](int i) { return m0 + tile_row<A_KC, TM>(ty, i); };
  auto col_of = [&](int j) { return n0 + tile_row<B_KC, TM>(tx, j); };

  if (pb.splits == 1) {
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int m = row_of(i);
      if (m >= d.m) continue;
#pragma unroll
      for (int j = 0; j < TM; ++j) {
        const int n = col_of(j);
        if (n >= d.n) continue;
        float* cp = d.c + (size_t)m * d.ldc + n;
        *cp = epilogue(d, acc[i][j], m, n, cp);
      }
    }
    return;
  }
  // split-K: publish my partial, the last CTA of this tile folds all partials in split order
  float* part = pb.part[prob];
  const size_t plane = (size_t)d.m * d.n;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = row_of(i);
    if (m >= d.m) continue;
#pragma unroll
    for (int j = 0; j < TM; ++j) {
      const int n = col_of(j);
      if (n < d.n) __stcg(part + (size_t)split * plane + (size_t)m * d.n + n, acc[i][j]);
    }
  }
  __threadfence();
  __syncthreads();
  unsigned int* ticket = pb.tickets + (size_t)prob * gridDim.x * gridDim.y + blockIdx.y * gridDim.x + blockIdx.x;
  if (tid == 0) s_ticket = atomicAdd(ticket, 1u);
  __syncthreads();
  if (s_ticket != (unsigned)pb.splits - 1) return;
  __threadfence();
  // fold the partials in split order; all 16 loads of a split are issued together (L2 latency, not
  // bandwidth, is the cost here), splits unrolled by two
  float sum[TM][TM];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TM; ++j) sum[i][j] = 0.f;
  int rows[TM], cols[TM];
#pragma unroll
  for (int i = 0; i < TM; ++i) { rows[i] = row_of(i); cols[i] = col_of(i); }
  for (int s0 = 0; s0 < pb.splits; s0 += 2) {
    float v0[TM][TM], v1[TM][TM];
    const bool two = s0 + 1 < pb.splits;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TM; ++j) {
        const bool ok = rows[i] < d.m && cols[j] < d.n;
        const size_t o = (size_t)rows[i] * d.n + cols[j];
        v0[i][j] = ok ? __ldcg(part + (size_t)s0 * plane + o) : 0.f;
        v1[i][j] = (ok && two) ? __ldcg(part + (size_t)(s0 + 1) * plane + o) : 0.f;
      }
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TM; ++j) sum[i][j] = (sum[i][j] + v0[i][j]) + v1[i][j];
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    if (rows[i] >= d.m) continue;
#pragma unroll
    for (int j = 0; j < TM; ++j) {
      if (cols[j] >= d.n) continue;
      float* cp = d.c + (size_t)rows[i] * d.ldc + cols[j];
      *cp = epilogue(d, sum[i][j], rows[i], cols[j], cp);
    }
  }
  if (tid == 0) *ticket = 0u;  // self-resetting: the workspace can be reused by the next launch
}


// -----------------------------------------------------------------------------------------------
// Tensor-core variant of the 64x64 kernel: same cp.async pipeline, split-K and epilogue, the inner
// product on mma.sync.m16n8k8 TF32 with the 3xTF32 split (x = big + small, both TF32;
// acc += small_a*big_b + big_a*small_b + big_a*big_b, fp32 accumulate).  The dropped small*small term is
// ~2^-22 relative, i.e. the result stays at fp32 SGEMM accuracy (parity tests: 1e-5), while the FMA-pipe
// ceiling of the CUDA-core kernel (~24 TFLOP/s, profiles/r1_experiments.md) no longer bounds the FC GEMMs.
// Warp w owns rows 16*(w&3).. and columns 32*(w>>2).. of the tile: one A fragment, four B fragments,
// 12 MMAs per k-step of 8.  [k][row] tiles use pitch 72 so that the fragment pattern (8t + g) is
// bank-conflict-free; [row][k] tiles keep pitch 20 (20g + t is conflict-free as well).
// -----------------------------------------------------------------------------------------------
constexpr int MPITCH_MN = 64 + 8;
constexpr int MMA_TILE_FLOATS = (64 * PITCH_KC > BK * MPITCH_MN) ? 64 * PITCH_KC : BK * MPITCH_MN;

template <bool KC>
__device__ __forceinline__ float tile_at(const float* s, int row, int k) {
  return KC ? s[row * PITCH_KC + k] : s[k * MPITCH_MN + row];
}
// x = big + small with big = x truncated to TF32 (explicit mask) and small = x - big, which is exact in fp32.
// `small` is handed to the tensor core as raw fp32 bits: the TF32 datapath reads only the top 19 bits, i.e.
// truncates it, leaving a residual below 2^-21 |x|.  (cvt.rna.tf32.f32 is emulated with 4 instructions on
// sm_100 -- 9 per split value; mask + subtract is 2.)
__device__ __forceinline__ void split_tf32(float x, uint32_t& big, uint32_t& small) {
  big = __float_as_uint(x) & 0xffffe000u;
  small = __float_as_uint(x - __uint_as_float(big));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(THREADS, 3) gemm_mma_kernel(const PipeBatch pb) {
  constexpr int BM = 64, BN = 64;
  extern __shared__ __align__(16) float gemm_smem[];
  float (*As)[MMA_TILE_FLOATS] = reinterpret_cast<float (*)[MMA_TILE_FLOATS]>(gemm_smem);
  float (*Bs)[MMA_TILE_FLOATS] = reinterpret_cast<float (*)[MMA_TILE_FLOATS]>(gemm_smem + STAGES * MMA_TILE_FLOATS);
  __shared__ unsigned int s_ticket;
  const int prob = blockIdx.z / pb.splits, split = blockIdx.z - prob * pb.splits;
  const GemmDesc d = prob == 0 ? pb.d[0] : (prob == 1 ? pb.d[1] : pb.d[2]);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (m0 >= d.m || n0 >= d.n) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wm = (warp & 3) * 16, wn = (warp >> 2) * 32;
  const int k_begin = split * pb.k_per_split;
  const int k_end = min(d.k, k_begin + pb.k_per_split);
  const int nk = k_end > k_begin ? (k_end - k_begin + BK - 1) / BK : 0;

  // [n-tile j][c register]: row g + 8 (r >> 1), column 8 j + 2 t + (r & 1).  The tensor core adds into its
  // accumulator with truncation, so `part` only ever holds ONE k-tile (16 k) and is then promoted into `acc`
  // with a round-to-nearest add: the long sum over K behaves like the CUDA-core kernel's.
  float acc[4][4], part[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[j][r] = 0.f;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) {
      const KSeg ks = k_segment(d, k_begin + s * BK, k_end);
      load_tile<A_KC, BM, MPITCH_MN>(As[s], ks.a, ks.lda, m0, d.m, ks.k0, ks.kmax, tid);
      load_tile<B_KC, BN, MPITCH_MN>(Bs[s], ks.b, ks.ldb, n0, d.n, ks.k0, ks.kmax, tid);
    }
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    const int nxt = kt + STAGES - 1;
    if (nxt < nk) {
      const KSeg ks = k_segment(d, k_begin + nxt * BK, k_end);
      load_tile<A_KC, BM, MPITCH_MN>(As[nxt % STAGES], ks.a, ks.lda, m0, d.m, ks.k0, ks.kmax, tid);
      load_tile<B_KC, BN, MPITCH_MN>(Bs[nxt % STAGES], ks.b, ks.ldb, n0, d.n, ks.k0, ks.kmax, tid);
    }
    cp_async_commit();
    const float* as = As[kt % STAGES];
    const float* bs = Bs[kt % STAGES];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) part[j][r] = 0.f;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 8) {
      uint32_t a_big[4], a_small[4];
      split_tf32(tile_at<A_KC>(as, wm + g, kk + t), a_big[0], a_small[0]);
      split_tf32(tile_at<A_KC>(as, wm + g + 8, kk + t), a_big[1], a_small[1]);
      split_tf32(tile_at<A_KC>(as, wm + g, kk + t + 4), a_big[2], a_small[2]);
      split_tf32(tile_at<A_KC>(as, wm + g + 8, kk + t + 4), a_big[3], a_small[3]);
      uint32_t b_big[4][2], b_small[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        split_tf32(tile_at<B_KC>(bs, wn + 8 * j + g, kk + t), b_big[j][0], b_small[j][0]);
        split_tf32(tile_at<B_KC>(bs, wn + 8 * j + g, kk + t + 4), b_big[j][1], b_small[j][1]);
      }
      // correction terms first, then the leading term; four independent accumulator chains per pass
#pragma unroll
      for (int j = 0; j < 4; ++j) mma_tf32(part[j], a_small, b_big[j]);
#pragma unroll
      for (int j = 0; j < 4; ++j) mma_tf32(part[j], a_big, b_small[j]);
#pragma unroll
      for (int j = 0; j < 4; ++j) mma_tf32(part[j], a_big, b_big[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[j][r] += part[j][r];
  }
  cp_async_wait<0>();

  auto row_of = [&](int r) { return m0 + wm + g + 8 * (r >> 1); };
  auto col_of = [&](int j, int r) { return n0 + wn + 8 * j + 2 * t + (r & 1); };

  if (pb.splits == 1) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int m = row_of(r), n = col_of(j, r);
        if (m >= d.m || n >= d.n) continue;
        float* cp = d.c + (size_t)m * d.ldc + n;
        *cp = epilogue(d, acc[j][r], m, n, cp);
      }
    return;
  }
  float* partials = pb.part[prob];
  const size_t plane = (size_t)d.m * d.n;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int m = row_of(r), n = col_of(j, r);
      if (m < d.m && n < d.n) __stcg(partials + (size_t)split * plane + (size_t)m * d.n + n, acc[j][r]);
    }
  __threadfence();
  __syncthreads();
  unsigned int* ticket = pb.tickets + (size_t)prob * gridDim.x * gridDim.y + blockIdx.y * gridDim.x + blockIdx.x;
  if (tid == 0) s_ticket = atomicAdd(ticket, 1u);
  __syncthreads();
  if (s_ticket != (unsigned)pb.splits - 1) return;
  __threadfence();
  // last CTA of the tile: fold the partials in split order (fixed order -> reproducible), 16 loads in flight
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[j][r] = 0.f;
  for (int sp = 0; sp < pb.splits; ++sp) {
    float v[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int m = row_of(r), n = col_of(j, r);
        v[j][r] = (m < d.m && n < d.n) ? __ldcg(partials + (size_t)sp * plane + (size_t)m * d.n + n) : 0.f;
      }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[j][r] += v[j][r];
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int m = row_of(r), n = col_of(j, r);
      if (m >= d.m || n >= d.n) continue;
      float* cp = d.c + (size_t)m * d.ldc + n;
      *cp = epilogue(d, acc[j][r], m, n, cp);
    }
  if (tid == 0) *ticket = 0u;
}


// -----------------------------------------------------------------------------------------------
// tcgen05 (UMMA) kernel for the large FC problems: 128x128 output tile per CTA, accumulators in TMEM,
// operands in shared memory in the canonical K-major SWIZZLE_128B UMMA layout, MMAs issued by one thread.
//
// fp32 accuracy from TF32 tensor cores (3xTF32): every operand element is split once, in shared memory, into
// big = x & ~0x1fff (exactly a TF32 number) and small = x - big (exact in fp32; the TF32 datapath reads its
// top 19 bits); per k-step of 8 the issuing thread queues small_a*big_b, big_a*small_b, big_a*big_b into the
// same TMEM accumulator.  The tensor core adds into its accumulator with truncation, so a chain is kept short:
// the K loop runs in groups of 128 k that alternate between two TMEM accumulators; while one group is being
// multiplied the previous one is drained (tcgen05.ld) into per-thread fp32 sums with round-to-nearest adds.
//
// Roles (9 warps): warps 0-7 load (cp.async, 16-byte chunks), split, drain and run the epilogue; warp 8 issues
// the MMAs.  Hand-offs are mbarriers: full[s] (256 producer arrivals) -> MMA; empty[s] (tcgen05.commit) ->
// producers; acc_full[b] (tcgen05.commit) -> drainers; acc_empty[b] (256 arrivals) -> MMA.
//
// Shared-memory operand tile (128 rows x 32 k, 16 KB): an atom is 8 rows of 128 bytes on a 1024-byte boundary;
// the 16-byte chunk kc of row t sits at 1024 (t / 8) + 128 (t % 8) + 16 (kc ^ (t % 8)) -- the XOR swizzle the
// tensor core undoes.  Descriptor per k-step j: start + 32 j, SBO (8-row group stride) = 1024.
//   * K-major source (src[t * ld + k]): cp.async writes each chunk straight to its final place; the thread
//     that loaded it rewrites it as big in place and writes small to the twin tile.
//   * MN-major source (src[k * ld + t]; the tcgen05 transpose path returned zeros for TF32 here): a chunk is
//     4 t of one k.  It lands in a thread-private slot, is read back to registers, transposed 4x4 across the
//     four lanes that hold k..k+3 (two shuffle rounds) into K-major chunks, split and stored.
// Every lane group of 8 touches 8 different bank groups and every global read is a 64-byte run.
// -----------------------------------------------------------------------------------------------
template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(UTHREADS, 1) gemm_umma_kernel(const PipeBatch pb) {
  extern __shared__ __align__(1024) unsigned char u_smem_raw[];
  // swizzle atoms must sit on 1024-byte boundaries of the shared-memory address space
  unsigned char* u_smem = u_smem_raw + ((1024u - (u_smem_addr(u_smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bar_full[USTAGES], bar_empty[USTAGES], bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t s_tmem;
    const int prob = blockIdx.z / pb.splits, split = blockIdx.z - prob * pb.splits;
  const GemmDesc d = prob == 0 ? pb.d[0] : (prob == 1 ? pb.d[1] : pb.d[2]);
  const int m0 = blockIdx.y * UM, n0 = blockIdx.x * UN;
  if (m0 >= d.m || n0 >= d.n) return;  // uniform for the CTA, before anything is allocated
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int k_begin = split * pb.k_per_split;
  const int k_end = min(d.k, k_begin + pb.k_per_split);
  const int nk = k_end > k_begin ? (k_end - k_begin + UK - 1) / UK : 0;
  const int ngroups = (nk + UGROUP - 1) / UGROUP;
  const bool tracing = pb.trace != nullptr && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
#define U_STAMP(i)                                                      \
  do {                                                                  \
    if (tracing) {                                                      \
      long long t_;                                                     \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));            \
      pb.trace[(i)] = t_;                                               \
    }                                                                   \
  } while (0)
  U_STAMP(0);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < USTAGES; ++s) {
      u_mbar_init(&bar_full[s], U_PRODUCERS);
      u_mbar_init(&bar_empty[s], 1);
    }
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      u_mbar_init(&bar_acc_full[b], 1);
      u_mbar_init(&bar_acc_empty[b], U_PRODUCERS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(u_smem_addr(&s_tmem)),
                 "r"(U_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  U_STAMP(1);

  float sum[64];  // producers: row 32 (warp & 3) + lane, columns 64 (warp >> 2) .. + 63 of the tile
#pragma unroll
  for (int i = 0; i < 64; ++i) sum[i] = 0.f;

  if (warp == U_PRODUCERS / 32) {
    // ===== MMA issuer: the whole warp converged, one lane elected inside each instruction (umma.cuh) ==========
    {
      // instruction descriptor: D fp32 [4,6)=1, A and B TF32 [7,10)=[10,13)=2, both K-major, N >> 3 at [17,23),
      // M >> 4 at [24,29)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(UN >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);
      const uint32_t smem_base = __shfl_sync(0xffffffffu, u_smem_addr(u_smem), 0);
      const uint32_t utmem = __shfl_sync(0xffffffffu, tmem, 0);
      const int unk = __shfl_sync(0xffffffffu, nk, 0);
      for (int kt = 0; kt < unk; ++kt) {
        const int stage = kt % USTAGES, g = kt / UGROUP, b = g & 1;
        if (kt % UGROUP == 0 && g >= 2) u_mbar_wait(&bar_acc_empty[b], (uint32_t)(((g >> 1) - 1) & 1));
        u_mbar_wait(&bar_full[stage], (uint32_t)((kt / USTAGES) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        u_mma_stage_elect(smem_base, (uint32_t)(stage * U_STAGE_BYTES), utmem + (uint32_t)(b * UN), kt % UGROUP == 0, idesc);
        u_commit_elect(&bar_empty[stage]);
        if (kt % UGROUP == UGROUP - 1 || kt == unk - 1) u_commit_elect(&bar_acc_full[b]);
      }
    }
    __syncwarp();
  } else {
    // ===== producers / drainers ===========================================================
    const uint32_t t_row = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * 64);
    int drained = 0;
    auto drain = [&](int gi) {
      const int b = gi & 1;
      u_mbar_wait(&bar_acc_full[b], (uint32_t)((gi >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v[16];
        u_tmem_ld16(t_row + (uint32_t)(b * UN + 16 * c), v);
#pragma unroll
        for (int i = 0; i < 16; ++i) sum[16 * c + i] += v[i];
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      u_mbar_arrive(&bar_acc_empty[b]);
    };
    auto issue_loads = [&](int kt) {
      unsigned char* st = u_smem + (kt % USTAGES) * U_STAGE_BYTES;
      const KSeg ks = k_segment(d, k_begin + kt * UK, k_end);
      u_load_tile<A_KC>(st, ks.a, ks.lda, m0, d.m, ks.k0, ks.kmax, warp, lane);
      u_load_tile<B_KC>(st + 2 * U_TILE_BYTES, ks.b, ks.ldb, n0, d.n, ks.k0, ks.kmax, warp, lane);
    };
#pragma unroll
    for (int s = 0; s < USTAGES - 1; ++s) {
      if (s < nk) issue_loads(s);
      cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
      const int nxt = kt + USTAGES - 1;
      if (nxt < nk) {
        // tile nxt reuses the stage of tile nxt - USTAGES: wait until the MMAs that read it are done
        if (nxt >= USTAGES) u_mbar_wait(&bar_empty[nxt % USTAGES], (uint32_t)(((nxt / USTAGES) - 1) & 1));
        issue_loads(nxt);
      }
      cp_async_commit();
      cp_async_wait<USTAGES - 1>();  // my chunks of tile kt have landed
      if (kt < 8) U_STAMP(2 + 4 * kt);
      unsigned char* st = u_smem + (kt % USTAGES) * U_STAGE_BYTES;
      float4 xa[4], xb[4];
      if (!A_KC || !B_KC) asm volatile("bar.sync 1, 256;" ::: "memory");  // every producer's rows of tile kt are in
      u_read_chunks<A_KC>(st, warp, lane, xa);
      u_read_chunks<B_KC>(st + 2 * U_TILE_BYTES, warp, lane, xb);
      if (!A_KC || !B_KC) asm volatile("bar.sync 1, 256;" ::: "memory");  // ... and read, before the zone is rewritten
      u_write_split<A_KC>(st, st + U_TILE_BYTES, warp, lane, xa);
      u_write_split<B_KC>(st + 2 * U_TILE_BYTES, st + 3 * U_TILE_BYTES, warp, lane, xb);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA unit
      u_mbar_arrive(&bar_full[kt % USTAGES]);
      if (kt < 8) U_STAMP(3 + 4 * kt);
      // drain a finished group once the MMA warp has the next group's first tile to chew on
      if (kt % UGROUP >= 1 && drained < kt / UGROUP) { drain(drained); ++drained; }
      if (kt < 8) U_STAMP(4 + 4 * kt);
    }
    cp_async_wait<0>();
    U_STAMP(40);
    while (drained < ngroups) { drain(drained); ++drained; }
    U_STAMP(41);
  }

  // ---- epilogue ------------------------------------------------------------------------------------------
  // Each producer thread holds one row x 64 columns of this CTA's partial tile.  Without split-K it is written
  // out directly.  With split-K the CTAs of a tile form a thread-block cluster (along z): every CTA parks its
  // partial tile in its own shared memory, and after a cluster barrier CTA r adds up -- in split order, so the
  // result is reproducible -- and stores a contiguous column range of the tile, reading the other CTAs' partials
  // over distributed shared memory.  No global partials, no tickets, no fences.
  const bool worker = warp < U_PRODUCERS / 32;
  const int row = 32 * (warp & 3) + lane;
  const int col0 = (warp >> 2) * 64;
  const bool vec_c = (d.ldc & 3) == 0 && aligned16_dev(d.c);

  auto store4 = [&](float4 v, int m, int n) {  // epilogue + store of columns n..n+3 of row m
    float* crow = d.c + (size_t)m * d.ldc;
    if (vec_c && n + 3 < d.n) {
      float4 o, old = make_float4(0.f, 0.f, 0.f, 0.f);
      if (d.beta) old = *reinterpret_cast<const float4*>(crow + n);
      o.x = epilogue(d, v.x, m, n + 0, &old.x);
      o.y = epilogue(d, v.y, m, n + 1, &old.y);
      o.z = epilogue(d, v.z, m, n + 2, &old.z);
      o.w = epilogue(d, v.w, m, n + 3, &old.w);
      *reinterpret_cast<float4*>(crow + n) = o;
    } else {
      const float e4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (n + e < d.n) crow[n + e] = epilogue(d, e4[e], m, n + e, crow + n + e);
    }
  };

  if (pb.splits == 1) {
    if (worker && m0 + row < d.m) {
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const int n = n0 + col0 + 4 * c;
        if (n < d.n) store4(make_float4(sum[4 * c], sum[4 * c + 1], sum[4 * c + 2], sum[4 * c + 3]), m0 + row, n);
      }
    }
  } else {
    cg::cluster_group cluster = cg::this_cluster();
    // partial tile as [32 column chunks][129 rows] float4 (pitch 129: chunk-fastest readers stay conflict-free);
    // the operand stages are free: every MMA of this CTA has completed (all groups drained)
    float4* red = reinterpret_cast<float4*>(u_smem);
    constexpr int RED_PITCH = 129;
    __syncthreads();
    if (worker) {
#pragma unroll
      for (int c = 0; c < 16; ++c)
        red[(col0 / 4 + c) * RED_PITCH + row] = make_float4(sum[4 * c], sum[4 * c + 1], sum[4 * c + 2], sum[4 * c + 3]);
    }
    U_STAMP(42);
    cluster.sync();
    U_STAMP(43);
    // 128 rows x (32 / splits) column chunks per CTA = 16 / splits items per thread; all 16 remote loads of a
    // thread are issued before the first add
    if (worker) {
      auto reduce = [&](auto splits_tag) {
        constexpr int SP = decltype(splits_tag)::value, CPC = 32 / SP, ITEMS = 128 * CPC / U_PRODUCERS;
        float4 v[ITEMS][SP];
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
          const int item = tid + it * U_PRODUCERS, chunk = split * CPC + item % CPC, r = item / CPC;
#pragma unroll
          for (int sp = 0; sp < SP; ++sp) v[it][sp] = cluster.map_shared_rank(red, sp)[chunk * RED_PITCH + r];
        }
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
          const int item = tid + it * U_PRODUCERS, chunk = split * CPC + item % CPC, r = item / CPC;
          float4 acc = v[it][0];
#pragma unroll
          for (int sp = 1; sp < SP; ++sp) { acc.x += v[it][sp].x; acc.y += v[it][sp].y; acc.z += v[it][sp].z; acc.w += v[it][sp].w; }
          const int m = m0 + r, n = n0 + 4 * chunk;
          if (m < d.m && n < d.n) store4(acc, m, n);
        }
      };
      if (pb.splits == 2) reduce(std::integral_constant<int, 2>{});
      else if (pb.splits == 4) reduce(std::integral_constant<int, 4>{});
      else reduce(std::integral_constant<int, 8>{});
    }
    cluster.sync();  // nobody leaves while a neighbour may still be reading its partial tile
  }
  U_STAMP(44);
  // every warp is done with its tcgen05.ld before the columns go back to the allocator
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(U_TMEM_COLS) : "memory");
  }
  U_STAMP(45);
#undef U_STAMP
}


// -----------------------------------------------------------------------------------------------
// Large problems (batch >= ~512): 128x128 tile, 8x8 outputs per thread, BK = 8.  Global loads are
// staged through registers (next tile prefetched while the current one is multiplied) and written to
// shared memory TRANSPOSED to [k][row], so that every fragment read is one 128-bit load per 4 rows and
// one k: 4 LDS.128 per 64 FMA, ~110 registers, two CTAs per SM.  Same split-K / ticket epilogue.
// -----------------------------------------------------------------------------------------------
constexpr int GB = 128, GK = 8, GPITCH = GB + 4;

template <bool KC>
__device__ __forceinline__ float4 big_load(const float* __restrict__ src, int ld, int t0, int tmax, int k0, int kmax,
                                           int tid) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (KC) {  // src[t * ld + k]: thread -> (row tid/2, k-quad tid%2)
    const int t = t0 + (tid >> 1), k = k0 + (tid & 1) * 4;
    if (t < tmax && k < kmax) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)t * ld + k));  // K % 4 == 0
  } else {   // src[k * ld + t]: thread -> (k tid/32, row-quad tid%32)
    const int k = k0 + (tid >> 5), t = t0 + (tid & 31) * 4;
    if (k < kmax && t < tmax) {
      const float* g = src + (size_t)k * ld + t;
      if (t + 3 < tmax) v = __ldg(reinterpret_cast<const float4*>(g));
      else { v.x = g[0]; if (t + 1 < tmax) v.y = g[1]; if (t + 2 < tmax) v.z = g[2]; }
    }
  }
  return v;
}
template <bool KC>
__device__ __forceinline__ void big_store(float (*s)[GPITCH], const float4& v, int tid) {
  if (KC) {
    const int row = tid >> 1, kq = (tid & 1) * 4;
    s[kq + 0][row] = v.x; s[kq + 1][row] = v.y; s[kq + 2][row] = v.z; s[kq + 3][row] = v.w;
  } else {
    *reinterpret_cast<float4*>(&s[tid >> 5][(tid & 31) * 4]) = v;
  }
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(THREADS, 2) gemm_big_kernel(const PipeBatch pb) {
  __shared__ __align__(16) float As[2][GK][GPITCH];
  __shared__ __align__(16) float Bs[2][GK][GPITCH];
  __shared__ unsigned int s_ticket;
  const int prob = blockIdx.z / pb.splits, split = blockIdx.z - prob * pb.splits;
  const GemmDesc d = prob == 0 ? pb.d[0] : (prob == 1 ? pb.d[1] : pb.d[2]);
  const int m0 = blockIdx.y * GB, n0 = blockIdx.x * GB;
  if (m0 >= d.m || n0 >= d.n) return;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int k_begin = split * pb.k_per_split;
  const int k_end = min(d.k, k_begin + pb.k_per_split);
  const int nk = k_end > k_begin ? (k_end - k_begin + GK - 1) / GK : 0;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 pa = big_load<A_KC>(d.a, d.lda, m0, d.m, k_begin, k_end, tid);
  float4 pbv = big_load<B_KC>(d.b, d.ldb, n0, d.n, k_begin, k_end, tid);
  if (nk > 0) {
    big_store<A_KC>(As[0], pa, tid);
    big_store<B_KC>(Bs[0], pbv, tid);
  }
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) {
      pa = big_load<A_KC>(d.a, d.lda, m0, d.m, k_begin + (kt + 1) * GK, k_end, tid);
      pbv = big_load<B_KC>(d.b, d.ldb, n0, d.n, k_begin + (kt + 1) * GK, k_end, tid);
    }
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      big_store<A_KC>(As[buf ^ 1], pa, tid);
      big_store<B_KC>(Bs[buf ^ 1], pbv, tid);
    }
    __syncthreads();
  }

  auto row_of = [&](int i) { return m0 + (i >> 2) * 64 + ty * 4 + (i & 3); };
  auto col_of = [&](int j) { return n0 + (j >> 2) * 64 + tx * 4 + (j & 3); };
  if (pb.splits == 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = row_of(i);
      if (m >= d.m) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = col_of(j);
        if (n >= d.n) continue;
        float* cp = d.c + (size_t)m * d.ldc + n;
        *cp = epilogue(d, acc[i][j], m, n, cp);
      }
    }
    return;
  }
  float* part = pb.part[prob];
  const size_t plane = (size_t)d.m * d.n;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = row_of(i);
    if (m >= d.m) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = col_of(j);
      if (n < d.n) __stcg(part + (size_t)split * plane + (size_t)m * d.n + n, acc[i][j]);
    }
  }
  __threadfence();
  __syncthreads();
  unsigned int* ticket = pb.tickets + (size_t)prob * gridDim.x * gridDim.y + blockIdx.y * gridDim.x + blockIdx.x;
  if (tid == 0) s_ticket = atomicAdd(ticket, 1u);
  __syncthreads();
  if (s_ticket != (unsigned)pb.splits - 1) return;
  __threadfence();
  // fold in split order; acc[][] is reused as the running sum, 8 loads in flight per row
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int sp = 0; sp < pb.splits; ++sp) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = row_of(i);
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = col_of(j);
        v[j] = (m < d.m && n < d.n) ? __ldcg(part + (size_t)sp * plane + (size_t)m * d.n + n) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] += v[j];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = row_of(i);
    if (m >= d.m) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = col_of(j);
      if (n >= d.n) continue;
      float* cp = d.c + (size_t)m * d.ldc + n;
      *cp = epilogue(d, acc[i][j], m, n, cp);
    }
  }
  if (tid == 0) *ticket = 0u;
}

bool pipe_ok(const GemmDesc& d) {
  // cp.async moves 16-byte chunks: bases 16-byte aligned, leading dimensions and (for k-contiguous
  // operands) K multiples of 4
  if (!aligned16(d.a) || !aligned16(d.b) || d.lda % 4 || d.ldb % 4) return false;
  if ((d.a_kc || d.b_kc) && d.k % 4) return false;
  if (d.k_split) {  // a k-tile (16 or 32 wide) must never straddle the two segments
    if (d.k_split % 32 || d.k_split >= d.k || !d.a2 || !d.b2) return false;
    if (!aligned16(d.a2) || !aligned16(d.b2) || d.lda2 % 4 || d.ldb2 % 4) return false;
  }
  return true;
}

}  // namespace

bool gemm_ksplit_ok(const GemmDesc& d) { return pipe_ok(d); }

// Zero the ticket block once per C-ABI call; every split-K launch leaves it zeroed again.
int prepare_gemm_workspace(void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!ws || ws_bytes < 65536) return GML_OK;
  GML_CUDA_TRY(cudaMemsetAsync(ws, 0, 65536, st));
  return GML_OK;
}

size_t gemm_workspace_bytes() {
  // Split-K is only used while tiles < 2 * SMs, with splits <= ceil(2 * SMs / tiles): the partial planes
  // never exceed (2 * SMs + tiles) tiles of 64 x 64 floats, i.e. < 4 * SMs tiles; plus the ticket block.
  return (size_t)4 * kNumSMs * 128 * 64 * sizeof(float) + 65536 + 1024;
}

// returns GML_E_UNSUPPORTED when the problems do not meet the pipeline's alignment rules
int launch_gemm_pipelined(const GemmDesc* descs, int count, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (count < 1 || count > 3) return GML_E_BADARG;
  PipeBatch pb;
  int max_m = 0, max_n = 0, min_k = 1 << 30;
  for (int i = 0; i < count; ++i) {
    if (descs[i].m <= 0 || descs[i].n <= 0 || descs[i].k <= 0) return GML_E_BADARG;
    if (!pipe_ok(descs[i])) return GML_E_UNSUPPORTED;
    if (descs[i].a_kc != descs[0].a_kc || descs[i].b_kc != descs[0].b_kc) return GML_E_UNSUPPORTED;
    pb.d[i] = descs[i];
    max_m = descs[i].m > max_m ? descs[i].m : max_m;
    max_n = descs[i].n > max_n ? descs[i].n : max_n;
    min_k = descs[i].k < min_k ? descs[i].k : min_k;
  }
  for (int i = count; i < 3; ++i) pb.d[i] = descs[0];
  for (int i = 1; i < count; ++i)
    if (descs[i].k != descs[0].k) return GML_E_UNSUPPORTED;
  // 128x128 tiles (gemm_big_kernel, 8x8 outputs per thread) measured SLOWER than the 64x64 cp.async kernel at
  // 3 CTAs per SM on every FC shape of the three blocks (profiles/r1_experiments.md), so they are opt-in only.
  bool any_ksplit = false;
  for (int i = 0; i < count; ++i) any_ksplit |= descs[i].k_split != 0;
  const bool big = g_gemm_big_tiles && !any_ksplit && (long)ceil_div(max_m, 128) * ceil_div(max_n, 128) * count >= 24 && min_k >= 64;
  // tcgen05 kernel: worth its fixed costs (TMEM allocation, 192 KB of shared memory, one CTA per SM) only for the
  // big FC problems (batch >= ~512 on the 256- and 512-channel blocks)
  // tensor cores only pay off on the large problems (measured cross-over, scripts/gemm_accuracy.py); everything
  // smaller stays on the CUDA-core kernel
  const bool large = (double)max_m * max_n * min_k * count >= 1e8;
  const bool umma = g_gemm_umma && !big && large && min_k >= 256 && max_m >= 128 && max_n >= 128;
  const bool mma_sync = g_gemm_tf32x3 && large;
  const int BM = (big || umma) ? 128 : 64, BN = BM;
  const int bk = umma ? UK : BK;
  const int tiles_m = ceil_div(max_m, BM), tiles_n = ceil_div(max_n, BN);
  const long tiles = (long)tiles_m * tiles_n * count;  // grid extent (CTAs outside their problem exit at once)
  long busy = 0;                                        // tiles that actually compute
  for (int i = 0; i < count; ++i) busy += (long)ceil_div(descs[i].m, BM) * ceil_div(descs[i].n, BN);
  const int nk = ceil_div(min_k, bk);
  // tcgen05 kernel: one CTA per SM (shared memory), so the busy CTAs must fit ONE wave: splits = floor(SMs / busy)
  int splits = umma ? (int)(kNumSMs / busy) : (int)((2 * kNumSMs + busy - 1) / busy);
  const int min_tiles_per_split = umma ? 2 : 4;
  if (splits > nk / min_tiles_per_split) splits = nk / min_tiles_per_split;
  if (splits > 16) splits = 16;
  if (splits < 1) splits = 1;
  if (umma) {  // the splits of a tile form one thread-block cluster: 1, 2, 4 or 8
    int p2 = 1;
    while (p2 * 2 <= splits && p2 < 8) p2 *= 2;
    splits = p2;
  }
  const size_t ticket_bytes = 65536;  // fixed-size ticket block at the head of the workspace
  if ((size_t)tiles * sizeof(unsigned int) > ticket_bytes) return GML_E_UNSUPPORTED;
  while (splits > 1 && !umma) {
    size_t need = ticket_bytes;
    for (int i = 0; i < count; ++i) need += round_up((size_t)splits * descs[i].m * descs[i].n * sizeof(float), 256);
    if (ws && need <= ws_bytes) break;
    --splits;
  }
  pb.splits = splits;
  pb.trace = g_gemm_trace;
  pb.k_per_split = ceil_div(nk, splits) * bk;
  pb.tickets = nullptr;
  pb.part[0] = pb.part[1] = pb.part[2] = nullptr;
  if (splits > 1 && !umma) {
    char* p = static_cast<char*>(ws);
    pb.tickets = reinterpret_cast<unsigned int*>(p);
    p += ticket_bytes;
    for (int i = 0; i < count; ++i) {
      pb.part[i] = reinterpret_cast<float*>(p);
      p += round_up((size_t)splits * descs[i].m * descs[i].n * sizeof(float), 256);
    }
    // tickets must be zero on entry (prepare_gemm_workspace) and are restored to zero by the kernel
  }
  dim3 grid(tiles_n, tiles_m, count * splits);
  {
    LaunchScope ls(kTagGemm, st);
    const bool akc = descs[0].a_kc != 0, bkc = descs[0].b_kc != 0;
#define GML_GEMM(AK, BKC)                                                                                   \
  do {                                                                                                      \
    if (umma) {                                                                                             \
      const size_t sm = (size_t)USTAGES * U_STAGE_BYTES + 1024;                                              \
      /* per device and cheap: set on every launch (a process may drive several GPUs) */                    \
      GML_CUDA_TRY(cudaFuncSetAttribute(gemm_umma_kernel<AK, BKC>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                        (int)sm));                                                          \
      cudaLaunchConfig_t cfg = {};                                                                          \
      cfg.gridDim = grid; cfg.blockDim = dim3(UTHREADS); cfg.dynamicSmemBytes = sm; cfg.stream = st;         \
      cudaLaunchAttribute attr[1];                                                                          \
      attr[0].id = cudaLaunchAttributeClusterDimension;                                                     \
      attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = splits;         \
      cfg.attrs = attr; cfg.numAttrs = 1;                                                                   \
      GML_CUDA_TRY(cudaLaunchKernelEx(&cfg, gemm_umma_kernel<AK, BKC>, pb));                                 \
    } else if (big) {                                                                                       \
      gemm_big_kernel<AK, BKC><<<grid, THREADS, 0, st>>>(pb);                                                \
    } else if (mma_sync) {                                                                                  \
      const size_t sm = 2 * STAGES * MMA_TILE_FLOATS * sizeof(float);                                        \
      gemm_mma_kernel<AK, BKC><<<grid, THREADS, sm, st>>>(pb);                                               \
    } else {                                                                                                \
      const size_t sm = 2 * STAGES * TileGeom<64>::kFloats * sizeof(float);                                  \
      gemm_pipe_kernel<AK, BKC, 4><<<grid, THREADS, sm, st>>>(pb);                                           \
    }                                                                                                       \
  } while (0)
    if (akc && bkc) GML_GEMM(true, true);
    else if (akc && !bkc) GML_GEMM(true, false);
    else if (!akc && bkc) GML_GEMM(false, true);
    else GML_GEMM(false, false);
#undef GML_GEMM
  }
  GML_LAUNCH_CHECK();
  return GML_OK;
}

}  // namespace gml
