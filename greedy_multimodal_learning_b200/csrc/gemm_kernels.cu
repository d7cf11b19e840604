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
#include "common.cuh"
#include "kernels.h"

namespace gml {

int g_gemm_big_tiles = 0;  // tunable "gemm_big_tiles"

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
  GemmDesc d[2];
  float* part[2];        // split-K partials per problem: [splits][m][n]
  unsigned int* tickets; // [count][tiles_m * tiles_n]
  int splits;
  int k_per_split;       // multiple of BK
};

__device__ __forceinline__ void cp_async16(float* dst_smem, const float* src, int src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// one operand tile for k in [k0, k0 + BK): KC = k-contiguous source (src[t * ld + k]), else src[k * ld + t]
template <bool KC, int BT>
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
      cp_async16(s + kk * TileGeom<BT>::kPitchMN + tq, g, bytes);
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
  const GemmDesc d = prob ? pb.d[1] : pb.d[0];
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
      load_tile<A_KC, BM>(As[s], d.a, d.lda, m0, d.m, k_begin + s * BK, k_end, tid);
      load_tile<B_KC, BN>(Bs[s], d.b, d.ldb, n0, d.n, k_begin + s * BK, k_end, tid);
    }
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    const int nxt = kt + STAGES - 1;
    if (nxt < nk) {
      load_tile<A_KC, BM>(As[nxt % STAGES], d.a, d.lda, m0, d.m, k_begin + nxt * BK, k_end, tid);
      load_tile<B_KC, BN>(Bs[nxt % STAGES], d.b, d.ldb, n0, d.n, k_begin + nxt * BK, k_end, tid);
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
  const GemmDesc d = prob ? pb.d[1] : pb.d[0];
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
  return true;
}

}  // namespace

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
  if (count < 1 || count > 2) return GML_E_BADARG;
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
  if (count == 1) pb.d[1] = descs[0];
  if (count == 2 && descs[0].k != descs[1].k) return GML_E_UNSUPPORTED;
  // 128x128 tiles (gemm_big_kernel, 8x8 outputs per thread) measured SLOWER than the 64x64 cp.async kernel at
  // 3 CTAs per SM on every FC shape of the three blocks (profiles/r1_experiments.md), so they are opt-in only.
  const bool big = g_gemm_big_tiles && (long)ceil_div(max_m, 128) * ceil_div(max_n, 128) * count >= 24 && min_k >= 64;
  const int BM = big ? 128 : 64, BN = BM;
  const int tiles_m = ceil_div(max_m, BM), tiles_n = ceil_div(max_n, BN);
  const long tiles = (long)tiles_m * tiles_n * count;
  const int nk = ceil_div(min_k, BK);
  int splits = (int)((2 * kNumSMs + tiles - 1) / tiles);
  if (splits > nk / 4) splits = nk / 4;  // at least 4 k-tiles per split
  if (splits > 16) splits = 16;
  if (splits < 1) splits = 1;
  const size_t ticket_bytes = 65536;  // fixed-size ticket block at the head of the workspace
  if ((size_t)tiles * sizeof(unsigned int) > ticket_bytes) return GML_E_UNSUPPORTED;
  while (splits > 1) {
    size_t need = ticket_bytes;
    for (int i = 0; i < count; ++i) need += round_up((size_t)splits * descs[i].m * descs[i].n * sizeof(float), 256);
    if (ws && need <= ws_bytes) break;
    --splits;
  }
  pb.splits = splits;
  pb.k_per_split = ceil_div(nk, splits) * BK;
  pb.tickets = nullptr;
  pb.part[0] = pb.part[1] = nullptr;
  if (splits > 1) {
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
    if (big) {                                                                                              \
      gemm_big_kernel<AK, BKC><<<grid, THREADS, 0, st>>>(pb);                                                \
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
