// Single-pass MMTM kernels: thread-block clusters keep a group of samples resident in shared
// memory, so the feature maps cross HBM exactly once per direction (forward 4u, backward 6u;
// u = N*C*HW*4 B per modality) instead of the 6u / 8u of the streaming two-pass path.
//
// Layout of the work
//   * a cluster of CS CTAs owns a group of G samples; CTA r holds channels [r*C/CS, (r+1)*C/CS)
//     of BOTH modalities of those samples.  Two geometries:
//       CS = 4, one  CTA per SM, ~200 KB of planes per CTA   (128x28^2: G=1, 256x14^2: G=2)
//       CS = 8, two CTAs per SM, ~100 KB of planes per CTA   (same G) -- the second resident CTA
//               streams while the first one sits in its FC / cluster-barrier phase
//   * the CTA's planes arrive as NCHUNK 1-D TMA bulk copies (cp.async.bulk, contiguous in NCHW),
//     each signalling its own mbarrier: the reduction pass consumes chunks as they land, and
//     while the output pass drains chunk j to HBM the NEXT group's chunk j is already being
//     fetched into the freed space (loads and stores overlap inside one CTA).
//   * the squeeze vector (2C floats per sample) and the hidden state (D floats) are exchanged
//     through distributed shared memory (every CTA writes its part into all CTAs of the cluster),
//     two cluster barriers per group; the FCs are per-sample GEMVs on CUDA cores, weights streamed
//     from L2 with several rows in flight per warp (latency, not bandwidth, is what matters there).
//   * persistent: clusters loop over groups with stride = number of resident clusters.
// Forward: resident = inputs.  Backward: resident = grad_out (needed twice: dot, then apply);
// the saved inputs stream through registers once for the dot.  Weight gradients are NOT formed
// here: dE / dH rows go to global memory and the batched GEMMs in fc_kernels.cu reduce them over
// the batch in a fixed order.
//
// Shapes outside the supported set fall back to the streaming kernels (capi.cu).
#include <cooperative_groups.h>

#include <map>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <utility>

#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace gml {

int g_fused_cluster = 0;   // tunables (gml_set_tunable): 0 = automatic
int g_fused_threads = 0;
int g_fused_kind = 0;      // 0 auto, 1 shared-memory resident, 2 L2 resident
int g_fused_occ = 4;       // L2-resident kernels: CTAs per SM the register budget is compiled for (4: 64 regs, 5: 48)
int g_fused_group_kb = 128;  // L2-resident kernels: take 2 samples per cluster while 2 slices <= this many KB per CTA
int g_fused_stash_kb = -1; // L2-resident kernels: shared memory per CTA used to stash planes between the passes; -1 = automatic:
                           // 24 KB (measured best; 46+ costs occupancy), none for a backward that has its weights in shared memory
int g_fused_hw_special = 1;  // L2-resident kernels: 1 = plane-size-specialised instantiation for 28 x 28 planes (tunable "fused_hw_special")
int g_fused_wsmem = -1;    // L2-resident kernels: FC weight slices prefetched into shared memory (cp.async, hidden behind pass 1);
                           // -1 auto (whatever fits at 4 CTAs per SM), else bit 0 = first FC, bit 1 = second FC
int g_fused_prefetch = 0;  // L2-resident kernels: bulk L2 prefetch look-ahead distance in groups (0 = off)
long long* g_fused_occ_trace = nullptr;  // debug: per-CTA {smid, start ns, end ns, 0} of the L2-resident kernels (4 slots per CTA)
long long* g_fused_trace = nullptr;  // debug: per-phase clock64() stamps of the first CTAs (device buffer)

namespace {

constexpr int kMaxCluster = 8;
constexpr int kMaxChunks = 16;
constexpr int kRowsPerBatch = 4;  // GEMV rows a warp keeps in flight at once

struct FusedCfg {
  int n, c, hw, d;
  int cs;        // CTAs per cluster
  int threads;   // threads per CTA
  int g;         // samples per group
  int cq, dq;    // C/cs, D/cs
  int pl;        // planes per CTA = g * 2 * cq
  int pc;        // planes per chunk
  int nchunk;    // pl / pc
  int n_groups;
  size_t data_bytes;
  long long* trace;  // nullptr unless phase tracing is on: [cta < 8][iter < 16][16 stamps]
  int prefetch;      // L2-resident kernels: issue bulk L2 prefetches of the CTA's planes up front
  int trace_first;   // first CTA of the traced window (L2-resident kernels)
  int keep_planes;   // L2-resident kernels: the first keep_planes planes of a CTA are stashed in shared memory
  long long* occ_trace;  // nullptr unless the residency trace is on
  int wsm;           // L2-resident kernels: bit 0 / bit 1 = weight slice of the first / second FC lives in shared memory
};

#define GML_STAMP(k)                                                                       \
  do {                                                                                     \
    if (f.trace && threadIdx.x == 0 && blockIdx.x < 8 && iter < 16)                        \
      f.trace[((size_t)blockIdx.x * 16 + iter) * 16 + (k)] = clock64();                     \
  } while (0)

// ---- PTX wrappers --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  // bounded spin: a lost transaction must surface as an error, never as a hung GPU
  for (uint32_t spins = 0; !ok; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!ok && spins > (1u << 26)) __trap();
  }
}
// 1-D TMA: global -> this CTA's shared memory, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// fire-and-forget HBM -> L2 bulk prefetch (no registers, no shared memory, no completion tracking)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(src), "r"(bytes), "l"(policy)
               : "memory");
}

struct Smem {
  float* data;
  uint64_t* bars;
  float* psum;    // [pl] plane sums / dots
  float* scale;   // [pl] per-plane multiplier of the output pass
  float* addv;    // [pl] per-plane additive term (backward)
  float* bias_h;  // [dq]      my slice of b_sq
  float* bias_g;  // [2][cq]   my slices of b_v, b_s
  float* vec_a;   // [g][2C]  z (fwd) / dE (bwd), full vector, filled by all CTAs of the cluster
  float* vec_b;   // [g][D]   h (fwd) / dH (bwd)
  float* part;    // cross-warp partial sums of the transposed GEMVs (backward)
  int* done;      // [nchunk] planes of chunk j already drained in the output pass
};

__host__ __device__ inline size_t extras_bytes(const FusedCfg& f, bool bwd) {
  size_t b = kMaxChunks * (sizeof(uint64_t) + sizeof(int)) + 3 * (size_t)f.pl * 4 + ((size_t)f.dq + 2 * f.cq) * 4 +
             (size_t)f.g * 2 * f.c * 4 + (size_t)f.g * f.d * 4;
  if (bwd) b += (size_t)f.threads * f.g * 4;  // part[slices][g][cols], slices * cols == threads
  return b + 16;
}

__device__ __forceinline__ Smem carve(unsigned char* base, const FusedCfg& f) {
  Smem s;
  s.data = reinterpret_cast<float*>(base);
  unsigned char* p = base + f.data_bytes;
  s.bars = reinterpret_cast<uint64_t*>(p); p += kMaxChunks * sizeof(uint64_t);
  s.done = reinterpret_cast<int*>(p); p += kMaxChunks * sizeof(int);
  s.vec_a = reinterpret_cast<float*>(p); p += (size_t)f.g * 2 * f.c * sizeof(float);   // 16-byte aligned
  s.vec_b = reinterpret_cast<float*>(p); p += (size_t)f.g * f.d * sizeof(float);       // 16-byte aligned
  s.psum = reinterpret_cast<float*>(p); p += f.pl * sizeof(float);
  s.scale = reinterpret_cast<float*>(p); p += f.pl * sizeof(float);
  s.addv = reinterpret_cast<float*>(p); p += f.pl * sizeof(float);
  s.bias_h = reinterpret_cast<float*>(p); p += f.dq * sizeof(float);
  s.bias_g = reinterpret_cast<float*>(p); p += 2 * f.cq * sizeof(float);
  s.part = reinterpret_cast<float*>(p);
  return s;
}

// plane p of this CTA -> (sample in group, modality, local channel)
__device__ __forceinline__ void plane_coords(const FusedCfg& f, int p, int& g, int& mod, int& cl) {
  const int sl = p / f.cq;
  cl = p - sl * f.cq;
  g = sl >> 1;
  mod = sl & 1;
}

// issue the bulk copies of one group (chunks [j0, j1)), called by one thread
__device__ __forceinline__ void issue_chunks(const FusedCfg& f, const Smem& s, const float* xa, const float* xb, int rank,
                                             int n0, int gcount, int j0, int j1, uint64_t policy) {
  const uint32_t chunk_bytes = (uint32_t)f.pc * f.hw * sizeof(float);
  for (int j = j0; j < j1; ++j) {
    const int p0 = j * f.pc;
    int g, mod, cl;
    plane_coords(f, p0, g, mod, cl);
    if (g >= gcount) break;
    const float* src = (mod ? xb : xa) + ((size_t)(n0 + g) * f.c + (size_t)rank * f.cq + cl) * f.hw;
    mbar_expect_tx(&s.bars[j], chunk_bytes);
    bulk_g2s(s.data + (size_t)p0 * f.hw, src, chunk_bytes, &s.bars[j], policy);
  }
}


// Output-pass bookkeeping: the group that drains the LAST plane of chunk j re-arms the chunk with
// the next group's data (no block-wide barrier in the streaming loops).
__device__ __forceinline__ void plane_drained(const FusedCfg& f, const Smem& s, int p, const float* xa, const float* xb,
                                              int rank, int next_n0, int next_gcount, uint64_t policy) {
  const int j = p / f.pc;
  // No fence here on purpose: every shared-memory read of this plane has already returned its value
  // (the stores that consume them have been issued), and a MEMBAR would also wait for those global
  // stores to be acknowledged -- a full round trip per plane.
  const int prev = atomicAdd(&s.done[j], 1);
  if (prev == f.pc - 1) {
    s.done[j] = 0;
    if (next_gcount > 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue_chunks(f, s, xa, xb, rank, next_n0, next_gcount, j, j + 1, policy);
    }
  }
}

// y[row] = <W[row0 + row, 0:K], x[g, 0:K]>, K % 4 == 0.  Each warp owns kRowsPerBatch rows at a time
// and issues all their weight loads before the first FMA (the weights come from L2: the cost is
// latency, so several rows must be in flight); lane r then finishes row r (see the end of the loop).
template <int T, int GMAX, int R, typename RowPtr, typename Epi>
__device__ __forceinline__ void gemv_rows_r(RowPtr rowptr, int nrows, int k, const float* x, int ldx, int gcount,
                                            Epi epi) {
  constexpr int kWarps = T / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k4 = k >> 2;
  for (int base = warp * R; base < nrows; base += kWarps * R) {
    float acc[R][GMAX];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int g = 0; g < GMAX; ++g) acc[r][g] = 0.f;
#pragma unroll 1
    for (int i = lane; i < k4; i += 32) {
      float4 wv[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int row = min(base + r, nrows - 1);  // clamp: duplicates are discarded in the epilogue
        wv[r] = *(reinterpret_cast<const float4*>(rowptr(row)) + i);  // generic load: global or shared memory
      }
#pragma unroll
      for (int g = 0; g < GMAX; ++g) {
        if (g < gcount) {
          const float4 xv = *reinterpret_cast<const float4*>(x + (size_t)g * ldx + 4 * i);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            acc[r][g] = fmaf(wv[r].x, xv.x, acc[r][g]); acc[r][g] = fmaf(wv[r].y, xv.y, acc[r][g]);
            acc[r][g] = fmaf(wv[r].z, xv.z, acc[r][g]); acc[r][g] = fmaf(wv[r].w, xv.w, acc[r][g]);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int g = 0; g < GMAX; ++g) acc[r][g] = warp_sum(acc[r][g]);
    // every lane holds every sum (xor butterfly): lane r finishes row base + r, all R of them at once -- one copy of the
    // epilogue, executed once per batch (R sequential single-lane epilogues, each a sigmoid or eight remote stores,
    // were the longest part of the FC phases)
    float mine[GMAX];
#pragma unroll
    for (int g = 0; g < GMAX; ++g) {
      mine[g] = acc[0][g];
#pragma unroll
      for (int r = 1; r < R; ++r) mine[g] = lane == r ? acc[r][g] : mine[g];
    }
    if (lane < R && base + lane < nrows) epi(base + lane, mine);
  }
}

// The rows may live in global or in shared memory (generic loads): one copy of the code serves both, which matters here --
// with a separate instantiation per address space and per row batch the forward kernel was 14k instructions and its
// once-per-CTA FC phases ran out of the instruction cache.
template <int T, int GMAX, typename RowPtr, typename Epi>
__device__ __forceinline__ void gemv_rows(RowPtr rowptr, int nrows, int k, const float* x, int ldx, int gcount,
                                          Epi epi) {
  // few rows: spread them over more warps (one L2 round trip either way)
  if (nrows <= 2 * (T / 32)) gemv_rows_r<T, GMAX, 2>(rowptr, nrows, k, x, ldx, gcount, epi);
  else gemv_rows_r<T, GMAX, kRowsPerBatch>(rowptr, nrows, k, x, ldx, gcount, epi);
}

// y[col] = sum_k x[g, k] * W[k, col0 + col] (transposed GEMV): thread = (k-slice, col), coalesced in
// col.  Partials are ADDED into s_part[slice][g][col]; the caller reduces the slices in a fixed order.
template <int T, int GMAX, bool SM = false>
__device__ __forceinline__ void gemv_cols_partial(const float* w, int ldw, int col0, int ncols, int k0,
                                                  int k1, const float* x, int ldx, int gcount, float* s_part) {
  const int slices = T / ncols;
  const int col = threadIdx.x % ncols, sl = threadIdx.x / ncols;
  const int span = (k1 - k0 + slices - 1) / slices;
  const int ka = k0 + sl * span, kb = min(k1, ka + span);
  float acc[GMAX];
#pragma unroll
  for (int g = 0; g < GMAX; ++g) acc[g] = 0.f;
  const float* wp = w + col0 + col;
#pragma unroll 16
  for (int kk = ka; kk < kb; ++kk) {
    // (the strided global reads want the non-coherent path: generic loads made this GEMV ~2x slower)
    const float wv = SM ? wp[(size_t)kk * ldw] : __ldg(wp + (size_t)kk * ldw);
#pragma unroll
    for (int g = 0; g < GMAX; ++g)
      if (g < gcount) acc[g] = fmaf(x[(size_t)g * ldx + kk], wv, acc[g]);
  }
#pragma unroll
  for (int g = 0; g < GMAX; ++g) s_part[((size_t)sl * GMAX + g) * ncols + col] += acc[g];
}

template <int T>
__device__ __forceinline__ void common_prologue(const FusedCfg& f, const Smem& s, const float* b_sq, const float* b_v,
                                                const float* b_s, int rank) {
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int j = 0; j < f.nchunk; ++j) {
      mbar_init(&s.bars[j], 1);
      s.done[j] = 0;
    }
    fence_mbar_init();
  }
  if (b_sq) {
    for (int i = tid; i < f.dq; i += T) s.bias_h[i] = __ldg(b_sq + rank * f.dq + i);
    for (int i = tid; i < 2 * f.cq; i += T)
      s.bias_g[i] = __ldg((i < f.cq ? b_v : b_s) + rank * f.cq + (i < f.cq ? i : i - f.cq));
  }
  __syncthreads();
}

// =============================================================================================
// forward
// =============================================================================================
template <int T, int L, int GMAX>
__global__ void __launch_bounds__(T, (T == 256 ? 2 : 1)) fused_fwd_kernel(const FusedFwdArgs a, const FusedCfg f) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cluster_id = blockIdx.x / f.cs, n_clusters = gridDim.x / f.cs;
  const Smem s = carve(smem_raw, f);
  const int tid = threadIdx.x;
  const uint64_t pol_stream = policy_evict_first();

  common_prologue<T>(f, s, a.b_sq, a.b_v, a.b_s, rank);
  if (tid == 0 && cluster_id < f.n_groups) {
    const int n0 = cluster_id * f.g;
    issue_chunks(f, s, a.a, a.b, rank, n0, min(f.g, f.n - n0), 0, f.nchunk, pol_stream);
  }
  cluster.sync();  // every CTA of the cluster is alive before anyone writes remote shared memory

  constexpr int kPlanesPerPass = T / L;
  const int lane = tid % L, grp_in_pass = tid / L;
  const int hw4 = f.hw >> 2;
  uint32_t parity = 0;
  int iter = 0;

  for (int grp = cluster_id; grp < f.n_groups; grp += n_clusters) {
    const int n0 = grp * f.g;
    const int gcount = min(f.g, f.n - n0);
    const int vplanes = gcount * 2 * f.cq;
    GML_STAMP(0);

    // ---- pass 1: plane sums straight out of shared memory as chunks land --------------------
    // group gi owns planes gi, gi + NG, ...: every warp has work in every chunk wave
    for (int p = grp_in_pass; p < vplanes; p += kPlanesPerPass) {
      mbar_wait(&s.bars[p / f.pc], parity);
      const float4* v = reinterpret_cast<const float4*>(s.data + (size_t)p * f.hw);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
      for (int i = lane; i < hw4; i += L) {
        const float4 x = v[i];
        a0 += x.x; a1 += x.y; a2 += x.z; a3 += x.w;
      }
      const float t = group_sum<L>((a0 + a1) + (a2 + a3));
      if (lane == 0) s.psum[p] = t;
    }
    GML_STAMP(1);
    __syncthreads();
    GML_STAMP(2);
    // ---- squeeze vector -> every CTA of the cluster (DSMEM) + global z ------------------------
    for (int p = tid; p < vplanes; p += T) {
      int g, mod, cl;
      plane_coords(f, p, g, mod, cl);
      const int k = mod * f.c + rank * f.cq + cl;
      const float mean = s.psum[p] / (float)f.hw;
      for (int dst = 0; dst < f.cs; ++dst) cluster.map_shared_rank(s.vec_a, dst)[g * 2 * f.c + k] = mean;
      a.z[(size_t)(n0 + g) * 2 * f.c + k] = mean;
    }
    cluster.sync();
    GML_STAMP(3);
    // ---- FC1: my slice of the hidden units, H = relu(Wsq z + bsq) ------------------------------
    gemv_rows<T, GMAX>([&](int r) { return a.w_sq + (size_t)(rank * f.dq + r) * 2 * f.c; }, f.dq, 2 * f.c, s.vec_a,
                       2 * f.c, gcount, [&](int r, const float* acc) {
                         const int dd = rank * f.dq + r;
                         const float bias = s.bias_h[r];
                         for (int g = 0; g < gcount; ++g) {
                           const float hval = fmaxf(acc[g] + bias, 0.f);
                           for (int dst = 0; dst < f.cs; ++dst)
                             cluster.map_shared_rank(s.vec_b, dst)[g * f.d + dd] = hval;
                           a.h[(size_t)(n0 + g) * f.d + dd] = hval;
                         }
                       });
    GML_STAMP(4);
    cluster.sync();
    GML_STAMP(5);
    // ---- FC2: gates of my channels, both modalities (rows [0,cq) = visual, [cq,2cq) = skeleton) --
    gemv_rows<T, GMAX>(
        [&](int r) {
          return r < f.cq ? a.w_v + (size_t)(rank * f.cq + r) * f.d : a.w_s + (size_t)(rank * f.cq + r - f.cq) * f.d;
        },
        2 * f.cq, f.d, s.vec_b, f.d, gcount, [&](int r, const float* acc) {
          const int mod = r >= f.cq, cl = r - mod * f.cq;
          const int ch = rank * f.cq + cl;
          const float bias = s.bias_g[r];
          float* gout = mod ? a.g_b : a.g_a;
          for (int g = 0; g < gcount; ++g) {
            const float gate = sigmoidf_ref(acc[g] + bias);
            s.scale[(g * 2 + mod) * f.cq + cl] = gate * a.gate_scale;
            gout[(size_t)(n0 + g) * f.c + ch] = gate;
          }
        });
    GML_STAMP(6);
    __syncthreads();
    GML_STAMP(7);
    // ---- pass 2: gate from shared memory, stream out; refill freed chunks with the next group --
    const int next = grp + n_clusters;
    const int next_n0 = next * f.g;
    const int next_gcount = next < f.n_groups ? min(f.g, f.n - next_n0) : 0;
    for (int p = grp_in_pass; p < vplanes; p += kPlanesPerPass) {
      int g, mod, cl;
      plane_coords(f, p, g, mod, cl);
      const float sc = s.scale[p];
      const float4* v = reinterpret_cast<const float4*>(s.data + (size_t)p * f.hw);
      float4* o = reinterpret_cast<float4*>((mod ? a.b_out : a.a_out) +
                                            ((size_t)(n0 + g) * f.c + (size_t)rank * f.cq + cl) * f.hw);
#pragma unroll 8
      for (int i = lane; i < hw4; i += L) {
        float4 x = v[i];
        x.x *= sc; x.y *= sc; x.z *= sc; x.w *= sc;
        stg_stream(o + i, x);
      }
      if (L < 32) __syncwarp();
      if (lane == 0) plane_drained(f, s, p, a.a, a.b, rank, next_n0, next_gcount, pol_stream);
    }
    GML_STAMP(8);
    __syncthreads();  // all refills of this iteration are issued; psum/scale may be reused
    GML_STAMP(9);
    parity ^= 1;
    ++iter;
  }
  cluster.sync();  // nobody exits while a sibling may still address its shared memory
}

// =============================================================================================
// backward
// =============================================================================================
template <int T, int L, int GMAX>
__global__ void __launch_bounds__(T, (T == 256 ? 2 : 1)) fused_bwd_kernel(const FusedBwdArgs a, const FusedCfg f) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cluster_id = blockIdx.x / f.cs, n_clusters = gridDim.x / f.cs;
  const Smem s = carve(smem_raw, f);
  const int tid = threadIdx.x;
  const uint64_t pol_stream = policy_evict_first();

  common_prologue<T>(f, s, nullptr, nullptr, nullptr, rank);
  if (tid == 0 && cluster_id < f.n_groups) {
    const int n0 = cluster_id * f.g;
    issue_chunks(f, s, a.go_a, a.go_b, rank, n0, min(f.g, f.n - n0), 0, f.nchunk, pol_stream);
  }
  cluster.sync();

  constexpr int kPlanesPerPass = T / L;
  const int lane = tid % L, grp_in_pass = tid / L;
  const int hw4 = f.hw >> 2;
  const int ncol_h = f.dq;        // outputs of the dH GEMV handled by this CTA
  const int ncol_z = 2 * f.cq;    // outputs of the dZ GEMV handled by this CTA
  uint32_t parity = 0;
  int iter = 0;

  for (int grp = cluster_id; grp < f.n_groups; grp += n_clusters) {
    const int n0 = grp * f.g;
    const int gcount = min(f.g, f.n - n0);
    const int vplanes = gcount * 2 * f.cq;
    GML_STAMP(0);

    // ---- early, latency-hiding loads of the per-plane gate and my slice of the ReLU mask -------
    float gate_pf = 0.f, h_pf = 0.f;
    if (tid < vplanes) {
      int g, mod, cl;
      plane_coords(f, tid, g, mod, cl);
      gate_pf = __ldg((mod ? a.g_b : a.g_a) + (size_t)(n0 + g) * f.c + rank * f.cq + cl);
    }
    if (tid < ncol_h * gcount) {
      const int g = tid / ncol_h, col = tid - g * ncol_h;
      h_pf = __ldg(a.h + (size_t)(n0 + g) * f.d + rank * f.dq + col);
    }
    // ---- pass 1: <grad_out (shared), input (global, streamed once)> per plane -----------------
    for (int p = grp_in_pass; p < vplanes; p += kPlanesPerPass) {
      int g, mod, cl;
      plane_coords(f, p, g, mod, cl);
      const float4* gv = reinterpret_cast<const float4*>(s.data + (size_t)p * f.hw);
      const float4* xv = reinterpret_cast<const float4*>((mod ? a.b : a.a) +
                                                         ((size_t)(n0 + g) * f.c + (size_t)rank * f.cq + cl) * f.hw);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      bool waited = false;
      for (int i0 = 0; i0 < hw4; i0 += 8 * L) {
        // up to 8 independent 128-bit global loads per lane in flight, issued BEFORE blocking on the
        // chunk barrier so their latency overlaps the wait
        float4 x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + lane + u * L;
          x[u] = i < hw4 ? ldg_hint(xv + i, pol_stream) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (!waited) {
          mbar_wait(&s.bars[p / f.pc], parity);
          waited = true;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + lane + u * L;
          if (i < hw4) {
            const float4 gg = gv[i];
            a0 = fmaf(gg.x, x[u].x, a0); a1 = fmaf(gg.y, x[u].y, a1);
            a2 = fmaf(gg.z, x[u].z, a2); a3 = fmaf(gg.w, x[u].w, a3);
          }
        }
      }
      const float t = group_sum<L>((a0 + a1) + (a2 + a3));
      if (lane == 0) s.psum[p] = t;
    }
    GML_STAMP(1);
    __syncthreads();
    GML_STAMP(2);
    // ---- dE of my channels -> all CTAs + global ------------------------------------------------
    if (tid < vplanes) {
      const int p = tid;
      int g, mod, cl;
      plane_coords(f, p, g, mod, cl);
      const int ch = rank * f.cq + cl;
      const float gate = gate_pf;
      const float de = s.psum[p] * a.gate_scale * gate * (1.f - gate);
      s.scale[p] = gate * a.gate_scale;
      for (int dst = 0; dst < f.cs; ++dst)
        cluster.map_shared_rank(s.vec_a, dst)[g * 2 * f.c + mod * f.c + ch] = de;
      (mod ? a.de_b : a.de_a)[(size_t)(n0 + g) * f.c + ch] = de;
    }
    for (int i = tid; i < T * GMAX; i += T) s.part[i] = 0.f;
    cluster.sync();
    GML_STAMP(3);
    // ---- dH for my slice of the hidden units: dE_a Wv + dE_b Ws, masked by H > 0 ----------------
    gemv_cols_partial<T, GMAX>(a.w_v, f.d, rank * f.dq, ncol_h, 0, f.c, s.vec_a, 2 * f.c, gcount, s.part);
    gemv_cols_partial<T, GMAX>(a.w_s, f.d, rank * f.dq, ncol_h, 0, f.c, s.vec_a + f.c, 2 * f.c, gcount, s.part);
    __syncthreads();
    {
      const int slices = T / ncol_h;
      if (tid < ncol_h * gcount) {
        const int g = tid / ncol_h, col = tid - g * ncol_h;
        float v = 0.f;
        for (int sl = 0; sl < slices; ++sl) v += s.part[((size_t)sl * GMAX + g) * ncol_h + col];
        const int dd = rank * f.dq + col;
        v = h_pf > 0.f ? v : 0.f;
        for (int dst = 0; dst < f.cs; ++dst) cluster.map_shared_rank(s.vec_b, dst)[g * f.d + dd] = v;
        a.dh[(size_t)(n0 + g) * f.d + dd] = v;
      }
    }
    GML_STAMP(4);
    cluster.sync();
    GML_STAMP(5);
    // ---- dZ of my channels: dH Wsq[:, my columns] -----------------------------------------------
    {
      // columns of this CTA: [rank*cq, +cq) of the visual half and [C + rank*cq, +cq) of the skeleton half
      const int slices = T / ncol_z;
      const int col = tid % ncol_z, sl = tid / ncol_z;
      const int gcol = (col < f.cq) ? rank * f.cq + col : f.c + rank * f.cq + (col - f.cq);
      const int span = (f.d + slices - 1) / slices;
      const int ka = sl * span, kb = min(f.d, ka + span);
      float acc[GMAX];
#pragma unroll
      for (int g = 0; g < GMAX; ++g) acc[g] = 0.f;
#pragma unroll 16
      for (int kk = ka; kk < kb; ++kk) {
        const float wv = __ldg(a.w_sq + (size_t)kk * 2 * f.c + gcol);
#pragma unroll
        for (int g = 0; g < GMAX; ++g)
          if (g < gcount) acc[g] = fmaf(s.vec_b[g * f.d + kk], wv, acc[g]);
      }
#pragma unroll
      for (int g = 0; g < GMAX; ++g) s.part[((size_t)sl * GMAX + g) * ncol_z + col] = acc[g];
      __syncthreads();
      for (int o = tid; o < ncol_z * gcount; o += T) {
        const int g = o / ncol_z, c2 = o - g * ncol_z;
        float v = 0.f;
        for (int s2 = 0; s2 < slices; ++s2) v += s.part[((size_t)s2 * GMAX + g) * ncol_z + c2];
        const int mod = c2 >= f.cq, cl = c2 - mod * f.cq;
        s.addv[(g * 2 + mod) * f.cq + cl] = v / (float)f.hw;  // MeanBackward: grad / HW
      }
    }
    GML_STAMP(6);
    __syncthreads();
    GML_STAMP(7);
    // ---- pass 2: d_input = grad_out * scale + ds / HW, refill with the next group ---------------
    const int next = grp + n_clusters;
    const int next_n0 = next * f.g;
    const int next_gcount = next < f.n_groups ? min(f.g, f.n - next_n0) : 0;
    for (int p = grp_in_pass; p < vplanes; p += kPlanesPerPass) {
      int g, mod, cl;
      plane_coords(f, p, g, mod, cl);
      const float sc = s.scale[p], ad = s.addv[p];
      const float4* v = reinterpret_cast<const float4*>(s.data + (size_t)p * f.hw);
      float4* o = reinterpret_cast<float4*>((mod ? a.d_b : a.d_a) +
                                            ((size_t)(n0 + g) * f.c + (size_t)rank * f.cq + cl) * f.hw);
#pragma unroll 8
      for (int i = lane; i < hw4; i += L) {
        float4 x = v[i];
        x.x = fmaf(x.x, sc, ad); x.y = fmaf(x.y, sc, ad); x.z = fmaf(x.z, sc, ad); x.w = fmaf(x.w, sc, ad);
        stg_stream(o + i, x);
      }
      if (L < 32) __syncwarp();
      if (lane == 0) plane_drained(f, s, p, a.go_a, a.go_b, rank, next_n0, next_gcount, pol_stream);
    }
    GML_STAMP(8);
    __syncthreads();
    GML_STAMP(9);
    parity ^= 1;
    ++iter;
  }
  cluster.sync();
}

// ---- host side ------------------------------------------------------------------------------
// cs = 4 -> one CTA per SM (~200 KB of planes), cs = 8 -> two CTAs per SM (~100 KB each)
bool make_cfg_cs(int n, int c, int hw, int d, int cs, int threads, FusedCfg* out) {
  if (n <= 0 || c % (4 * cs) != 0 || d % (4 * cs) != 0 || hw % 4 != 0) return false;
  FusedCfg f;
  f.n = n; f.c = c; f.hw = hw; f.d = d; f.cs = cs; f.threads = threads;
  f.cq = c / cs; f.dq = d / cs;
  const size_t slice = (size_t)2 * f.cq * hw * sizeof(float);  // one sample, both modalities, this CTA
  const size_t budget = cs == 4 ? 204800 : 102400;
  if (slice == 0 || slice > budget) return false;
  f.g = (int)(budget / slice);
  if (f.g > 2) f.g = 2;  // GMAX
  if (f.g > n) f.g = n;
  f.pl = f.g * 2 * f.cq;
  if (f.pl > threads || f.dq * f.g > threads) return false;  // one thread per plane / hidden unit in the small steps
  // transposed GEMVs map one thread per (slice, column): the column counts must divide the block
  if (f.dq > threads || 2 * f.cq > threads || threads % f.dq != 0 || threads % (2 * f.cq) != 0) return false;
  // chunking: equal chunks that never straddle a (sample, modality) slice and are TMA-sized
  int pc = f.cq;
  while (f.pl / pc < 8 && pc % 2 == 0 && ((size_t)(pc / 2) * hw * 4) % 16 == 0 && (size_t)(pc / 2) * hw * 4 >= 8192)
    pc /= 2;
  f.pc = pc;
  f.nchunk = f.pl / pc;
  if (f.nchunk > kMaxChunks || f.nchunk * pc != f.pl) return false;
  if (((size_t)pc * hw * 4) % 16 != 0 || (size_t)pc * hw * 4 >= (1u << 20)) return false;  // mbarrier tx-count range
  f.n_groups = (n + f.g - 1) / f.g;
  f.data_bytes = (size_t)f.pl * hw * sizeof(float);
  f.trace = g_fused_trace;
  f.trace_first = 0;
  f.prefetch = 0;
  f.keep_planes = 0;
  f.wsm = 0;
  f.occ_trace = nullptr;
  const size_t total = f.data_bytes + extras_bytes(f, true);
  if (total > (cs == 4 ? 232448u : 115000u)) return false;
  *out = f;
  return true;
}

bool make_cfg(int n, int c, int hw, int d, FusedCfg* out) {
  const int want_cs = g_fused_cluster;
  const int thr4 = g_fused_threads ? g_fused_threads : 512;
  const int thr8 = g_fused_threads ? g_fused_threads : 256;
  if (want_cs == 4) return make_cfg_cs(n, c, hw, d, 4, thr4, out);
  if (want_cs == 8) return make_cfg_cs(n, c, hw, d, 8, thr8 > 256 ? 256 : thr8, out);
  if (make_cfg_cs(n, c, hw, d, 8, 256, out)) return true;
  return make_cfg_cs(n, c, hw, d, 4, thr4, out);
}

int lanes_for(int hw) {
  const int items = hw / 4;
  int l = 8;
  while (l < 32 && items > l * 8) l <<= 1;
  return l;
}

template <typename Args, typename K>
int do_launch(K kern, const Args& args, const FusedCfg& f, bool bwd, cudaStream_t st, int tag) {
  const size_t smem = f.data_bytes + extras_bytes(f, bwd);
  GML_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(f.threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = f.cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  // persistent grid: as many clusters as can be co-resident, capped by the work
  cfg.gridDim = dim3(kNumSMs * 2 / f.cs * f.cs);
  static std::mutex mu;
  static std::map<std::pair<const void*, size_t>, int> cache;
  int max_clusters = 0;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_pair((const void*)kern, smem * 16 + (size_t)f.cs);
    auto it = cache.find(key);
    if (it != cache.end()) {
      max_clusters = it->second;
    } else {
      if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess || max_clusters <= 0) {
        cudaGetLastError();
        max_clusters = f.cs == 4 ? 33 : 28;  // conservative fallbacks
      }
      cache[key] = max_clusters;
    }
  }
  int clusters = f.n_groups < max_clusters ? f.n_groups : max_clusters;
  if (clusters < 1) clusters = 1;
  cfg.gridDim = dim3(clusters * f.cs);
  {
    LaunchScope ls(tag, st);
    GML_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, args, f));
  }
  GML_LAUNCH_CHECK();
  return GML_OK;
}

template <int T, typename Args>
int dispatch_fwd(const Args& args, const FusedCfg& f, cudaStream_t st) {
  const int l = lanes_for(f.hw);
#define GML_FWD(LL, GG) return do_launch(fused_fwd_kernel<T, LL, GG>, args, f, false, st, kTagFusedFwd)
  if (f.g == 1) { if (l == 32) GML_FWD(32, 1); if (l == 16) GML_FWD(16, 1); GML_FWD(8, 1); }
  if (l == 32) GML_FWD(32, 2);
  if (l == 16) GML_FWD(16, 2);
  GML_FWD(8, 2);
#undef GML_FWD
}
template <int T, typename Args>
int dispatch_bwd(const Args& args, const FusedCfg& f, cudaStream_t st) {
  const int l = lanes_for(f.hw);
#define GML_BWD(LL, GG) return do_launch(fused_bwd_kernel<T, LL, GG>, args, f, true, st, kTagFusedBwd)
  if (f.g == 1) { if (l == 32) GML_BWD(32, 1); if (l == 16) GML_BWD(16, 1); GML_BWD(8, 1); }
  if (l == 32) GML_BWD(32, 2);
  if (l == 16) GML_BWD(16, 2);
  GML_BWD(8, 2);
#undef GML_BWD
}


// =============================================================================================
// "L2-resident" variant: same cluster decomposition and exchange, but the planes are NOT parked
// in shared memory.  Pass 1 streams them from HBM with an L2 evict_last hint, pass 2 re-reads them
// (L2 hits, demoted with evict_first) a few microseconds later.  Shared memory use is a few KB, so
// four CTAs (four independent groups) share an SM and the FC / cluster-barrier chain of one group
// hides behind the streaming of the others.  One cluster per group, no persistent loop.
// The in-flight footprint is (resident CTAs) x (slice per CTA) ~ 60 MB of the 126 MB L2.
// =============================================================================================
struct L2Smem {
  float* vec_a; float* vec_b; float* psum; float* scale; float* addv; float* bias_h; float* bias_g; float* part;
};
__host__ __device__ inline size_t l2_small_bytes(const FusedCfg& f, bool bwd) {
  size_t b = (size_t)f.g * 2 * f.c * 4 + (size_t)f.g * f.d * 4 + 3 * (size_t)f.pl * 4 + ((size_t)f.dq + 2 * f.cq) * 4;
  if (bwd) b += (size_t)f.threads * f.g * 4;
  return (b + 16 + 127) / 128 * 128;
}
// weight slices a CTA keeps in shared memory (floats): first FC 2C x D/cs, second FC 2C/cs x D, in either direction
__host__ __device__ inline size_t l2_w1_floats(const FusedCfg& f) { return (f.wsm & 1) ? (size_t)2 * f.c * f.dq : 0; }
__host__ __device__ inline size_t l2_w2_floats(const FusedCfg& f) { return (f.wsm & 2) ? (size_t)2 * f.cq * f.d : 0; }
__host__ __device__ inline size_t l2_smem_bytes(const FusedCfg& f, bool bwd) {
  return l2_small_bytes(f, bwd) + (size_t)f.keep_planes * f.hw * 4 + (l2_w1_floats(f) + l2_w2_floats(f)) * 4;
}
__device__ __forceinline__ void cpa16(float* dst_smem, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cpa_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ L2Smem l2_carve(unsigned char* p, const FusedCfg& f) {
  L2Smem s;
  s.vec_a = reinterpret_cast<float*>(p); p += (size_t)f.g * 2 * f.c * 4;
  s.vec_b = reinterpret_cast<float*>(p); p += (size_t)f.g * f.d * 4;
  s.psum = reinterpret_cast<float*>(p); p += f.pl * 4;
  s.scale = reinterpret_cast<float*>(p); p += f.pl * 4;
  s.addv = reinterpret_cast<float*>(p); p += f.pl * 4;
  s.bias_h = reinterpret_cast<float*>(p); p += f.dq * 4;
  s.bias_g = reinterpret_cast<float*>(p); p += 2 * f.cq * 4;
  s.part = reinterpret_cast<float*>(p);
  return s;
}

// ---- cluster exchange without barriers: a value is stored into a peer's shared memory with st.async, which also
// counts its bytes on an mbarrier in that peer; the peer waits for "all bytes of the vector have arrived"
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr),
               "r"(__float_as_uint(v)), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// the two exchange barriers live in the last 16 bytes of the "small" region (l2_small_bytes rounds up past them)
__device__ __forceinline__ uint64_t* l2_xchg_bars(unsigned char* smem, const FusedCfg& f, bool bwd) {
  return reinterpret_cast<uint64_t*>(smem + l2_small_bytes(f, bwd) - 16);
}

__device__ __forceinline__ void occ_stamp(const FusedCfg& f, int slot) {
#ifndef GML_L2_TRACE
  (void)f; (void)slot;   // phase / residency tracing of the cluster kernels is compiled in with -DGML_L2_TRACE only
  return;
#endif
  if (f.occ_trace && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (slot == 1) {
      unsigned sm;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
      f.occ_trace[(size_t)blockIdx.x * 4 + 0] = sm;
    }
    f.occ_trace[(size_t)blockIdx.x * 4 + slot] = (long long)t;
  }
}

// HW4C != 0: the plane size (in float4) is a compile-time constant -- for 28 x 28 planes (196 vectors, 32 lanes per plane) six
// of the eight load slots of a lane are then unconditionally valid, one is valid for lanes 0-3 and one never: their
// predicates, and the instructions of the dead slot, fold away in the issue-bound streaming loops.
template <int T, int L, int GMAX, int OCC, int HW4C = 0>
__global__ void __launch_bounds__(T, OCC) l2_fwd_kernel(const FusedFwdArgs a, const FusedCfg f) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int grp = blockIdx.x / f.cs;
  const L2Smem s = l2_carve(smem_raw, f);
  // per-thread stash: a thread parks the values it loaded in pass 1 and picks the SAME elements up in
  // pass 2 -- shared memory used as extra register space, no synchronisation, no L2 re-read
  float4* stash = reinterpret_cast<float4*>(smem_raw + l2_small_bytes(f, false));
  const int tid = threadIdx.x;
  occ_stamp(f, 1);
  uint64_t* xbar = l2_xchg_bars(smem_raw, f, false);
  {
    const int gc = min(f.g, f.n - grp * f.g);
    if (tid == 0) {
      mbar_init(&xbar[0], 1);
      mbar_init(&xbar[1], 1);
      fence_mbar_init();
      mbar_expect_tx(&xbar[0], (uint32_t)(gc * 2 * f.c * 4));   // the whole squeeze vector(s), from all CTAs
      mbar_expect_tx(&xbar[1], (uint32_t)(gc * f.d * 4));       // the whole hidden vector(s)
    }
    __syncwarp();
    cluster_arrive();  // matched by cluster_wait() before the first remote store: peers' barriers are initialised by then
  }
  const uint64_t pol_keep = policy_evict_last(), pol_drop = policy_evict_first();
  for (int i = tid; i < f.dq; i += T) s.bias_h[i] = __ldg(a.b_sq + rank * f.dq + i);
  for (int i = tid; i < 2 * f.cq; i += T)
    s.bias_g[i] = __ldg((i < f.cq ? a.b_v : a.b_s) + rank * f.cq + (i < f.cq ? i : i - f.cq));
  // This CTA's rows of the FC weights (W_sq rows [rank*dq, +dq), W_v / W_s rows [rank*cq, +cq): contiguous blocks)
  // start their way into shared memory now and land while pass 1 streams: the GEMVs of the chain then run without an
  // L2 round trip, which under load costs ~3000 cycles each (profiles/r1_l2_resident_phase_trace.txt: fc1, fc2)
  float* wsm1 = reinterpret_cast<float*>(smem_raw + l2_small_bytes(f, false) + (size_t)f.keep_planes * f.hw * 4);
  float* wsm2 = wsm1 + l2_w1_floats(f);
  if (f.wsm & 1) {
    const float* src = a.w_sq + (size_t)rank * f.dq * 2 * f.c;
    for (int i = tid * 4; i < f.dq * 2 * f.c; i += T * 4) cpa16(wsm1 + i, src + i);
  }
  if (f.wsm & 2) {
    const int nw = f.cq * f.d;
    const float* sv = a.w_v + (size_t)rank * nw;
    const float* ss = a.w_s + (size_t)rank * nw;
    for (int i = tid * 4; i < nw; i += T * 4) { cpa16(wsm2 + i, sv + i); cpa16(wsm2 + nw + i, ss + i); }
  }

  constexpr int kPlanesPerPass = T / L;
  const int lane = tid % L, grp_in_pass = tid / L;
  const int hw4 = HW4C ? HW4C : (f.hw >> 2);
  const int n0 = grp * f.g;
  const int gcount = min(f.g, f.n - n0);
  const int vplanes = gcount * 2 * f.cq;
  // phase stamps: trace the CTAs of the LAST-launched clusters' neighbours too -> use a mid-grid window
  const int iter = 0;
  const bool stamp_me = f.trace && threadIdx.x == 0 && blockIdx.x >= f.trace_first && blockIdx.x < f.trace_first + 8;
#ifdef GML_L2_TRACE
#define GML_STAMP2(k) do { if (stamp_me) f.trace[((size_t)(blockIdx.x - f.trace_first) * 16 + iter) * 16 + (k)] = clock64(); } while (0)
#else
#define GML_STAMP2(k) do { (void)stamp_me; (void)iter; } while (0)
#endif
  GML_STAMP2(0);

  // ---- pass 1: plane sums from HBM, lines asked to stay in L2 --------------------------------
  for (int p = grp_in_pass; p < vplanes; p += kPlanesPerPass) {
    int g, mod, cl;
    plane_coords(f, p, g, mod, cl);
    const float4* xv = reinterpret_cast<const float4*>((mod ? a.b : a.a) +
                                                       ((size_t)(n0 + g) * f.c + (size_t)rank * f.cq + cl) * f.hw);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const bool kept = p < f.keep_planes;
    const uint64_t pol = kept ? pol_drop : pol_keep;  // stashed planes need no L2 residency
    for (int i0 = 0; i0 < hw4; i0 += 8 * L) {
      float4 x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + lane + u * L;
        x[u] = i < hw4 ? ldg_hint(xv + i, pol) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + lane + u * L;
        if (kept && i < hw4) stash[(size_t)p * hw4 + i] = x[u];
        a0 += x[u].x; a1 += x[u].y; a2 += x[u].z; a3 += x[u].w;
      }
    }
    const float t = group_sum<L>((a0 + a1) + (a2 + a3));
    if (lane == 0) s.psum[p] = t;
  }
  GML_STAMP2(1);
  if (f.wsm) cpa_wait_all();  // this thread's weight copies; the barrier below publishes them to the block
  // The squeeze vector travels to every CTA of the cluster as st.async stores counted on the receiver's mbarrier: no
  // cluster barrier and no MEMBAR on the chain (a cluster.sync() here cost ~2000 cycles, profiles/r2_cluster_kernels.md).
  // z and h go out to global memory only after the chain.
  __syncthreads();
  cluster_wait();
  GML_STAMP2(2);
  {
    const uint32_t va = smem_u32(s.vec_a), xb = smem_u32(&xbar[0]);
    for (int p = tid; p < vplanes; p += T) {
      int g, mod, cl;
      plane_coords(f, p, g, mod, cl);
      const int k = mod * f.c + rank * f.cq + cl;
      const float mean = s.psum[p] / (float)f.hw;
      for (int dst = 0; dst < f.cs; ++dst)
        st_async_f32(mapa_u32(va + (uint32_t)(g * 2 * f.c + k) * 4u, dst), mean, mapa_u32(xb, dst));
    }
  }
  mbar_wait(&xbar[0], 0);
  GML_STAMP2(3);
  {
    // FC1 rows of this CTA: shared-memory copy or global memory (row pointer = base + r * 2C either way)
    const float* w1 = (f.wsm & 1) ? wsm1 : a.w_sq + (size_t)rank * f.dq * 2 * f.c;
    const uint32_t vb0 = smem_u32(s.vec_b), xb = smem_u32(&xbar[1]);
    gemv_rows<T, GMAX>([&](int r) { return w1 + (size_t)r * 2 * f.c; }, f.dq, 2 * f.c, s.vec_a, 2 * f.c, gcount,
                       [&](int r, const float* acc) {
                         const int dd = rank * f.dq + r;
                         const float bias = s.bias_h[r];
                         for (int g = 0; g < gcount; ++g) {
                           const float hval = fmaxf(acc[g] + bias, 0.f);
                           const uint32_t vb = vb0 + (uint32_t)(g * f.d + dd) * 4u;
                           for (int dst = 0; dst < f.cs; ++dst) st_async_f32(mapa_u32(vb, dst), hval, mapa_u32(xb, dst));
                         }
                       });
  }
  GML_STAMP2(4);
  mbar_wait(&xbar[1], 0);
  GML_STAMP2(5);
  {
    // FC2 rows: [W_v rows ; W_s rows] of this CTA's channels, contiguous in the shared-memory copy
    const float* wv = (f.wsm & 2) ? wsm2 : a.w_v + (size_t)rank * f.cq * f.d;
    const float* ws = (f.wsm & 2) ? wsm2 + (size_t)f.cq * f.d : a.w_s + (size_t)rank * f.cq * f.d;
    gemv_rows<T, GMAX>([&](int r) { return r < f.cq ? wv + (size_t)r * f.d : ws + (size_t)(r - f.cq) * f.d; }, 2 * f.cq,
                       f.d, s.vec_b, f.d, gcount, [&](int r, const float* acc) {
                         const int mod = r >= f.cq, cl = r - mod * f.cq;
                         const int ch = rank * f.cq + cl;
                         const float bias = s.bias_g[r];
                         float* gout = mod ? a.g_b : a.g_a;
                         for (int g = 0; g < gcount; ++g) {
                           const float gate = sigmoidf_ref(acc[g] + bias);
                           s.scale[(g * 2 + mod) * f.cq + cl] = gate * a.gate_scale;
                           gout[(size_t)(n0 + g) * f.c + ch] = gate;
                         }
                       });
  }
  GML_STAMP2(6);
  __syncthreads();
  GML_STAMP2(7);
  // deferred exports of this CTA's share of z and h (fire and forget: nothing below waits for them)
  for (int p = tid; p < vplanes; p += T) {
    int g, mod, cl;
    plane_coords(f, p, g, mod, cl);
    a.z[(size_t)(n0 + g) * 2 * f.c + mod * f.c + rank * f.cq + cl] = s.psum[p] / (float)f.hw;
  }
  for (int i = tid; i < f.dq * gcount; i += T) {
    const int g = i / f.dq, dd = rank * f.dq + (i - g * f.dq);
    a.h[(size_t)(n0 + g) * f.d + dd] = s.vec_b[g * f.d + dd];
  }
  // ---- pass 2: re-read (L2), gate, stream out ---------------------------------------------------
  for (int p = grp_in_pass; p < vplanes; p += kPlanesPerPass) {
    int g, mod, cl;
    plane_coords(f, p, g, mod, cl);
    const float sc = s.scale[p];
    const size_t off = ((size_t)(n0 + g) * f.c + (size_t)rank * f.cq + cl) * f.hw;
    const float4* xv = reinterpret_cast<const float4*>((mod ? a.b : a.a) + off);
    float4* o = reinterpret_cast<float4*>((mod ? a.b_out : a.a_out) + off);
    const bool kept = p < f.keep_planes;
    for (int i0 = 0; i0 < hw4; i0 += 8 * L) {
      float4 x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + lane + u * L;
        if (i < hw4) x[u] = kept ? stash[(size_t)p * hw4 + i] : ldg_hint(xv + i, pol_drop);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + lane + u * L;
        if (i < hw4) {
          x[u].x *= sc; x[u].y *= sc; x[u].z *= sc; x[u].w *= sc;
          stg_stream(o + i, x[u]);
        }
      }
    }
  }
  GML_STAMP2(8);
  GML_STAMP2(9);
#ifdef GML_L2_TRACE
  __syncthreads();  // the end stamp is the CTA's, not warp 0's
  occ_stamp(f, 2);
#endif
  // no trailing cluster barrier: remote shared-memory writes only happen before the third barrier
#undef GML_STAMP2
}

// STASH = false: no per-thread stash code in the streaming loops at all.  The loops are issue-bound (57 % of the issue
// slots busy), and with the weight slices in shared memory the backward gains more from the shorter loops than it loses
// in L2 re-reads (8-CTA clusters, batch 1024: 0.461 -> 0.451 ms); the forward keeps its stash (0.314 vs 0.321 ms).
template <int T, int L, int GMAX, int OCC, bool STASH, int HW4C = 0>
__global__ void __launch_bounds__(T, OCC) l2_bwd_kernel(const FusedBwdArgs a, const FusedCfg f) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int grp = blockIdx.x / f.cs;
  const L2Smem s = l2_carve(smem_raw, f);
  float4* stash = reinterpret_cast<float4*>(smem_raw + l2_small_bytes(f, true));  // see l2_fwd_kernel
  const int tid = threadIdx.x;
  occ_stamp(f, 1);
  uint64_t* xbar = l2_xchg_bars(smem_raw, f, true);
  {
    const int gc = min(f.g, f.n - grp * f.g);
    if (tid == 0) {
      mbar_init(&xbar[0], 1);
      mbar_init(&xbar[1], 1);
      fence_mbar_init();
      mbar_expect_tx(&xbar[0], (uint32_t)(gc * 2 * f.c * 4));   // dE of both modalities, from all CTAs
      mbar_expect_tx(&xbar[1], (uint32_t)(gc * f.d * 4));       // dH
    }
    __syncwarp();
    cluster_arrive();
    for (int i = tid; i < T * GMAX; i += T) s.part[i] = 0.f;   // (published by the __syncthreads after pass 1)
  }
  const uint64_t pol_keep = policy_evict_last(), pol_drop = policy_evict_first();
  constexpr int kPlanesPerPass = T / L;
  const int lane = tid % L, grp_in_pass = tid / L;
  const int hw4 = HW4C ? HW4C : (f.hw >> 2);
  const int ncol_h = f.dq, ncol_z = 2 * f.cq;
  const int n0 = grp * f.g;
  const int gcount = min(f.g, f.n - n0);
  const int vplanes = gcount * 2 * f.cq;
  const int iter = 0;
  const bool stamp_me = f.trace && threadIdx.x == 0 && blockIdx.x >= f.trace_first && blockIdx.x < f.trace_first + 8;
#ifdef GML_L2_TRACE
#define GML_STAMP2(k) do { if (stamp_me) f.trace[((size_t)(blockIdx.x - f.trace_first) * 16 + iter) * 16 + (k)] = clock64(); } while (0)
#else
#define GML_STAMP2(k) do { (void)stamp_me; (void)iter; } while (0)
#endif
  GML_STAMP2(0);

  float gate_pf = 0.f, h_pf = 0.f;
  if (tid < vplanes) {
    int g, mod, cl;
    plane_coords(f, tid, g, mod, cl);
    gate_pf = __ldg((mod ? a.g_b : a.g_a) + (size_t)(n0 + g) * f.c + rank * f.cq + cl);
  }
  if (tid < ncol_h * gcount) {
    const int g = tid / ncol_h, col = tid - g * ncol_h;
    h_pf = __ldg(a.h + (size_t)(n0 + g) * f.d + rank * f.dq + col);
  }
  // weight slices into shared memory behind pass 1 (see l2_fwd_kernel): wsm1 = [W_v[:, cols] ; W_s[:, cols]] as
  // [2C][dq] (cols = this CTA's dq hidden units), wsm2 = W_sq[:, this CTA's 2 cq squeeze columns] as [D][2 cq]
  float* wsm1 = reinterpret_cast<float*>(smem_raw + l2_small_bytes(f, true) + (size_t)f.keep_planes * f.hw * 4);
  float* wsm2 = wsm1 + l2_w1_floats(f);
  if (f.wsm & 1) {
    const int qpr = f.dq >> 2;  // 16-byte pieces per row
    for (int i = tid; i < 2 * f.c * qpr; i += T) {
      const int k = i / qpr, q = i - k * qpr;
      const float* src = (k < f.c ? a.w_v + (size_t)k * f.d : a.w_s + (size_t)(k - f.c) * f.d) + rank * f.dq + 4 * q;
      cpa16(wsm1 + (size_t)k * f.dq + 4 * q, src);
    }
  }
  if (f.wsm & 2) {
    const int qpr = ncol_z >> 2;
    for (int i = tid; i < f.d * qpr; i += T) {
      const int k = i / qpr, col = 4 * (i - k * qpr);
      const int gcol = col < f.cq ? rank * f.cq + col : f.c + rank * f.cq + (col - f.cq);
      cpa16(wsm2 + (size_t)k * ncol_z + col, a.w_sq + (size_t)k * 2 * f.c + gcol);
    }
  }
  // ---- pass 1: <grad_out (kept in L2), input (streamed once)> per plane --------------------------
  for (int p = grp_in_pass; p < vplanes; p += kPlanesPerPass) {
    int g, mod, cl;
    plane_coords(f, p, g, mod, cl);
    const size_t off = ((size_t)(n0 + g) * f.c + (size_t)rank * f.cq + cl) * f.hw;
    const float4* gv = reinterpret_cast<const float4*>((mod ? a.go_b : a.go_a) + off);
    const float4* xv = reinterpret_cast<const float4*>((mod ? a.b : a.a) + off);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const bool kept = STASH && p < f.keep_planes;
    const uint64_t polg = kept ? pol_drop : pol_keep;
    for (int i0 = 0; i0 < hw4; i0 += 4 * L) {
      float4 x[4], gg[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + lane + u * L;
        if (i < hw4) {
          gg[u] = ldg_hint(gv + i, polg);
          x[u] = ldg_hint(xv + i, pol_drop);
        } else {
          gg[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          x[u] = gg[u];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + lane + u * L;
        if (kept && i < hw4) stash[(size_t)p * hw4 + i] = gg[u];
        a0 = fmaf(gg[u].x, x[u].x, a0); a1 = fmaf(gg[u].y, x[u].y, a1);
        a2 = fmaf(gg[u].z, x[u].z, a2); a3 = fmaf(gg[u].w, x[u].w, a3);
      }
    }
    const float t = group_sum<L>((a0 + a1) + (a2 + a3));
    if (lane == 0) s.psum[p] = t;
  }
  GML_STAMP2(1);
  if (f.wsm) cpa_wait_all();
  // dE of this CTA's channels -> every CTA of the cluster (st.async + mbarrier, see l2_fwd_kernel); dE and dH are
  // exported to global memory after the chain
  __syncthreads();
  cluster_wait();
  GML_STAMP2(2);
  if (tid < vplanes) {
    const int p = tid;
    int g, mod, cl;
    plane_coords(f, p, g, mod, cl);
    const int ch = rank * f.cq + cl;
    const float de = s.psum[p] * a.gate_scale * gate_pf * (1.f - gate_pf);
    s.scale[p] = gate_pf * a.gate_scale;
    const uint32_t va = smem_u32(s.vec_a) + (uint32_t)(g * 2 * f.c + mod * f.c + ch) * 4u, xb = smem_u32(&xbar[0]);
    for (int dst = 0; dst < f.cs; ++dst) st_async_f32(mapa_u32(va, dst), de, mapa_u32(xb, dst));
  }
  mbar_wait(&xbar[0], 0);
  GML_STAMP2(3);
  {
    // [W_v[:, cols] ; W_s[:, cols]]: shared-memory copy ([2C][dq]) or global memory ([C][D] each, column offset rank*dq)
    if (f.wsm & 1) {
      gemv_cols_partial<T, GMAX, true>(wsm1, f.dq, 0, ncol_h, 0, f.c, s.vec_a, 2 * f.c, gcount, s.part);
      gemv_cols_partial<T, GMAX, true>(wsm1 + (size_t)f.c * f.dq, f.dq, 0, ncol_h, 0, f.c, s.vec_a + f.c, 2 * f.c,
                                       gcount, s.part);
    } else {
      gemv_cols_partial<T, GMAX>(a.w_v, f.d, rank * f.dq, ncol_h, 0, f.c, s.vec_a, 2 * f.c, gcount, s.part);
      gemv_cols_partial<T, GMAX>(a.w_s, f.d, rank * f.dq, ncol_h, 0, f.c, s.vec_a + f.c, 2 * f.c, gcount, s.part);
    }
  }
  __syncthreads();
  {
    const int slices = T / ncol_h;
    if (tid < ncol_h * gcount) {
      const int g = tid / ncol_h, col = tid - g * ncol_h;
      float v = 0.f;
      for (int sl = 0; sl < slices; ++sl) v += s.part[((size_t)sl * GMAX + g) * ncol_h + col];
      const int dd = rank * f.dq + col;
      v = h_pf > 0.f ? v : 0.f;
      const uint32_t vb = smem_u32(s.vec_b) + (uint32_t)(g * f.d + dd) * 4u, xb = smem_u32(&xbar[1]);
      for (int dst = 0; dst < f.cs; ++dst) st_async_f32(mapa_u32(vb, dst), v, mapa_u32(xb, dst));
    }
  }
  GML_STAMP2(4);
  mbar_wait(&xbar[1], 0);
  GML_STAMP2(5);
  {
    const int slices = T / ncol_z;
    const int col = tid % ncol_z, sl = tid / ncol_z;
    const int gcol = (col < f.cq) ? rank * f.cq + col : f.c + rank * f.cq + (col - f.cq);
    const int span = (f.d + slices - 1) / slices;
    const int ka = sl * span, kb = min(f.d, ka + span);
    float acc[GMAX];
#pragma unroll
    for (int g = 0; g < GMAX; ++g) acc[g] = 0.f;
    if (f.wsm & 2) {
#pragma unroll 16
      for (int kk = ka; kk < kb; ++kk) {
        const float wv = wsm2[(size_t)kk * ncol_z + col];
#pragma unroll
        for (int g = 0; g < GMAX; ++g)
          if (g < gcount) acc[g] = fmaf(s.vec_b[g * f.d + kk], wv, acc[g]);
      }
    } else {
#pragma unroll 16
      for (int kk = ka; kk < kb; ++kk) {
        const float wv = __ldg(a.w_sq + (size_t)kk * 2 * f.c + gcol);
#pragma unroll
        for (int g = 0; g < GMAX; ++g)
          if (g < gcount) acc[g] = fmaf(s.vec_b[g * f.d + kk], wv, acc[g]);
      }
    }
#pragma unroll
    for (int g = 0; g < GMAX; ++g) s.part[((size_t)sl * GMAX + g) * ncol_z + col] = acc[g];
    __syncthreads();
    for (int o = tid; o < ncol_z * gcount; o += T) {
      const int g = o / ncol_z, c2 = o - g * ncol_z;
      float v = 0.f;
      for (int s2 = 0; s2 < slices; ++s2) v += s.part[((size_t)s2 * GMAX + g) * ncol_z + c2];
      const int mod = c2 >= f.cq, cl = c2 - mod * f.cq;
      s.addv[(g * 2 + mod) * f.cq + cl] = v / (float)f.hw;
    }
  }
  GML_STAMP2(6);
  __syncthreads();
  GML_STAMP2(7);
  if (tid < vplanes) {   // deferred exports: this CTA's share of dE and dH
    int g, mod, cl;
    plane_coords(f, tid, g, mod, cl);
    const int ch = rank * f.cq + cl;
    (mod ? a.de_b : a.de_a)[(size_t)(n0 + g) * f.c + ch] = s.vec_a[g * 2 * f.c + mod * f.c + ch];
  }
  if (tid < ncol_h * gcount) {
    const int g = tid / ncol_h, dd = rank * f.dq + (tid - g * ncol_h);
    a.dh[(size_t)(n0 + g) * f.d + dd] = s.vec_b[g * f.d + dd];
  }
  // ---- pass 2: d_input = grad_out (L2 re-read) * scale + ds / HW ---------------------------------
  for (int p = grp_in_pass; p < vplanes; p += kPlanesPerPass) {
    int g, mod, cl;
    plane_coords(f, p, g, mod, cl);
    const float sc = s.scale[p], ad = s.addv[p];
    const size_t off = ((size_t)(n0 + g) * f.c + (size_t)rank * f.cq + cl) * f.hw;
    const float4* gv = reinterpret_cast<const float4*>((mod ? a.go_b : a.go_a) + off);
    float4* o = reinterpret_cast<float4*>((mod ? a.d_b : a.d_a) + off);
    const bool kept = STASH && p < f.keep_planes;
    for (int i0 = 0; i0 < hw4; i0 += 8 * L) {
      float4 x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + lane + u * L;
        if (i < hw4) x[u] = kept ? stash[(size_t)p * hw4 + i] : ldg_hint(gv + i, pol_drop);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + lane + u * L;
        if (i < hw4) {
          x[u].x = fmaf(x[u].x, sc, ad); x[u].y = fmaf(x[u].y, sc, ad);
          x[u].z = fmaf(x[u].z, sc, ad); x[u].w = fmaf(x[u].w, sc, ad);
          stg_stream(o + i, x[u]);
        }
      }
    }
  }
  GML_STAMP2(8);
  GML_STAMP2(9);
#ifdef GML_L2_TRACE
  __syncthreads();
  occ_stamp(f, 2);
#endif
  // no trailing cluster barrier (see forward)
#undef GML_STAMP2
}

bool make_cfg_l2(int n, int c, int hw, int d, int cs, bool bwd, FusedCfg* out) {
  if (n <= 0 || c % (4 * cs) != 0 || d % (4 * cs) != 0 || hw % 4 != 0) return false;
  FusedCfg f;
  f.n = n; f.c = c; f.hw = hw; f.d = d; f.cs = cs; f.threads = 256;
  f.cq = c / cs; f.dq = d / cs;
  const size_t slice = (size_t)2 * f.cq * hw * sizeof(float);
  f.g = slice * 2 <= (size_t)g_fused_group_kb * 1024 ? 2 : 1;  // two samples per cluster while the CTA's share stays small
  if (f.g > n) f.g = n;
  f.pl = f.g * 2 * f.cq;
  if (f.pl > f.threads || f.dq * f.g > f.threads) return false;
  if (f.dq > f.threads || 2 * f.cq > f.threads || f.threads % f.dq != 0 || f.threads % (2 * f.cq) != 0) return false;
  f.pc = f.pl; f.nchunk = 1;
  f.n_groups = (n + f.g - 1) / f.g;
  f.data_bytes = 0;
  f.trace = g_fused_trace;
  f.occ_trace = g_fused_occ_trace;
  f.trace_first = (f.n_groups / 2) * cs;
  f.prefetch = ((size_t)hw * 4) % 16 == 0 ? g_fused_prefetch : 0;  // look-ahead distance in groups
  // shared memory per CTA: [exchange buffers][stash][weight slices].  Weight slices (2 x 2CD/cs floats) are taken when
  // they still leave >= 8 KB of stash at four CTAs per SM; a forced mask may cost occupancy instead.
  f.wsm = 0; f.keep_planes = 0;
  const size_t small = l2_small_bytes(f, bwd), plane = (size_t)hw * 4;
  const size_t w1 = (size_t)2 * c * f.dq * 4, w2 = (size_t)2 * f.cq * d * 4;
  const size_t lim4 = 46 * 1024;  // measured (profiles/r2_sweep.md): 44.3 KB per CTA runs at full speed, 49.4 KB does not
  int mask = g_fused_wsmem;
  if (mask < 0) mask = (small + w1 + w2 + 8 * 1024 <= lim4) ? 3 : 0;
  const size_t wbytes = ((mask & 1) ? w1 : 0) + ((mask & 2) ? w2 : 0);
  size_t stash = g_fused_stash_kb >= 0 ? (size_t)g_fused_stash_kb * 1024 : ((bwd && (mask & 3) == 3) ? 0 : 24 * 1024);
  if (wbytes) {
    size_t lim = lim4;
    if (small + wbytes > lim) lim = 75 * 1024;    // 3 CTAs per SM
    if (small + wbytes > lim) lim = 113 * 1024;   // 2
    if (small + wbytes > lim) lim = 227 * 1024;   // 1
    if (small + wbytes > lim) return false;
    if (stash > lim - small - wbytes) stash = lim - small - wbytes;
  }
  f.wsm = mask & 3;
  f.keep_planes = (int)(stash / plane);
  if (f.keep_planes > f.pl) f.keep_planes = f.pl;
  if (l2_smem_bytes(f, bwd) > 227 * 1024) return false;
  if ((long long)f.n_groups * cs > 0x7fffffffLL) return false;
  *out = f;
  return true;
}

template <typename Args, typename K>
int do_launch_l2(K kern, const Args& args, const FusedCfg& f, bool bwd, cudaStream_t st, int tag) {
  const size_t smem = l2_smem_bytes(f, bwd);
  {
    static const bool dbg = getenv("GML_DEBUG_CFG") != nullptr;
    if (dbg)
      fprintf(stderr, "[gml] l2 %s: cs=%d g=%d wsm=%d (tunable %d) keep=%d smem=%zu stash_kb=%d\n", bwd ? "bwd" : "fwd", f.cs,
              f.g, f.wsm, g_fused_wsmem, f.keep_planes, smem, g_fused_stash_kb);
  }
  if (smem > 48 * 1024) GML_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (f.cs > 8) GML_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(f.threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = f.cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cfg.gridDim = dim3((unsigned)f.n_groups * f.cs);
  {
    LaunchScope ls(tag, st);
    GML_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, args, f));
  }
  GML_LAUNCH_CHECK();
  return GML_OK;
}
template <typename Args>
int dispatch_l2_fwd(const Args& args, const FusedCfg& f, cudaStream_t st) {
  const int l = lanes_for(f.hw);
#define GML_L2F(LL, GG)                                                                              \
  do {                                                                                               \
    if (g_fused_occ == 5) return do_launch_l2(l2_fwd_kernel<256, LL, GG, 5>, args, f, false, st, kTagFusedFwd); \
    return do_launch_l2(l2_fwd_kernel<256, LL, GG, 4>, args, f, false, st, kTagFusedFwd);             \
  } while (0)
  if (g_fused_hw_special && l == 32 && f.hw == 784 && f.g == 1 && g_fused_occ != 5)   // 28 x 28 planes
    return do_launch_l2(l2_fwd_kernel<256, 32, 1, 4, 196>, args, f, false, st, kTagFusedFwd);
  if (f.g == 1) { if (l == 32) GML_L2F(32, 1); if (l == 16) GML_L2F(16, 1); GML_L2F(8, 1); }
  if (l == 32) GML_L2F(32, 2);
  if (l == 16) GML_L2F(16, 2);
  GML_L2F(8, 2);
#undef GML_L2F
}
template <typename Args>
int dispatch_l2_bwd(const Args& args, const FusedCfg& f, cudaStream_t st) {
  const int l = lanes_for(f.hw);
#define GML_L2B(LL, GG)                                                                              \
  do {                                                                                               \
    if (f.keep_planes == 0) return do_launch_l2(l2_bwd_kernel<256, LL, GG, 4, false>, args, f, true, st, kTagFusedBwd); \
    if (g_fused_occ == 5) return do_launch_l2(l2_bwd_kernel<256, LL, GG, 5, true>, args, f, true, st, kTagFusedBwd); \
    return do_launch_l2(l2_bwd_kernel<256, LL, GG, 4, true>, args, f, true, st, kTagFusedBwd);        \
  } while (0)
  if (g_fused_hw_special && l == 32 && f.hw == 784 && f.g == 1 && g_fused_occ != 5) {
    if (f.keep_planes == 0) return do_launch_l2(l2_bwd_kernel<256, 32, 1, 4, false, 196>, args, f, true, st, kTagFusedBwd);
    return do_launch_l2(l2_bwd_kernel<256, 32, 1, 4, true, 196>, args, f, true, st, kTagFusedBwd);
  }
  if (f.g == 1) { if (l == 32) GML_L2B(32, 1); if (l == 16) GML_L2B(16, 1); GML_L2B(8, 1); }
  if (l == 32) GML_L2B(32, 2);
  if (l == 16) GML_L2B(16, 2);
  GML_L2B(8, 2);
#undef GML_L2B
}

}  // namespace

static bool use_l2_kind() { return g_fused_kind == 2 || g_fused_kind == 0; }
int g_fused_weight_ratio_x100 = 100;  // tunable: max (FC weight bytes) / (group feature-map bytes), in percent
static bool weights_ok(const FusedCfg& f) {
  // per-sample FC weights are re-read from L2 for every group: only worth it while they are small
  // next to the group's feature-map bytes (MMTM4's 512x7^2 goes the streaming way)
  const double w_bytes = 16.0 * f.c * f.d;
  const double group_bytes = 2.0 * f.g * 2.0 * f.c * f.hw * 4.0;
  return w_bytes * 100.0 <= group_bytes * g_fused_weight_ratio_x100;
}
static bool pick_cfg(int n, int c, int hw, int d, FusedCfg* f, bool* l2, bool bwd) {
  if (use_l2_kind()) {
    // measured on B200 (profiles/): the forward prefers 8-CTA clusters (smaller L2 footprint in flight),
    // the backward 4-CTA clusters (fewer, larger transposed GEMVs)
    // ... unless 8-CTA clusters can keep both weight slices in shared memory and the batch is large: 128x28^2 backward
    // at batch 1024 0.488 ms (cs 8, weights in shared memory) vs 0.515 (cs 4) vs 0.620 (cs 8, weights from L2); at batch
    // 128-256 cs 4 is still ahead (0.091 vs 0.095 ms at 128)
    if (bwd && !g_fused_cluster && n >= 512 && make_cfg_l2(n, c, hw, d, 8, bwd, f) && f->wsm == 3 && weights_ok(*f)) {
      *l2 = true;
      return true;
    }
    const int first = g_fused_cluster ? g_fused_cluster : (bwd ? 4 : 8);
    const int second = g_fused_cluster ? 0 : (bwd ? 8 : 4);
    if (make_cfg_l2(n, c, hw, d, first, bwd, f) && weights_ok(*f)) { *l2 = true; return true; }
    if (second && make_cfg_l2(n, c, hw, d, second, bwd, f) && weights_ok(*f)) { *l2 = true; return true; }
    if (g_fused_kind == 2) return false;
  }
  *l2 = false;
  return make_cfg(n, c, hw, d, f);
}

bool fused_supported(int n, int c_v, int c_s, int hw_v, int hw_s, int d, int mode) {
  if (mode != GML_MODE_NORMAL) return false;
  if (c_v != c_s || hw_v != hw_s) return false;
  FusedCfg f;
  bool l2 = false;
  if (!pick_cfg(n, c_v, hw_v, d, &f, &l2, false)) return false;
  if (!weights_ok(f)) return false;
  FusedCfg fb;
  return pick_cfg(n, c_v, hw_v, d, &fb, &l2, true) && weights_ok(fb);
}

// Forward only: at large batch the batched FC GEMMs (tcgen05) get cheap, while the cluster kernel re-reads the FC
// weights from L2 once per group.  Measured on B200 (profiles/r1_sweep.md): 256x14^2 at batch 1024 runs 0.297 ms
// streaming vs 0.321 ms fused; 128x28^2 (weights 16 % of a group) stays fused at every batch.
bool fused_fwd_preferred(int n, int c, int hw, int d) {
  FusedCfg f;
  bool l2 = false;
  if (!pick_cfg(n, c, hw, d, &f, &l2, false)) return false;
  const double w_bytes = 16.0 * f.c * f.d, group_bytes = 2.0 * f.g * 2.0 * f.c * f.hw * 4.0;
  return !(n >= 1024 && w_bytes > 0.5 * group_bytes);
}

int launch_fused_fwd(const FusedFwdArgs& args, cudaStream_t st) {
  FusedCfg f;
  bool l2 = false;
  if (!pick_cfg(args.n, args.c, args.hw, args.d, &f, &l2, false)) return GML_E_UNSUPPORTED;
  if (!aligned16(args.a) || !aligned16(args.b) || !aligned16(args.a_out) || !aligned16(args.b_out) ||
      !aligned16(args.w_sq) || !aligned16(args.w_v) || !aligned16(args.w_s))
    return GML_E_UNSUPPORTED;
  if (l2) return dispatch_l2_fwd(args, f, st);
  return f.threads == 512 ? dispatch_fwd<512>(args, f, st) : dispatch_fwd<256>(args, f, st);
}

int launch_fused_bwd(const FusedBwdArgs& args, cudaStream_t st) {
  FusedCfg f;
  bool l2 = false;
  if (!pick_cfg(args.n, args.c, args.hw, args.d, &f, &l2, true)) return GML_E_UNSUPPORTED;
  if (!aligned16(args.go_a) || !aligned16(args.go_b) || !aligned16(args.a) || !aligned16(args.b) ||
      !aligned16(args.d_a) || !aligned16(args.d_b))
    return GML_E_UNSUPPORTED;
  if (l2 && f.wsm && (!aligned16(args.w_sq) || !aligned16(args.w_v) || !aligned16(args.w_s))) f.wsm = 0;
  if (l2) return dispatch_l2_bwd(args, f, st);
  return f.threads == 512 ? dispatch_bwd<512>(args, f, st) : dispatch_bwd<256>(args, f, st);
}

}  // namespace gml
