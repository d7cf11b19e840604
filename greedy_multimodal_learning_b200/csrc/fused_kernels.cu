// Single-pass MMTM kernels: thread-block clusters keep a group of samples resident in shared
// memory, so the feature maps cross HBM exactly once per direction (forward 4u, backward 6u;
// u = N*C*HW*4 B per modality) instead of the 6u / 8u of the streaming two-pass path.
//
// Layout of the work
//   * a cluster of 4 CTAs (1 CTA per SM, ~210 KB dynamic shared memory each) owns a group of G
//     samples; CTA r holds channels [r*C/4, (r+1)*C/4) of BOTH modalities of those samples:
//     G * 2 * (C/4) planes of HW floats = 200,704 B for 128x28^2 (G=1) and 256x14^2 (G=2).
//   * the CTA's planes arrive as NCHUNK 1-D TMA bulk copies (cp.async.bulk, contiguous in NCHW),
//     each signalling its own mbarrier: the reduction pass consumes chunks as they land, and
//     while the output pass drains chunk j to HBM the NEXT group's chunk j is already being
//     fetched into the freed space (loads and stores overlap inside one CTA).
//   * the squeeze vector (2C floats per sample) and the hidden state (D floats) are exchanged
//     through distributed shared memory (every CTA writes its part into all four CTAs), two
//     cluster barriers per group; the FCs are per-sample GEMVs on CUDA cores with the weights
//     streamed from L2 (4C^2 floats per group: 0.16x / 0.65x of the group's HBM bytes).
//   * persistent: clusters loop over groups with stride = number of resident clusters.
// Forward: resident = inputs.  Backward: resident = grad_out (needed twice: dot, then apply);
// the saved inputs stream through registers once for the dot.  Weight gradients are NOT formed
// here: dE / dH rows go to global memory and the batched GEMMs in fc_kernels.cu reduce them over
// the batch in a fixed order.
//
// Shapes outside the supported set fall back to the streaming kernels (capi.cu).
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace gml {

namespace {

constexpr int kCluster = 4;
constexpr int kThreadsF = 256;
constexpr int kWarpsF = kThreadsF / 32;
constexpr int kMaxChunks = 16;
constexpr size_t kDataBudget = 204800;  // bytes of resident planes per CTA

struct FusedCfg {
  int n, c, hw, d;
  int g;         // samples per group
  int cq, dq;    // C/4, D/4
  int pl;        // planes per CTA = g * 2 * cq
  int pc;        // planes per chunk
  int nchunk;    // pl / pc
  int n_groups;
  size_t data_bytes;
};

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

struct Smem {
  float* data;
  uint64_t* bars;
  float* psum;    // [pl] plane sums / dots
  float* scale;   // [pl] per-plane multiplier of the output pass
  float* addv;    // [pl] per-plane additive term (backward)
  float* vec_a;   // [g][2C]  z (fwd) / dE (bwd), full vector, filled by all 4 CTAs
  float* vec_b;   // [g][D]   h (fwd) / dH (bwd)
  float* part;    // cross-warp partial sums of the transposed GEMVs (backward)
};

__device__ __forceinline__ Smem carve(unsigned char* base, const FusedCfg& f, bool bwd) {
  Smem s;
  s.data = reinterpret_cast<float*>(base);
  unsigned char* p = base + f.data_bytes;
  s.bars = reinterpret_cast<uint64_t*>(p); p += kMaxChunks * sizeof(uint64_t);
  s.psum = reinterpret_cast<float*>(p); p += f.pl * sizeof(float);
  s.scale = reinterpret_cast<float*>(p); p += f.pl * sizeof(float);
  s.addv = reinterpret_cast<float*>(p); p += f.pl * sizeof(float);
  s.vec_a = reinterpret_cast<float*>(p); p += (size_t)f.g * 2 * f.c * sizeof(float);
  s.vec_b = reinterpret_cast<float*>(p); p += (size_t)f.g * f.d * sizeof(float);
  s.part = reinterpret_cast<float*>(p);
  (void)bwd;
  return s;
}

size_t smem_bytes(const FusedCfg& f, bool bwd) {
  size_t b = f.data_bytes + kMaxChunks * sizeof(uint64_t) + 3 * (size_t)f.pl * 4 + (size_t)f.g * 2 * f.c * 4 +
             (size_t)f.g * f.d * 4;
  if (bwd) b += (size_t)kThreadsF * f.g * 4;  // part[slices][g][outs], slices * outs == 256
  return b + 16;
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

// y[row] = <W[row0 + row, 0:K], x[g, 0:K]> for rows handled warp-per-row, K % 4 == 0
template <int GMAX, typename Epi>
__device__ __forceinline__ void gemv_rows(const float* __restrict__ w, int ldw, int row0, int nrows, int k,
                                          const float* x, int ldx, int gcount, Epi epi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k4 = k >> 2;
  for (int r = warp; r < nrows; r += kWarpsF) {
    const float4* wr = reinterpret_cast<const float4*>(w + (size_t)(row0 + r) * ldw);
    float acc[GMAX];
#pragma unroll
    for (int g = 0; g < GMAX; ++g) acc[g] = 0.f;
    for (int i = lane; i < k4; i += 32) {
      const float4 wv = __ldg(wr + i);
#pragma unroll
      for (int g = 0; g < GMAX; ++g) {
        if (g < gcount) {
          const float4 xv = *reinterpret_cast<const float4*>(x + (size_t)g * ldx + 4 * i);
          acc[g] = fmaf(wv.x, xv.x, acc[g]); acc[g] = fmaf(wv.y, xv.y, acc[g]);
          acc[g] = fmaf(wv.z, xv.z, acc[g]); acc[g] = fmaf(wv.w, xv.w, acc[g]);
        }
      }
    }
#pragma unroll
    for (int g = 0; g < GMAX; ++g) acc[g] = warp_sum(acc[g]);
    if (lane == 0) epi(r, acc);
  }
}

// y[col] = sum_k x[g, k] * W[k, col0 + col] (transposed GEMV): thread = (k-slice, col), coalesced in col.
// Partials go to s_part[slice][g][col]; the caller reduces the slices in a fixed order.
template <int GMAX>
__device__ __forceinline__ void gemv_cols_partial(const float* __restrict__ w, int ldw, int col0, int ncols, int k0,
                                                  int k1, const float* x, int ldx, int gcount, float* s_part) {
  const int slices = kThreadsF / ncols;
  const int col = threadIdx.x % ncols, sl = threadIdx.x / ncols;
  const int span = (k1 - k0 + slices - 1) / slices;
  const int ka = k0 + sl * span, kb = min(k1, ka + span);
  float acc[GMAX];
#pragma unroll
  for (int g = 0; g < GMAX; ++g) acc[g] = 0.f;
  const float* wp = w + col0 + col;
#pragma unroll 4
  for (int kk = ka; kk < kb; ++kk) {
    const float wv = __ldg(wp + (size_t)kk * ldw);
#pragma unroll
    for (int g = 0; g < GMAX; ++g)
      if (g < gcount) acc[g] = fmaf(x[(size_t)g * ldx + kk], wv, acc[g]);
  }
#pragma unroll
  for (int g = 0; g < GMAX; ++g) s_part[((size_t)sl * GMAX + g) * ncols + col] += acc[g];
}

// =============================================================================================
// forward
// =============================================================================================
template <int L, int GMAX>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreadsF, 1)
    fused_fwd_kernel(const FusedFwdArgs a, const FusedCfg f) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cluster_id = blockIdx.x / kCluster, n_clusters = gridDim.x / kCluster;
  const Smem s = carve(smem_raw, f, false);
  const int tid = threadIdx.x;
  const uint64_t pol_stream = policy_evict_first();

  if (tid == 0) {
    for (int j = 0; j < f.nchunk; ++j) mbar_init(&s.bars[j], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0 && cluster_id < f.n_groups) {
    const int n0 = cluster_id * f.g;
    issue_chunks(f, s, a.a, a.b, rank, n0, min(f.g, f.n - n0), 0, f.nchunk, pol_stream);
  }
  cluster.sync();  // every CTA of the cluster is alive before anyone writes remote shared memory

  constexpr int kPlanesPerPass = kThreadsF / L;
  const int lane = tid % L, grp_in_pass = tid / L;
  const int hw4 = f.hw >> 2;
  const float inv_gate_scale_unused = 0.f; (void)inv_gate_scale_unused;
  uint32_t parity = 0;

  for (int grp = cluster_id; grp < f.n_groups; grp += n_clusters) {
    const int n0 = grp * f.g;
    const int gcount = min(f.g, f.n - n0);
    const int vplanes = gcount * 2 * f.cq;
    const int vchunks = vplanes / f.pc;

    // ---- pass 1: plane sums straight out of shared memory as chunks land --------------------
    for (int j = 0; j < vchunks; ++j) {
      mbar_wait(&s.bars[j], parity);
      for (int pp = grp_in_pass; pp < f.pc; pp += kPlanesPerPass) {
        const int p = j * f.pc + pp;
        const float4* v = reinterpret_cast<const float4*>(s.data + (size_t)p * f.hw);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
        for (int i = lane; i < hw4; i += L) {
          const float4 x = v[i];
          a0 += x.x; a1 += x.y; a2 += x.z; a3 += x.w;
        }
        const float t = group_sum<L>((a0 + a1) + (a2 + a3));
        if (lane == 0) s.psum[p] = t;
      }
    }
    __syncthreads();
    // ---- squeeze vector -> every CTA of the cluster (DSMEM) + global z ------------------------
    for (int p = tid; p < vplanes; p += kThreadsF) {
      int g, mod, cl;
      plane_coords(f, p, g, mod, cl);
      const int k = mod * f.c + rank * f.cq + cl;
      const float mean = s.psum[p] / (float)f.hw;
#pragma unroll
      for (int dst = 0; dst < kCluster; ++dst) cluster.map_shared_rank(s.vec_a, dst)[g * 2 * f.c + k] = mean;
      a.z[(size_t)(n0 + g) * 2 * f.c + k] = mean;
    }
    cluster.sync();
    // ---- FC1: my quarter of the hidden units, H = relu(Wsq z + bsq) ---------------------------
    gemv_rows<GMAX>(a.w_sq, 2 * f.c, rank * f.dq, f.dq, 2 * f.c, s.vec_a, 2 * f.c, gcount,
                    [&](int r, const float* acc) {
                      const int dd = rank * f.dq + r;
                      const float bias = __ldg(a.b_sq + dd);
                      for (int g = 0; g < gcount; ++g) {
                        const float hval = fmaxf(acc[g] + bias, 0.f);
#pragma unroll
                        for (int dst = 0; dst < kCluster; ++dst)
                          cluster.map_shared_rank(s.vec_b, dst)[g * f.d + dd] = hval;
                        a.h[(size_t)(n0 + g) * f.d + dd] = hval;
                      }
                    });
    cluster.sync();
    // ---- FC2: gates of my channels, both modalities -------------------------------------------
    for (int mod = 0; mod < 2; ++mod) {
      const float* w = mod ? a.w_s : a.w_v;
      const float* bias_v = mod ? a.b_s : a.b_v;
      float* gout = mod ? a.g_b : a.g_a;
      gemv_rows<GMAX>(w, f.d, rank * f.cq, f.cq, f.d, s.vec_b, f.d, gcount, [&](int r, const float* acc) {
        const int ch = rank * f.cq + r;
        const float bias = __ldg(bias_v + ch);
        for (int g = 0; g < gcount; ++g) {
          const float gate = sigmoidf_ref(acc[g] + bias);
          s.scale[(g * 2 + mod) * f.cq + r] = gate * a.gate_scale;
          gout[(size_t)(n0 + g) * f.c + ch] = gate;
        }
      });
    }
    __syncthreads();
    // ---- pass 2: gate from shared memory, stream out; refill freed chunks with the next group --
    const int next = grp + n_clusters;
    const int next_n0 = next * f.g;
    const int next_gcount = next < f.n_groups ? min(f.g, f.n - next_n0) : 0;
    for (int j = 0; j < vchunks; ++j) {
      for (int pp = grp_in_pass; pp < f.pc; pp += kPlanesPerPass) {
        const int p = j * f.pc + pp;
        int g, mod, cl;
        plane_coords(f, p, g, mod, cl);
        const float sc = s.scale[p];
        const float4* v = reinterpret_cast<const float4*>(s.data + (size_t)p * f.hw);
        float4* o = reinterpret_cast<float4*>((mod ? a.b_out : a.a_out) +
                                              ((size_t)(n0 + g) * f.c + (size_t)rank * f.cq + cl) * f.hw);
#pragma unroll 4
        for (int i = lane; i < hw4; i += L) {
          float4 x = v[i];
          x.x *= sc; x.y *= sc; x.z *= sc; x.w *= sc;
          stg_stream(o + i, x);
        }
      }
      __syncthreads();  // chunk j fully read -> its space may be overwritten by the async proxy
      if (tid == 0 && next_gcount > 0) issue_chunks(f, s, a.a, a.b, rank, next_n0, next_gcount, j, j + 1, pol_stream);
    }
    // chunks of the next group that lie beyond this group's valid range (only if this group was
    // partial, which can only be the last group) need no refill.
    if (tid == 0 && next_gcount > 0 && vchunks < f.nchunk)
      issue_chunks(f, s, a.a, a.b, rank, next_n0, next_gcount, vchunks, f.nchunk, pol_stream);
    parity ^= 1;
  }
  cluster.sync();  // nobody exits while a sibling may still address its shared memory
}

// =============================================================================================
// backward
// =============================================================================================
template <int L, int GMAX>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreadsF, 1)
    fused_bwd_kernel(const FusedBwdArgs a, const FusedCfg f) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cluster_id = blockIdx.x / kCluster, n_clusters = gridDim.x / kCluster;
  const Smem s = carve(smem_raw, f, true);
  const int tid = threadIdx.x;
  const uint64_t pol_stream = policy_evict_first();

  if (tid == 0) {
    for (int j = 0; j < f.nchunk; ++j) mbar_init(&s.bars[j], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0 && cluster_id < f.n_groups) {
    const int n0 = cluster_id * f.g;
    issue_chunks(f, s, a.go_a, a.go_b, rank, n0, min(f.g, f.n - n0), 0, f.nchunk, pol_stream);
  }
  cluster.sync();

  constexpr int kPlanesPerPass = kThreadsF / L;
  const int lane = tid % L, grp_in_pass = tid / L;
  const int hw4 = f.hw >> 2;
  const int ncol_h = f.dq;        // outputs of the dH GEMV handled by this CTA
  const int ncol_z = 2 * f.cq;    // outputs of the dZ GEMV handled by this CTA
  uint32_t parity = 0;

  for (int grp = cluster_id; grp < f.n_groups; grp += n_clusters) {
    const int n0 = grp * f.g;
    const int gcount = min(f.g, f.n - n0);
    const int vplanes = gcount * 2 * f.cq;
    const int vchunks = vplanes / f.pc;

    // ---- pass 1: <grad_out (shared), input (global, streamed once)> per plane -----------------
    for (int j = 0; j < vchunks; ++j) {
      // issue the global loads of the first planes before blocking on the barrier
      mbar_wait(&s.bars[j], parity);
      for (int pp = grp_in_pass; pp < f.pc; pp += kPlanesPerPass) {
        const int p = j * f.pc + pp;
        int g, mod, cl;
        plane_coords(f, p, g, mod, cl);
        const float4* gv = reinterpret_cast<const float4*>(s.data + (size_t)p * f.hw);
        const float4* xv = reinterpret_cast<const float4*>((mod ? a.b : a.a) +
                                                           ((size_t)(n0 + g) * f.c + (size_t)rank * f.cq + cl) * f.hw);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int i = lane;
        for (; i + 3 * L < hw4; i += 4 * L) {
          float4 x[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) x[u] = ldg_hint(xv + i + u * L, pol_stream);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 gg = gv[i + u * L];
            a0 = fmaf(gg.x, x[u].x, a0); a1 = fmaf(gg.y, x[u].y, a1);
            a2 = fmaf(gg.z, x[u].z, a2); a3 = fmaf(gg.w, x[u].w, a3);
          }
        }
        for (; i < hw4; i += L) {
          const float4 x = ldg_hint(xv + i, pol_stream), gg = gv[i];
          a0 = fmaf(gg.x, x.x, a0); a1 = fmaf(gg.y, x.y, a1); a2 = fmaf(gg.z, x.z, a2); a3 = fmaf(gg.w, x.w, a3);
        }
        const float t = group_sum<L>((a0 + a1) + (a2 + a3));
        if (lane == 0) s.psum[p] = t;
      }
    }
    __syncthreads();
    // ---- dE of my channels -> all CTAs + global ------------------------------------------------
    for (int p = tid; p < vplanes; p += kThreadsF) {
      int g, mod, cl;
      plane_coords(f, p, g, mod, cl);
      const int ch = rank * f.cq + cl;
      const float gate = __ldg((mod ? a.g_b : a.g_a) + (size_t)(n0 + g) * f.c + ch);
      const float de = s.psum[p] * a.gate_scale * gate * (1.f - gate);
      s.scale[p] = gate * a.gate_scale;
#pragma unroll
      for (int dst = 0; dst < kCluster; ++dst)
        cluster.map_shared_rank(s.vec_a, dst)[g * 2 * f.c + mod * f.c + ch] = de;
      (mod ? a.de_b : a.de_a)[(size_t)(n0 + g) * f.c + ch] = de;
    }
    for (int i = tid; i < kThreadsF * GMAX; i += kThreadsF) s.part[i] = 0.f;
    cluster.sync();
    // ---- dH for my quarter of the hidden units: dE_a Wv + dE_b Ws, masked by H > 0 -------------
    gemv_cols_partial<GMAX>(a.w_v, f.d, rank * f.dq, ncol_h, 0, f.c, s.vec_a, 2 * f.c, gcount, s.part);
    gemv_cols_partial<GMAX>(a.w_s, f.d, rank * f.dq, ncol_h, 0, f.c, s.vec_a + f.c, 2 * f.c, gcount, s.part);
    __syncthreads();
    {
      const int slices = kThreadsF / ncol_h;
      for (int o = tid; o < ncol_h * gcount; o += kThreadsF) {
        const int g = o / ncol_h, col = o - g * ncol_h;
        float v = 0.f;
        for (int sl = 0; sl < slices; ++sl) v += s.part[((size_t)sl * GMAX + g) * ncol_h + col];
        const int dd = rank * f.dq + col;
        const float hval = __ldg(a.h + (size_t)(n0 + g) * f.d + dd);
        v = hval > 0.f ? v : 0.f;
#pragma unroll
        for (int dst = 0; dst < kCluster; ++dst) cluster.map_shared_rank(s.vec_b, dst)[g * f.d + dd] = v;
        a.dh[(size_t)(n0 + g) * f.d + dd] = v;
      }
    }
    __syncthreads();
    for (int i = tid; i < kThreadsF * GMAX; i += kThreadsF) s.part[i] = 0.f;
    cluster.sync();
    // ---- dZ of my channels: dH Wsq[:, my columns] -----------------------------------------------
    {
      // columns of this CTA: [rank*cq, +cq) of the visual half and [C + rank*cq, +cq) of the skeleton half
      const int slices = kThreadsF / ncol_z;
      const int col = tid % ncol_z, sl = tid / ncol_z;
      const int gcol = (col < f.cq) ? rank * f.cq + col : f.c + rank * f.cq + (col - f.cq);
      const int span = (f.d + slices - 1) / slices;
      const int ka = sl * span, kb = min(f.d, ka + span);
      float acc[GMAX];
#pragma unroll
      for (int g = 0; g < GMAX; ++g) acc[g] = 0.f;
#pragma unroll 4
      for (int kk = ka; kk < kb; ++kk) {
        const float wv = __ldg(a.w_sq + (size_t)kk * 2 * f.c + gcol);
#pragma unroll
        for (int g = 0; g < GMAX; ++g)
          if (g < gcount) acc[g] = fmaf(s.vec_b[g * f.d + kk], wv, acc[g]);
      }
#pragma unroll
      for (int g = 0; g < GMAX; ++g) s.part[((size_t)sl * GMAX + g) * ncol_z + col] = acc[g];
      __syncthreads();
      for (int o = tid; o < ncol_z * gcount; o += kThreadsF) {
        const int g = o / ncol_z, c2 = o - g * ncol_z;
        float v = 0.f;
        for (int s2 = 0; s2 < slices; ++s2) v += s.part[((size_t)s2 * GMAX + g) * ncol_z + c2];
        const int mod = c2 >= f.cq, cl = c2 - mod * f.cq;
        s.addv[(g * 2 + mod) * f.cq + cl] = v / (float)f.hw;  // MeanBackward: grad / HW
      }
    }
    __syncthreads();
    // ---- pass 2: d_input = grad_out * scale + ds / HW, refill with the next group ---------------
    const int next = grp + n_clusters;
    const int next_n0 = next * f.g;
    const int next_gcount = next < f.n_groups ? min(f.g, f.n - next_n0) : 0;
    for (int j = 0; j < vchunks; ++j) {
      for (int pp = grp_in_pass; pp < f.pc; pp += kPlanesPerPass) {
        const int p = j * f.pc + pp;
        int g, mod, cl;
        plane_coords(f, p, g, mod, cl);
        const float sc = s.scale[p], ad = s.addv[p];
        const float4* v = reinterpret_cast<const float4*>(s.data + (size_t)p * f.hw);
        float4* o = reinterpret_cast<float4*>((mod ? a.d_b : a.d_a) +
                                              ((size_t)(n0 + g) * f.c + (size_t)rank * f.cq + cl) * f.hw);
#pragma unroll 4
        for (int i = lane; i < hw4; i += L) {
          float4 x = v[i];
          x.x = fmaf(x.x, sc, ad); x.y = fmaf(x.y, sc, ad); x.z = fmaf(x.z, sc, ad); x.w = fmaf(x.w, sc, ad);
          stg_stream(o + i, x);
        }
      }
      __syncthreads();
      if (tid == 0 && next_gcount > 0)
        issue_chunks(f, s, a.go_a, a.go_b, rank, next_n0, next_gcount, j, j + 1, pol_stream);
    }
    if (tid == 0 && next_gcount > 0 && vchunks < f.nchunk)
      issue_chunks(f, s, a.go_a, a.go_b, rank, next_n0, next_gcount, vchunks, f.nchunk, pol_stream);
    parity ^= 1;
  }
  cluster.sync();
}

// ---- host side ------------------------------------------------------------------------------
bool make_cfg(int n, int c, int hw, int d, FusedCfg* out) {
  if (n <= 0 || c % (4 * kCluster) != 0 || d % (4 * kCluster) != 0 || hw % 4 != 0) return false;
  FusedCfg f;
  f.n = n; f.c = c; f.hw = hw; f.d = d;
  f.cq = c / kCluster; f.dq = d / kCluster;
  const size_t slice = (size_t)2 * f.cq * hw * sizeof(float);  // one sample, both modalities, this CTA
  if (slice == 0 || slice > kDataBudget) return false;
  f.g = (int)(kDataBudget / slice);
  if (f.g > 2) f.g = 2;  // GMAX
  if (f.g > n) f.g = n;
  f.pl = f.g * 2 * f.cq;
  // transposed GEMVs map one thread per (slice, column): the column counts must divide the block
  if (kThreadsF % f.dq != 0 || kThreadsF % (2 * f.cq) != 0 || f.dq > kThreadsF || 2 * f.cq > kThreadsF) return false;
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
  *out = f;
  return true;
}

int lanes_for(int hw) {
  const int items = hw / 4;
  int l = 8;
  while (l < 32 && items > l * 8) l <<= 1;
  return l;
}

template <typename Kern>
int launch_cluster(Kern kern, const void* args_ptr, const FusedCfg& f, bool bwd, cudaStream_t st, int tag);

}  // namespace

bool fused_supported(int n, int c_v, int c_s, int hw_v, int hw_s, int d, int mode) {
  if (mode != GML_MODE_NORMAL) return false;
  if (c_v != c_s || hw_v != hw_s) return false;
  FusedCfg f;
  if (!make_cfg(n, c_v, hw_v, d, &f)) return false;
  // per-sample FC weights are re-read from L2 for every group: only worth it while they are
  // small next to the group's feature-map bytes (MMTM4's 512x7^2 goes the streaming way)
  const double w_bytes = 4.0 * (2.0 * c_v * d + 2.0 * c_v * d);
  const double group_bytes = 2.0 * f.g * 2.0 * c_v * hw_v * 4.0;
  if (w_bytes > 1.0 * group_bytes) return false;
  return true;
}

namespace {
template <typename Args, typename K>
int do_launch(K kern, const Args& args, const FusedCfg& f, bool bwd, cudaStream_t st, int tag) {
  const size_t smem = smem_bytes(f, bwd);
  if (smem > 232448) return GML_E_UNSUPPORTED;
  GML_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // persistent grid: as many clusters as can be co-resident (1 CTA per SM), capped by the work
  static int max_clusters_cache[2] = {0, 0};
  int& max_clusters = max_clusters_cache[bwd ? 1 : 0];
  if (max_clusters == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kNumSMs / kCluster * kCluster);
    cfg.blockDim = dim3(kThreadsF);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) != cudaSuccess || nc <= 0) {
      cudaGetLastError();
      nc = kNumSMs / kCluster - 4;  // conservative: 33 clusters of 4 fit a B200 (SURVEY / microarch notes)
    }
    max_clusters = nc;
  }
  int clusters = f.n_groups < max_clusters ? f.n_groups : max_clusters;
  if (clusters < 1) clusters = 1;
  {
    LaunchScope ls(tag, st);
    kern<<<clusters * kCluster, kThreadsF, smem, st>>>(args, f);
  }
  GML_LAUNCH_CHECK();
  return GML_OK;
}
}  // namespace

int launch_fused_fwd(const FusedFwdArgs& args, cudaStream_t st) {
  FusedCfg f;
  if (!make_cfg(args.n, args.c, args.hw, args.d, &f)) return GML_E_UNSUPPORTED;
  if (!aligned16(args.a) || !aligned16(args.b) || !aligned16(args.a_out) || !aligned16(args.b_out) ||
      !aligned16(args.w_sq) || !aligned16(args.w_v) || !aligned16(args.w_s))
    return GML_E_UNSUPPORTED;
  const int l = lanes_for(f.hw);
  if (f.g == 1) {
    if (l == 32) return do_launch(fused_fwd_kernel<32, 1>, args, f, false, st, kTagFusedFwd);
    if (l == 16) return do_launch(fused_fwd_kernel<16, 1>, args, f, false, st, kTagFusedFwd);
    return do_launch(fused_fwd_kernel<8, 1>, args, f, false, st, kTagFusedFwd);
  }
  if (l == 32) return do_launch(fused_fwd_kernel<32, 2>, args, f, false, st, kTagFusedFwd);
  if (l == 16) return do_launch(fused_fwd_kernel<16, 2>, args, f, false, st, kTagFusedFwd);
  return do_launch(fused_fwd_kernel<8, 2>, args, f, false, st, kTagFusedFwd);
}

int launch_fused_bwd(const FusedBwdArgs& args, cudaStream_t st) {
  FusedCfg f;
  if (!make_cfg(args.n, args.c, args.hw, args.d, &f)) return GML_E_UNSUPPORTED;
  if (!aligned16(args.go_a) || !aligned16(args.go_b) || !aligned16(args.a) || !aligned16(args.b) ||
      !aligned16(args.d_a) || !aligned16(args.d_b))
    return GML_E_UNSUPPORTED;
  const int l = lanes_for(f.hw);
  if (f.g == 1) {
    if (l == 32) return do_launch(fused_bwd_kernel<32, 1>, args, f, true, st, kTagFusedBwd);
    if (l == 16) return do_launch(fused_bwd_kernel<16, 1>, args, f, true, st, kTagFusedBwd);
    return do_launch(fused_bwd_kernel<8, 1>, args, f, true, st, kTagFusedBwd);
  }
  if (l == 32) return do_launch(fused_bwd_kernel<32, 2>, args, f, true, st, kTagFusedBwd);
  if (l == 16) return do_launch(fused_bwd_kernel<16, 2>, args, f, true, st, kTagFusedBwd);
  return do_launch(fused_bwd_kernel<8, 2>, args, f, true, st, kTagFusedBwd);
}

}  // namespace gml
