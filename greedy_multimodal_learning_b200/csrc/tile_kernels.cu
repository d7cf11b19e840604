// Tile pipeline: ONE persistent kernel per MMTM block and direction (normal mode).
//
// The batch is cut into tiles of M samples.  A tile goes through four stages
//     R  plane reduction (forward: sums -> squeeze;  backward: <grad_out, input> -> dE)
//     F1 first batched FC of the tile on the tensor cores  (forward: H = relu(Z Wsq^T + b); backward: dH)
//     F2 second batched FC                                  (forward: gates;                backward: dZ)
//     S  plane scaling (forward: A' = A g;  backward: dA = grad_out g + dZ / HW)
// and the stages of DIFFERENT tiles overlap inside one launch: while the tensor cores run the FCs of tile
// t, the other SMs already reduce tile t + 1 .. t + LAG and scale tile t - 1.  S re-reads what R read at most
// LAG + 1 tiles (tens of MB) earlier, so it hits L2: the feature maps cross HBM once per direction (forward 4u,
// backward 6u) although the FCs are batched over 128 samples (the weights are read once per tile, not once per
// sample as in the cluster kernels of fused_kernels.cu).
//
// CTAs are persistent, one per SM, and take one of two roles when they start (arrival order):
//   * stream CTAs: warp 8 is the loader -- it draws R / S work items (one chunk of P planes, ~25 KB, of one
//     modality) from a global ticket counter, waits for the item's dependency (S needs the tile's gates), and
//     moves the chunk global -> shared with 1-D TMA bulk copies into a ring of slots (mbarrier full/empty pairs);
//     warps 0-7 reduce or scale the chunk out of shared memory (any plane size: 49-float planes need no
//     alignment tricks) and store results / outputs with plain coalesced 128-bit stores.
//   * GEMM CTAs: draw F items (128 x 128 output tile x one K split) from a second ticket counter and run them
//     on tcgen05 with the 3xTF32 split (umma.cuh; same scheme as gemm_umma_kernel): 8 warps stage operands with
//     cp.async into the K-major SWIZZLE_128B layout and split them, warp 8 issues the MMAs into two alternating
//     TMEM accumulators (<= 128 k each), the workers drain them with round-to-nearest adds.  K splits of a tile
//     are folded by the last split to finish, in split order (bit-reproducible).  The backward's first items
//     transpose the weights once so that every GEMM is K-major x K-major.
// Stages are ordered by release/acquire counters in global memory.  Every wait points at work with a LOWER ticket,
// and a ticket is only ever held by a running CTA, so the pipeline cannot deadlock whatever the number of
// resident CTAs; all waits are bounded (trap, never hang).
//
// Replaces, for mode 0: reference src/balanced_mmtm.py:93-111,128-133,154 (forward) and the autograd graph of those
// lines (backward).
#include <mutex>

#include "common.cuh"
#include "kernels.h"
#include "umma.cuh"

namespace gml {

int g_tile_kind = 0;        // tunable "tile_kind": 0 automatic, 1 whenever the shape is supported, 2 never
int g_tile_lag = 2;         // tunable "tile_lag": tiles between a tile's R and S stage in the stream queue
int g_tile_gemm_ctas = 0;   // tunable "tile_gemm_ctas": 0 = from the FLOP/byte estimate
int g_tile_m = 0;           // tunable "tile_m": samples per tile, 0 = automatic (~25 MB of feature map per tile)
int g_tile_chunk_kb = 28;   // tunable "tile_chunk_kb": upper bound of a chunk (one work item) in KB
int g_tile_min_mb = 8;      // tunable "tile_min_mb": automatic mode takes the tile path from this many MB per modality

namespace {

constexpr int kThreads = UTHREADS;  // 288: warps 0-7 workers, warp 8 control (loader / MMA issuer)
constexpr int kMaxSlots = 8;
constexpr int kMaxNT = 16;          // 128-wide column tiles of a GEMM stage (N <= 2048)
constexpr int kMaxSplits = 8;
constexpr int kRingTiles = 4;       // split-K partial planes are recycled every kRingTiles tiles
constexpr int kCtrHead = 8;
constexpr int kCtrPerTile = 4 + 2 * kMaxNT;
constexpr unsigned long long kTimeoutNs = 4000000000ull;

enum { kItemR = 0, kItemS = 1, kItemStop = 2 };
enum { kEpiRelu = 0, kEpiSigmoid = 1, kEpiMask = 2, kEpiDiv = 3 };

struct GemmStage {
  const float* a; const float* a2;   // A rows = samples, K-major; k >= k_split comes from a2 (k - k_split)
  int lda, lda2, k_split;
  const float* b; const float* b2;   // B rows = output columns, K-major; rows >= n_split come from b2 (n - n_split)
  int ldb, n_split;
  int n_total, k_total;
  int n_tiles, splits, k_per_split;  // k_per_split is a multiple of UK
  float* out; float* out2;           // column < n_split -> out[row * ldo + col], else out2[row * ldo + col - n_split]
  int ldo;
  const float* bias; const float* bias2;
  const float* mask; int ldmask;     // kEpiMask: keep where mask > 0
  int epi; float div;
};

struct ColItem {   // out[j] = sum_i x[i * ld + j]; optional running-mean update (balanced_mmtm.py:113-114)
  const float* x; float* out; int rows, cols, ld;
  float* run_v; float* run_s; float n_total, step;
};

struct TileParams {
  const float* x[2];      // resident operand per modality: forward inputs, backward grad_out
  const float* y[2];      // backward: saved inputs (read once); forward: unused
  float* out[2];          // forward A', B'; backward dA, dB
  const float* gate[2];   // [N*C] gates per modality (forward: written by F2 in this launch)
  const float* add[2];    // backward S: dZ / HW per plane (written by F2 in this launch)
  float* rout[2];         // R output per modality; forward: z + mod * C with row stride 2C; backward: dE flat
  int rout_ld;            // row stride of rout (forward 2C, backward C)
  GemmStage st[2];
  ColItem cs[4];
  int n_cs;
  // backward prologue: transposed weights
  const float* w_v; const float* w_s; const float* w_sq;
  float* w_cat_t; float* w_sq_t;   // [D, 2C], [2C, D]; nullptr in the forward
  int n, c, hw, d, bwd;
  int m_tile, n_tiles, p, lanes, lag, n_gemm;
  int slots, nbuf;
  uint32_t chunk_bytes, slot_bytes, hw_magic;
  float gate_scale;
  unsigned* ctr;
  float* part;
  size_t part_tile_floats;   // partial planes of one tile: (sum over stages of n_tiles * splits) * 128 * 128
};

// ---- small PTX helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// bounded wait on a monotonically increasing counter written with release semantics by other CTAs
__device__ __forceinline__ void wait_counter(const unsigned* p, unsigned need) {
  if (ld_acquire_u32(p) >= need) return;
  const unsigned long long t0 = global_ns();
  unsigned ns = 32;
  while (ld_acquire_u32(p) < need) {
    __nanosleep(ns);
    if (ns < 512) ns <<= 1;
    if (global_ns() - t0 > kTimeoutNs) __trap();  // a lost signal must surface as an error, never as a hung GPU
  }
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(u_smem_addr(bar)), "r"(bytes) : "memory");
}
// 1-D TMA: global -> this CTA's shared memory, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          u_smem_addr(dst)),
      "l"(src), "r"(bytes), "r"(u_smem_addr(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ float ldcg_f32(const float* p) { return __ldcg(p); }

struct SlotMeta { int kind, mod, tile, q0; };

__device__ __forceinline__ int tile_rows(const TileParams& P, int t) {
  const int r = P.n - t * P.m_tile;
  return r < P.m_tile ? r : P.m_tile;
}
__device__ __forceinline__ unsigned* tile_ctr(const TileParams& P, int t) { return P.ctr + kCtrHead + (size_t)t * kCtrPerTile; }

// =================================================================================================================
// stream role
// =================================================================================================================
// shared memory of a stream CTA: [slots x (nbuf chunks | gate[P] | add[P])] [metas] [barriers]
struct StreamSmem {
  unsigned char* base;
  uint32_t slot_bytes, chunk_bytes;
  int p;
  __device__ __forceinline__ float* chunk(int slot, int b) const {
    return reinterpret_cast<float*>(base + (size_t)slot * slot_bytes + (size_t)b * chunk_bytes);
  }
  __device__ __forceinline__ float* gate(int slot, int nbuf) const {
    return reinterpret_cast<float*>(base + (size_t)slot * slot_bytes + (size_t)nbuf * chunk_bytes);
  }
  __device__ __forceinline__ float* add(int slot, int nbuf) const { return gate(slot, nbuf) + p; }
};

__device__ void stream_loader(const TileParams& P, const StreamSmem& sm, SlotMeta* metas, uint64_t* full, uint64_t* empty) {
  const uint64_t pol_keep = policy_evict_last(), pol_drop = policy_evict_first();
  const int T = P.n_tiles, nseg = 2 * (T + P.lag);
  const size_t planes_per_tile = (size_t)P.m_tile * P.c;
  int seg = 0;
  unsigned seg_start = 0;
  auto seg_tile = [&](int j) { return (j & 1) ? (j >> 1) - P.lag : (j >> 1); };
  auto seg_size = [&](int j) -> unsigned {
    const int t = seg_tile(j);
    if (t < 0 || t >= T) return 0u;
    return 2u * (unsigned)((size_t)tile_rows(P, t) * P.c / P.p);
  };
  unsigned uses = 0;
  int slot = 0;
  const unsigned need_f2 = (unsigned)P.st[1].n_tiles;
  // tickets are drawn two at a time and one draw ahead: the atomic's round trip overlaps the current items
  unsigned cur = atomicAdd(&P.ctr[0], 2u);
  for (;;) {
    const unsigned nxt = atomicAdd(&P.ctr[0], 2u);
    for (unsigned ticket = cur; ticket < cur + 2u; ++ticket) {
      while (seg < nseg && ticket >= seg_start + seg_size(seg)) { seg_start += seg_size(seg); ++seg; }
      if (uses >= (unsigned)P.slots) u_mbar_wait(&empty[slot], ((uses / P.slots) - 1u) & 1u);
      SlotMeta m;
      if (seg >= nseg) {  // queue exhausted: tell the workers
        m.kind = kItemStop; m.mod = 0; m.tile = 0; m.q0 = 0;
        metas[slot] = m;
        u_mbar_arrive(&full[slot]);
        return;
      }
      const int t = seg_tile(seg);
      const unsigned idx = ticket - seg_start;
      const unsigned chunks = (unsigned)((size_t)tile_rows(P, t) * P.c / P.p);
      m.kind = (seg & 1) ? kItemS : kItemR;
      m.mod = idx >= chunks ? 1 : 0;
      m.tile = t;
      m.q0 = (int)((size_t)t * planes_per_tile + (size_t)(idx - (m.mod ? chunks : 0u)) * P.p);
      if (m.kind == kItemS) {
        wait_counter(tile_ctr(P, t) + 2, need_f2);
        asm volatile("fence.proxy.async;" ::: "memory");  // the gates were written through the generic proxy
      }
      metas[slot] = m;
      const size_t off = (size_t)m.q0 * P.hw;
      const uint32_t vec_bytes = (uint32_t)P.p * 4u;
      if (m.kind == kItemR) {
        if (P.bwd) {
          mbar_expect_tx(&full[slot], 2 * P.chunk_bytes + vec_bytes);
          bulk_g2s(sm.chunk(slot, 0), P.x[m.mod] + off, P.chunk_bytes, &full[slot], pol_keep);
          bulk_g2s(sm.chunk(slot, 1), P.y[m.mod] + off, P.chunk_bytes, &full[slot], pol_drop);
          bulk_g2s(sm.gate(slot, P.nbuf), P.gate[m.mod] + m.q0, vec_bytes, &full[slot], pol_keep);
        } else {
          mbar_expect_tx(&full[slot], P.chunk_bytes);
          bulk_g2s(sm.chunk(slot, 0), P.x[m.mod] + off, P.chunk_bytes, &full[slot], pol_keep);
        }
      } else {
        mbar_expect_tx(&full[slot], P.chunk_bytes + vec_bytes * (P.bwd ? 2u : 1u));
        bulk_g2s(sm.chunk(slot, 0), P.x[m.mod] + off, P.chunk_bytes, &full[slot], pol_drop);
        bulk_g2s(sm.gate(slot, P.nbuf), P.gate[m.mod] + m.q0, vec_bytes, &full[slot], pol_drop);
        if (P.bwd) bulk_g2s(sm.add(slot, P.nbuf), P.add[m.mod] + m.q0, vec_bytes, &full[slot], pol_drop);
      }
      ++uses;
      if (++slot == P.slots) slot = 0;
    }
    cur = nxt;
  }
}

// plane sums (forward) or <grad_out, input> dots (backward) of the P planes of a chunk
__device__ __forceinline__ void reduce_chunk(const TileParams& P, const SlotMeta& m, const float* b0, const float* b1,
                                             const float* sgate, int tid) {
  const int L = P.lanes, lane_in = tid & (L - 1), grp = tid / L, ngrp = U_PRODUCERS / L;
  const int hw = P.hw;
  // first plane of the chunk: sample n, channel c0 (a chunk never straddles samples: P divides C)
  const int n = m.q0 / P.c, c0 = m.q0 - n * P.c;
  float* dst = P.rout[m.mod] + (size_t)n * P.rout_ld + c0;
  // the trip count is uniform over the CTA (shuffles below): inactive groups run with `act == false`
  for (int base = 0; base < P.p; base += ngrp) {
    const int pl = base + grp;
    const bool act = pl < P.p;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (act) {
      if ((hw & 3) == 0) {
        const int hw4 = hw >> 2;
        const float4* v = reinterpret_cast<const float4*>(b0) + (size_t)pl * hw4;
        if (P.bwd) {
          const float4* w = reinterpret_cast<const float4*>(b1) + (size_t)pl * hw4;
#pragma unroll 4
          for (int i = lane_in; i < hw4; i += L) {
            const float4 g = v[i], x = w[i];
            a0 = fmaf(g.x, x.x, a0); a1 = fmaf(g.y, x.y, a1); a2 = fmaf(g.z, x.z, a2); a3 = fmaf(g.w, x.w, a3);
          }
        } else {
#pragma unroll 4
          for (int i = lane_in; i < hw4; i += L) {
            const float4 x = v[i];
            a0 += x.x; a1 += x.y; a2 += x.z; a3 += x.w;
          }
        }
      } else {
        const float* v = b0 + (size_t)pl * hw;
        const float* w = b1 + (size_t)pl * hw;
        int i = lane_in;
        if (P.bwd) {
          for (; i + 3 * L < hw; i += 4 * L) {
            a0 = fmaf(v[i], w[i], a0); a1 = fmaf(v[i + L], w[i + L], a1);
            a2 = fmaf(v[i + 2 * L], w[i + 2 * L], a2); a3 = fmaf(v[i + 3 * L], w[i + 3 * L], a3);
          }
          for (; i < hw; i += L) a0 = fmaf(v[i], w[i], a0);
        } else {
          for (; i + 3 * L < hw; i += 4 * L) { a0 += v[i]; a1 += v[i + L]; a2 += v[i + 2 * L]; a3 += v[i + 3 * L]; }
          for (; i < hw; i += L) a0 += v[i];
        }
      }
    }
    float t = (a0 + a1) + (a2 + a3);
    for (int o = L >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (act && lane_in == 0) {
      if (P.bwd) {
        const float g = sgate[pl];
        dst[pl] = t * P.gate_scale * g * (1.f - g);   // dE = dg * g (1 - g)
      } else {
        dst[pl] = t / (float)hw;                       // squeeze
      }
    }
  }
}

// plane of element e of a chunk: e / HW by multiplication (exact for e < 2^20, 2 <= HW < 2^12; HW = 1 has no 32-bit magic)
__device__ __forceinline__ int plane_of(unsigned e, const TileParams& P) {
  return P.hw == 1 ? (int)e : (int)__umulhi(e, P.hw_magic);
}

// out = x * (gate * gate_scale) [+ add]: flat 128-bit walk over the chunk, the plane of an element from its index
__device__ __forceinline__ void scale_chunk(const TileParams& P, const SlotMeta& m, const float* b0, const float* sgate,
                                            const float* sadd, int tid) {
  const int hw = P.hw;
  const int nvec = (int)(P.chunk_bytes >> 4);
  const float4* v = reinterpret_cast<const float4*>(b0);
  float4* o = reinterpret_cast<float4*>(P.out[m.mod] + (size_t)m.q0 * hw);
  const float gs = P.gate_scale;
  if ((hw & 3) == 0) {
#pragma unroll 2
    for (int i = tid; i < nvec; i += U_PRODUCERS) {
      const int pl = plane_of((unsigned)(4 * i), P);
      const float sc = sgate[pl] * gs;
      float4 x = v[i];
      if (P.bwd) {
        const float ad = sadd[pl];
        x.x = fmaf(x.x, sc, ad); x.y = fmaf(x.y, sc, ad); x.z = fmaf(x.z, sc, ad); x.w = fmaf(x.w, sc, ad);
      } else {
        x.x *= sc; x.y *= sc; x.z *= sc; x.w *= sc;
      }
      stg_stream(o + i, x);
    }
  } else {
#pragma unroll 2
    for (int i = tid; i < nvec; i += U_PRODUCERS) {
      const unsigned e = 4u * (unsigned)i;
      const int p0 = plane_of(e, P), p1 = plane_of(e + 1, P), p2 = plane_of(e + 2, P), p3 = plane_of(e + 3, P);
      float4 x = v[i];
      if (P.bwd) {
        x.x = fmaf(x.x, sgate[p0] * gs, sadd[p0]); x.y = fmaf(x.y, sgate[p1] * gs, sadd[p1]);
        x.z = fmaf(x.z, sgate[p2] * gs, sadd[p2]); x.w = fmaf(x.w, sgate[p3] * gs, sadd[p3]);
      } else {
        x.x *= sgate[p0] * gs; x.y *= sgate[p1] * gs; x.z *= sgate[p2] * gs; x.w *= sgate[p3] * gs;
      }
      stg_stream(o + i, x);
    }
  }
}

__device__ void stream_workers(const TileParams& P, const StreamSmem& sm, const SlotMeta* metas, uint64_t* full,
                               uint64_t* empty, int tid) {
  const int lane = tid & 31;
  int slot = 0;
  uint32_t phase = 0;
  for (;;) {
    u_mbar_wait(&full[slot], phase);
    const SlotMeta m = metas[slot];
    if (m.kind == kItemStop) return;
    if (m.kind == kItemR) {
      reduce_chunk(P, m, sm.chunk(slot, 0), sm.chunk(slot, 1), sm.gate(slot, P.nbuf), tid);
      __syncwarp();
      if (lane == 0) {
        u_mbar_arrive(&empty[slot]);
        red_release_add(tile_ctr(P, m.tile) + 0, 1u);  // 8 arrivals (one per worker warp) per R item
      }
    } else {
      scale_chunk(P, m, sm.chunk(slot, 0), sm.gate(slot, P.nbuf), sm.add(slot, P.nbuf), tid);
      __syncwarp();
      if (lane == 0) u_mbar_arrive(&empty[slot]);
    }
    if (++slot == P.slots) { slot = 0; phase ^= 1u; }
  }
}

// =================================================================================================================
// GEMM role
// =================================================================================================================
struct GemmBars { uint64_t* full; uint64_t* empty; uint64_t* acc_full; uint64_t* acc_empty; };

struct GemmItem { int stage, tile, ntile, split; };

// K-major operand tile (128 rows x 32 k) via cp.async, rows resolved by `rowptr` (nullptr -> zero fill)
template <typename RowPtr>
__device__ __forceinline__ void t_load_tile(unsigned char* tile, RowPtr rowptr, const float* dummy, int t0, int k0,
                                            int kmax, int warp, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int t, kc;
    u_chunk<true>(i, warp, lane, t, kc);
    const uint32_t land = u_kmajor_off(t, kc);
    const float* row = rowptr(t0 + t);
    const int k = k0 + kc * 4;
    int bytes = row ? (kmax - k) * 4 : 0;
    bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
    cp_async16(reinterpret_cast<float*>(tile + land), bytes > 0 ? row + k : dummy, bytes);
  }
}

// One 128 x 128 x [k0, k1) product on tcgen05 (3xTF32).  Workers leave with the fp32 result of their row x 64
// columns in `sum`.  kbase / gbase: k-tiles and accumulator groups this CTA has pushed through its barriers so far.
__device__ __forceinline__ void gemm_item_mainloop(const TileParams& P, const GemmStage& g, const GemmItem& it,
                                                   unsigned char* u_smem, const GemmBars& bars, uint32_t tmem,
                                                   uint32_t kbase, uint32_t gbase, int nk, float (&sum)[64], int warp,
                                                   int lane) {
  const int m0 = it.tile * P.m_tile, m_end = m0 + tile_rows(P, it.tile);
  const int n0 = it.ntile * UN;
  const int k_begin = it.split * g.k_per_split;
  const int k_end = min(g.k_total, k_begin + g.k_per_split);
  if (warp == U_PRODUCERS / 32) {
    // ===== MMA issuer ==========================================================================================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(UN >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);
      for (int kt = 0; kt < nk; ++kt) {
        const uint32_t KT = kbase + (uint32_t)kt, GI = gbase + (uint32_t)(kt / UGROUP);
        const uint32_t stage = KT % USTAGES, b = GI & 1u;
        if (kt % UGROUP == 0 && GI >= 2) u_mbar_wait(&bars.acc_empty[b], ((GI >> 1) - 1u) & 1u);
        u_mbar_wait(&bars.full[stage], (KT / USTAGES) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_big = u_smem_addr(u_smem + stage * U_STAGE_BYTES), a_small = a_big + U_TILE_BYTES;
        const uint32_t b_big = a_big + 2 * U_TILE_BYTES, b_small = a_big + 3 * U_TILE_BYTES;
        const uint32_t acc = tmem + b * (uint32_t)UN;
#pragma unroll
        for (int j = 0; j < UK / 8; ++j) {
          const uint32_t o = (uint32_t)j * 32u;
          const uint64_t da_b = u_desc(a_big + o), da_s = u_desc(a_small + o);
          const uint64_t db_b = u_desc(b_big + o), db_s = u_desc(b_small + o);
          u_mma_tf32(acc, da_s, db_b, idesc, (kt % UGROUP != 0 || j != 0) ? 1u : 0u);
          u_mma_tf32(acc, da_b, db_s, idesc, 1u);
          u_mma_tf32(acc, da_b, db_b, idesc, 1u);
        }
        u_commit(&bars.empty[stage]);
        if (kt % UGROUP == UGROUP - 1 || kt == nk - 1) u_commit(&bars.acc_full[b]);
      }
    }
    __syncwarp();
    return;
  }
  // ===== producers / drainers ==================================================================================
#pragma unroll
  for (int i = 0; i < 64; ++i) sum[i] = 0.f;
  const int ngroups = (nk + UGROUP - 1) / UGROUP;
  const uint32_t t_row = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * 64);
  int drained = 0;
  auto drain = [&](int gi) {
    const uint32_t GI = gbase + (uint32_t)gi, b = GI & 1u;
    u_mbar_wait(&bars.acc_full[b], (GI >> 1) & 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float v[16];
      u_tmem_ld16(t_row + b * (uint32_t)UN + (uint32_t)(16 * c), v);
#pragma unroll
      for (int i = 0; i < 16; ++i) sum[16 * c + i] += v[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    u_mbar_arrive(&bars.acc_empty[b]);
  };
  auto issue_loads = [&](int kt) {
    const uint32_t KT = kbase + (uint32_t)kt;
    // the stage was last read by the MMAs of k-tile KT - USTAGES
    if (KT >= (uint32_t)USTAGES) u_mbar_wait(&bars.empty[KT % USTAGES], ((KT / USTAGES) - 1u) & 1u);
    unsigned char* st = u_smem + (KT % USTAGES) * U_STAGE_BYTES;
    const int k0 = k_begin + kt * UK;
    // A: second K segment from a2
    const bool second = g.k_split && k0 >= g.k_split;
    const float* abase = second ? g.a2 : g.a;
    const int lda = second ? g.lda2 : g.lda;
    const int ka = second ? k0 - g.k_split : k0;
    const int kamax = second ? k_end - g.k_split : (g.k_split ? min(k_end, g.k_split) : k_end);
    t_load_tile(st, [&](int r) { return r < m_end ? abase + (size_t)r * lda : (const float*)nullptr; }, g.a, m0, ka,
                kamax, warp, lane);
    t_load_tile(st + 2 * U_TILE_BYTES,
                [&](int r) {
                  return r < g.n_split ? g.b + (size_t)r * g.ldb
                                       : (r < g.n_total ? g.b2 + (size_t)(r - g.n_split) * g.ldb : (const float*)nullptr);
                },
                g.a, n0, k0, k_end, warp, lane);
  };
#pragma unroll
  for (int s = 0; s < USTAGES - 1; ++s) {
    if (s < nk) issue_loads(s);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    const int nxt = kt + USTAGES - 1;
    if (nxt < nk) issue_loads(nxt);
    cp_async_commit();
    cp_async_wait<USTAGES - 1>();  // my chunks of k-tile kt have landed
    const uint32_t KT = kbase + (uint32_t)kt;
    unsigned char* st = u_smem + (KT % USTAGES) * U_STAGE_BYTES;
    float4 xa[4], xb[4];
    u_read_chunks<true>(st, warp, lane, xa);
    u_read_chunks<true>(st + 2 * U_TILE_BYTES, warp, lane, xb);
    u_write_split<true>(st, st + U_TILE_BYTES, warp, lane, xa);
    u_write_split<true>(st + 2 * U_TILE_BYTES, st + 3 * U_TILE_BYTES, warp, lane, xb);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA unit
    u_mbar_arrive(&bars.full[KT % USTAGES]);
    if (kt % UGROUP >= 1 && drained < kt / UGROUP) { drain(drained); ++drained; }
  }
  cp_async_wait<0>();
  while (drained < ngroups) { drain(drained); ++drained; }
}

// epilogue + store of the 64 columns a worker holds
__device__ __forceinline__ void gemm_store(const TileParams& P, const GemmStage& g, const GemmItem& it,
                                           const float (&sum)[64], int warp, int lane) {
  const int m0 = it.tile * P.m_tile, m_end = m0 + tile_rows(P, it.tile);
  const int row = m0 + 32 * (warp & 3) + lane;
  if (row >= m_end) return;
  const int nb = it.ntile * UN + (warp >> 2) * 64;
  const bool vec = (g.n_split & 3) == 0 && (g.ldo & 3) == 0 && (g.n_total & 3) == 0;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const int n = nb + 4 * c;
    if (n >= g.n_total) break;
    float r[4] = {sum[4 * c], sum[4 * c + 1], sum[4 * c + 2], sum[4 * c + 3]};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int col = n + e;
      if (col >= g.n_total) continue;
      const bool hi = col >= g.n_split;
      float v = r[e];
      if (g.bias) v += hi ? g.bias2[col - g.n_split] : g.bias[col];
      if (g.epi == kEpiRelu) v = fmaxf(v, 0.f);
      else if (g.epi == kEpiSigmoid) v = sigmoidf_ref(v);
      else if (g.epi == kEpiMask) v = __ldg(g.mask + (size_t)row * g.ldmask + col) > 0.f ? v : 0.f;
      else v = v / g.div;
      r[e] = v;
    }
    const bool hi = n >= g.n_split;
    float* dst = (hi ? g.out2 + (size_t)row * g.ldo + (n - g.n_split) : g.out + (size_t)row * g.ldo + n);
    if (vec) {
      *reinterpret_cast<float4*>(dst) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = n + e;
        if (col >= g.n_total) continue;
        if (col >= g.n_split) g.out2[(size_t)row * g.ldo + (col - g.n_split)] = r[e];
        else g.out[(size_t)row * g.ldo + col] = r[e];
      }
    }
  }
}

// 32 x 32 transposes by the worker warps of the GEMM CTAs (backward prologue):
//   w_cat_t[d, k] = k < C ? w_v[k, d] : w_s[k - C, d]   ([D, 2C]);   w_sq_t[j, d] = w_sq[d, j]   ([2C, D])
__device__ void transpose_weights(const TileParams& P, int grank, float* scratch, int warp, int lane) {
  const int tc = P.c / 32, td = P.d / 32;
  const int per = tc * td;           // tiles of one [C, D] matrix
  const int total = 4 * per;         // w_v, w_s, w_sq (2C columns = two halves)
  float (*s)[33] = reinterpret_cast<float (*)[33]>(scratch + (size_t)warp * 32 * 33);
  for (int tile = grank * 8 + warp; tile < total; tile += P.n_gemm * 8) {
    const int mat = tile / per, r = tile - mat * per;
    const float* in; float* out; int ld_in, ld_out, r0, c0, out_col_off;
    if (mat < 2) {          // in [C, D] -> out [D, 2C] at column offset mat * C
      in = mat ? P.w_s : P.w_v; ld_in = P.d; out = P.w_cat_t; ld_out = 2 * P.c; out_col_off = mat * P.c;
      r0 = (r / td) * 32; c0 = (r % td) * 32;
    } else {                // in [D, 2C] (column half mat - 2) -> out [2C, D]
      in = P.w_sq + (mat - 2) * P.c; ld_in = 2 * P.c; out = P.w_sq_t + (size_t)(mat - 2) * P.c * P.d; ld_out = P.d;
      out_col_off = 0;
      r0 = (r / tc) * 32; c0 = (r % tc) * 32;
    }
#pragma unroll 8
    for (int i = 0; i < 32; ++i) s[i][lane] = __ldg(in + (size_t)(r0 + i) * ld_in + c0 + lane);
    __syncwarp();
#pragma unroll 8
    for (int i = 0; i < 32; ++i) out[(size_t)(c0 + i) * ld_out + out_col_off + r0 + lane] = s[lane][i];
    __syncwarp();
  }
}

// column sums over the batch (gate sum + running mean; bias gradients): 32 columns per item, 8 row groups
__device__ void colsum_item(const ColItem& ci, int colblock, float* scratch, int tid) {
  float (*part)[33] = reinterpret_cast<float (*)[33]>(scratch);
  const int tx = tid & 31, ty = tid >> 5;
  const int j = colblock * 32 + tx;
  float acc = 0.f;
  if (tid < U_PRODUCERS && j < ci.cols) {
    const float* col = ci.x + j;
    int i = ty;
    for (; i + 15 * 8 < ci.rows; i += 16 * 8) {
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = ldcg_f32(col + (size_t)(i + 8 * u) * ci.ld);
#pragma unroll
      for (int u = 0; u < 16; ++u) acc += v[u];
    }
    for (; i < ci.rows; i += 8) acc += ldcg_f32(col + (size_t)i * ci.ld);
  }
  if (tid < U_PRODUCERS) part[ty][tx] = acc;
  __syncthreads();
  if (tid < 32 && j < ci.cols) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += part[r][tx];
    ci.out[j] = t;
    if (ci.run_v) {
      const float mean = t / ci.n_total;
      ci.run_v[j] = (mean + ci.run_v[j] * ci.step) / (ci.step + 1.f);
      if (ci.run_s) ci.run_s[j] = (mean + ci.run_s[j] * ci.step) / (ci.step + 1.f);
    }
  }
  __syncthreads();
}

__device__ void gemm_role(const TileParams& P, int grank, unsigned char* u_smem, int tid) {
  __shared__ __align__(8) uint64_t bar_full[USTAGES], bar_empty[USTAGES], bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t s_tmem;
  __shared__ unsigned s_ticket, s_flag;
  const int warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < USTAGES; ++s) { u_mbar_init(&bar_full[s], U_PRODUCERS); u_mbar_init(&bar_empty[s], 1); }
#pragma unroll
    for (int b = 0; b < 2; ++b) { u_mbar_init(&bar_acc_full[b], 1); u_mbar_init(&bar_acc_empty[b], U_PRODUCERS); }
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
  const GemmBars bars{bar_full, bar_empty, bar_acc_full, bar_acc_empty};

  if (P.w_cat_t) {  // backward: K-major copies of the weights for the dH / dZ products
    if (warp < 8) transpose_weights(P, grank, reinterpret_cast<float*>(u_smem), warp, lane);
    __threadfence();
    __syncthreads();
    if (tid == 0) red_release_add(&P.ctr[3], 1u);
  }

  const int T = P.n_tiles;
  const int i1 = P.st[0].n_tiles * P.st[0].splits, i2 = P.st[1].n_tiles * P.st[1].splits;
  const unsigned per_tile = (unsigned)(i1 + i2);
  int cs_blocks[5];
  cs_blocks[0] = 0;
  for (int i = 0; i < 4; ++i) cs_blocks[i + 1] = cs_blocks[i] + (i < P.n_cs ? (P.cs[i].cols + 31) / 32 : 0);
  const unsigned total = per_tile * (unsigned)T + (unsigned)cs_blocks[4];
  uint32_t kbase = 0, gbase = 0;
  float sum[64];

  for (;;) {
    if (tid == 0) s_ticket = atomicAdd(&P.ctr[1], 1u);
    __syncthreads();
    const unsigned ticket = s_ticket;
    if (ticket >= total) break;
    if (ticket >= per_tile * (unsigned)T) {
      // ---- column-sum item: needs every tile's R, F1 and F2 -----------------------------------------------------
      if (tid == 0) {
        for (int t = 0; t < T; ++t) {
          const unsigned rneed = 16u * (unsigned)((size_t)tile_rows(P, t) * P.c / P.p);
          wait_counter(tile_ctr(P, t) + 0, rneed);
          wait_counter(tile_ctr(P, t) + 1, (unsigned)P.st[0].n_tiles);
          wait_counter(tile_ctr(P, t) + 2, (unsigned)P.st[1].n_tiles);
        }
      }
      __syncthreads();
      int blk = (int)(ticket - per_tile * (unsigned)T), which = 0;
      while (which < 3 && blk >= cs_blocks[which + 1]) ++which;
      colsum_item(P.cs[which], blk - cs_blocks[which], reinterpret_cast<float*>(u_smem), tid);
      continue;
    }
    GemmItem it;
    it.tile = (int)(ticket / per_tile);
    const int r = (int)(ticket - (unsigned)it.tile * per_tile);
    it.stage = r >= i1 ? 1 : 0;
    const GemmStage& g = P.st[it.stage];
    const int rr = it.stage ? r - i1 : r;
    it.ntile = rr / g.splits;
    it.split = rr - it.ntile * g.splits;
    unsigned* tc = tile_ctr(P, it.tile);
    if (tid == 0) {
      if (P.w_cat_t) wait_counter(&P.ctr[3], (unsigned)P.n_gemm);
      if (it.stage == 0) {
        const unsigned rneed = 16u * (unsigned)((size_t)tile_rows(P, it.tile) * P.c / P.p);  // 2 mods x 8 warps
        wait_counter(tc + 0, rneed);
        // the partial planes of this ring position were last used by tile - kRingTiles: fully folded?
        if (it.tile >= kRingTiles) wait_counter(tile_ctr(P, it.tile - kRingTiles) + 2, (unsigned)P.st[1].n_tiles);
      } else {
        wait_counter(tc + 1, (unsigned)P.st[0].n_tiles);
      }
    }
    __syncthreads();
    const int k_begin = it.split * g.k_per_split;
    const int k_end = min(g.k_total, k_begin + g.k_per_split);
    const int nk = k_end > k_begin ? (k_end - k_begin + UK - 1) / UK : 0;
    gemm_item_mainloop(P, g, it, u_smem, bars, tmem, kbase, gbase, nk, sum, warp, lane);
    kbase += (uint32_t)nk;
    gbase += (uint32_t)((nk + UGROUP - 1) / UGROUP);

    bool finish = true;
    if (g.splits > 1) {
      // publish my partial (thread-major layout: perfectly coalesced, and the folding CTA uses the same mapping)
      float* part = P.part + (size_t)(it.tile % kRingTiles) * P.part_tile_floats +
                    ((size_t)(it.stage ? i1 : 0) + (size_t)it.ntile * g.splits) * (UM * UN);
      if (warp < 8) {
        float4* mine = reinterpret_cast<float4*>(part + (size_t)it.split * (UM * UN)) + tid;
#pragma unroll
        for (int c = 0; c < 16; ++c)
          __stcg(mine + c * U_PRODUCERS, make_float4(sum[4 * c], sum[4 * c + 1], sum[4 * c + 2], sum[4 * c + 3]));
      }
      __threadfence();
      __syncthreads();
      if (tid == 0) s_flag = atomicAdd(tc + 4 + it.stage * kMaxNT + it.ntile, 1u);
      __syncthreads();
      finish = s_flag == (unsigned)g.splits - 1u;
      if (finish && warp < 8) {
        __threadfence();
        // fold in split order (fixed order -> bit-reproducible whichever CTA arrives last)
#pragma unroll
        for (int i = 0; i < 64; ++i) sum[i] = 0.f;
        for (int sp = 0; sp < g.splits; ++sp) {
          const float4* src = reinterpret_cast<const float4*>(part + (size_t)sp * (UM * UN)) + tid;
          float4 v[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) v[c] = __ldcg(src + c * U_PRODUCERS);
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            sum[4 * c] += v[c].x; sum[4 * c + 1] += v[c].y; sum[4 * c + 2] += v[c].z; sum[4 * c + 3] += v[c].w;
          }
        }
      }
    }
    if (finish) {
      if (warp < 8) gemm_store(P, g, it, sum, warp, lane);
      __threadfence();
      __syncthreads();
      if (tid == 0) red_release_add(tc + 1 + it.stage, 1u);
    }
  }
  // every warp is done with its tcgen05.ld before the columns go back to the allocator
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(U_TMEM_COLS) : "memory");
  }
}

__global__ void __launch_bounds__(kThreads, 1) tile_pipeline_kernel(const __grid_constant__ TileParams P) {
  extern __shared__ __align__(1024) unsigned char t_smem_raw[];
  unsigned char* smem = t_smem_raw + ((1024u - (u_smem_addr(t_smem_raw) & 1023u)) & 1023u);
  __shared__ unsigned s_role;
  __shared__ __align__(8) uint64_t s_full[kMaxSlots], s_empty[kMaxSlots];
  __shared__ SlotMeta s_meta[kMaxSlots];
  const int tid = threadIdx.x;
  if (tid == 0) s_role = atomicAdd(&P.ctr[2], 1u);
  __syncthreads();
  const int role = (int)s_role;
  if (role < P.n_gemm) {
    gemm_role(P, role, smem, tid);
    return;
  }
  if (tid == 0) {
    for (int s = 0; s < P.slots; ++s) { u_mbar_init(&s_full[s], 1); u_mbar_init(&s_empty[s], U_PRODUCERS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  StreamSmem sm{smem, P.slot_bytes, P.chunk_bytes, P.p};
  if (tid >= U_PRODUCERS) {
    if (tid == U_PRODUCERS) stream_loader(P, sm, s_meta, s_full, s_empty);
    __syncwarp();
  } else {
    stream_workers(P, sm, s_meta, s_full, s_empty, tid);
  }
}

// ---- host side -------------------------------------------------------------------------------------------------
struct TileCfg {
  int m_tile, n_tiles, p, lanes, n_gemm, slots, nbuf, lag;
  uint32_t chunk_bytes, slot_bytes;
  int nt[2], splits[2], kps[2];
  size_t part_tile_floats, ctr_bytes, part_bytes, smem_bytes;
};

int sm_count() {
  static int cached[16] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return kNumSMs;
  if (!cached[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMs;
    cached[dev] = n;
  }
  return cached[dev];
}

bool make_tile_cfg(int n, int c, int hw, int d, bool bwd, TileCfg* o) {
  if (n <= 0 || c % 32 != 0 || d % 32 != 0 || hw <= 0 || hw >= 4096) return false;
  if (2 * c > kMaxNT * UN || d > kMaxNT * UN) return false;
  if ((long long)n * c >= (1LL << 31) / 2) return false;
  TileCfg f;
  // chunk: P planes, P | C, P % 4 == 0 (16-byte aligned gate vectors and chunk bytes), as large as the bound allows
  const size_t bound = (size_t)g_tile_chunk_kb * 1024;
  f.p = 0;
  for (int p = c; p >= 4; --p) {
    if (c % p || p % 4) continue;
    if ((size_t)p * hw * 4 <= bound) { f.p = p; break; }
  }
  if (!f.p) return false;
  f.chunk_bytes = (uint32_t)((size_t)f.p * hw * 4);
  int lanes = 1;
  while (lanes < 32 && lanes * 2 * f.p <= U_PRODUCERS) lanes *= 2;
  f.lanes = lanes;
  f.nbuf = bwd ? 2 : 1;
  f.slot_bytes = (uint32_t)round_up((size_t)f.nbuf * f.chunk_bytes + 2 * (size_t)f.p * 4, 128);
  const size_t gemm_smem = (size_t)USTAGES * U_STAGE_BYTES;
  const size_t budget = 224 * 1024 - 1024;
  int slots = (int)(budget / f.slot_bytes);
  if (slots > kMaxSlots) slots = kMaxSlots;
  if (slots < 2) return false;
  f.slots = slots;
  size_t sm = (size_t)slots * f.slot_bytes;
  if (sm < gemm_smem) sm = gemm_smem;
  f.smem_bytes = sm + 1024;
  // tile: ~25 MB of resident feature map (both modalities), at least ~4 tiles for the pipeline, <= 128 samples
  const size_t per_sample = (size_t)2 * c * hw * 4;
  long m = g_tile_m > 0 ? g_tile_m : (long)((25u << 20) / per_sample);
  if (m > UM) m = UM;
  if (g_tile_m <= 0 && m > 8) m &= ~7L;
  if (g_tile_m <= 0 && m > (n + 3) / 4) m = (n + 3) / 4;
  if (m < 1) m = 1;
  f.m_tile = (int)m;
  f.n_tiles = (n + f.m_tile - 1) / f.m_tile;
  f.lag = g_tile_lag < 1 ? 1 : g_tile_lag;
  // GEMM stages: forward (K = 2C -> N = D), (K = D -> N = 2C); backward (K = 2C -> N = D), (K = D -> N = 2C)
  const int kk[2] = {2 * c, d}, nn[2] = {d, 2 * c};
  size_t items = 0, units = 0;
  for (int s = 0; s < 2; ++s) {
    f.nt[s] = ceil_div(nn[s], UN);
    const int nk = ceil_div(kk[s], UK);
    int sp = ceil_div(nk, 8);
    if (sp > kMaxSplits) sp = kMaxSplits;
    f.kps[s] = ceil_div(nk, sp) * UK;
    f.splits[s] = ceil_div(kk[s], f.kps[s]);
    items += (size_t)f.nt[s] * f.splits[s];
    units += (size_t)f.nt[s] * nk;
  }
  f.part_tile_floats = items * UM * UN;
  f.part_bytes = (size_t)kRingTiles * f.part_tile_floats * 4;
  f.ctr_bytes = round_up((size_t)(kCtrHead + (size_t)f.n_tiles * kCtrPerTile) * 4, 256);
  // GEMM CTAs: k-tile units of one tile x ~0.6 us each, against the tile's streaming time at ~6.5 TB/s
  int g = g_tile_gemm_ctas;
  if (g <= 0) {
    const double step_us = (double)f.m_tile * per_sample * (bwd ? 3.0 : 2.0) / 6.5e6;
    g = (int)(units * 0.6 / (step_us > 0.1 ? step_us : 0.1)) + 2;
    if (g > 40) g = 40;
  }
  const int sms = sm_count();
  if (g > sms / 3) g = sms / 3;
  if (g < 1) g = 1;
  f.n_gemm = g;
  *o = f;
  return true;
}

void fill_common(TileParams& P, const TileCfg& f, int n, int c, int hw, int d, bool bwd, void* ws) {
  P.n = n; P.c = c; P.hw = hw; P.d = d; P.bwd = bwd ? 1 : 0;
  P.m_tile = f.m_tile; P.n_tiles = f.n_tiles; P.p = f.p; P.lanes = f.lanes; P.lag = f.lag; P.n_gemm = f.n_gemm;
  P.slots = f.slots; P.nbuf = f.nbuf; P.chunk_bytes = f.chunk_bytes; P.slot_bytes = f.slot_bytes;
  P.hw_magic = (uint32_t)(((1ull << 32) + (unsigned)hw - 1) / (unsigned)hw);
  P.ctr = reinterpret_cast<unsigned*>(ws);
  P.part = reinterpret_cast<float*>(static_cast<char*>(ws) + f.ctr_bytes);
  P.part_tile_floats = f.part_tile_floats;
  for (int s = 0; s < 2; ++s) {
    P.st[s].n_tiles = f.nt[s]; P.st[s].splits = f.splits[s]; P.st[s].k_per_split = f.kps[s];
    P.st[s].a2 = nullptr; P.st[s].lda2 = 0; P.st[s].k_split = 0; P.st[s].b2 = nullptr;
    P.st[s].out2 = nullptr; P.st[s].bias = nullptr; P.st[s].bias2 = nullptr; P.st[s].mask = nullptr; P.st[s].ldmask = 0;
    P.st[s].div = 1.f;
  }
  P.n_cs = 0;
  P.w_cat_t = nullptr; P.w_sq_t = nullptr; P.w_v = P.w_s = P.w_sq = nullptr;
}

int launch_pipeline(const TileParams& P, const TileCfg& f, int tag, cudaStream_t st) {
  GML_CUDA_TRY(cudaMemsetAsync(P.ctr, 0, f.ctr_bytes, st));
  static std::mutex mu;
  static bool attr_set[16] = {false};
  int dev = 0;
  GML_CUDA_TRY(cudaGetDevice(&dev));
  {
    std::lock_guard<std::mutex> lk(mu);
    if (dev < 0 || dev >= 16 || !attr_set[dev]) {
      cudaFuncAttributes fa;
      GML_CUDA_TRY(cudaFuncGetAttributes(&fa, tile_pipeline_kernel));
      GML_CUDA_TRY(cudaFuncSetAttribute(tile_pipeline_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        227 * 1024 - (int)fa.sharedSizeBytes));
      if (dev >= 0 && dev < 16) attr_set[dev] = true;
    }
  }
  int grid = sm_count();
  // no more stream CTAs than there are items in flight for them (tiny problems)
  {
    LaunchScope ls(tag, st);
    tile_pipeline_kernel<<<grid, kThreads, f.smem_bytes, st>>>(P);
  }
  GML_LAUNCH_CHECK();
  return GML_OK;
}

}  // namespace

bool tile_supported(int n, int c_v, int c_s, int hw_v, int hw_s, int d, int mode) {
  if (g_tile_kind == 2) return false;
  if (mode != GML_MODE_NORMAL || c_v != c_s || hw_v != hw_s) return false;
  TileCfg f;
  return make_tile_cfg(n, c_v, hw_v, d, false, &f) && make_tile_cfg(n, c_v, hw_v, d, true, &f);
}

bool tile_preferred(int n, int c, int hw, int d) {
  (void)d;
  if (g_tile_kind == 1) return true;
  return (size_t)n * c * hw * 4 >= ((size_t)g_tile_min_mb << 20);
}

size_t tile_fwd_workspace_bytes(int n, int c, int hw, int d) {
  TileCfg f;
  if (!make_tile_cfg(n, c, hw, d, false, &f)) return 0;
  return f.ctr_bytes + f.part_bytes + 256;
}

size_t tile_bwd_workspace_bytes(int n, int c, int hw, int d) {
  TileCfg f;
  if (!make_tile_cfg(n, c, hw, d, true, &f)) return 0;
  return f.ctr_bytes + f.part_bytes + 2 * round_up((size_t)2 * c * d * 4, 256) + 256;
}

int launch_tile_fwd(const FusedFwdArgs& a, float* gate_sum, float* run_v, float* run_s, float step, void* ws,
                    size_t ws_bytes, cudaStream_t st) {
  TileCfg f;
  if (!make_tile_cfg(a.n, a.c, a.hw, a.d, false, &f)) return GML_E_UNSUPPORTED;
  if (!ws || ws_bytes < f.ctr_bytes + f.part_bytes) return GML_E_WORKSPACE;
  const void* al[] = {a.a, a.b, a.a_out, a.b_out, a.w_sq, a.w_v, a.w_s, a.z, a.h, a.g_a, a.g_b, ws};
  for (const void* p : al)
    if (!aligned16(p)) return GML_E_UNSUPPORTED;
  TileParams P;
  fill_common(P, f, a.n, a.c, a.hw, a.d, false, ws);
  P.x[0] = a.a; P.x[1] = a.b; P.y[0] = P.y[1] = nullptr;
  P.out[0] = a.a_out; P.out[1] = a.b_out;
  P.gate[0] = a.g_a; P.gate[1] = a.g_b;
  P.add[0] = P.add[1] = nullptr;
  P.rout[0] = a.z; P.rout[1] = a.z + a.c; P.rout_ld = 2 * a.c;
  P.gate_scale = a.gate_scale;
  GemmStage& s0 = P.st[0];   // H = relu(Z Wsq^T + bsq)
  s0.a = a.z; s0.lda = 2 * a.c; s0.b = a.w_sq; s0.ldb = 2 * a.c; s0.n_split = a.d; s0.n_total = a.d; s0.k_total = 2 * a.c;
  s0.out = a.h; s0.ldo = a.d; s0.bias = a.b_sq; s0.epi = kEpiRelu;
  GemmStage& s1 = P.st[1];   // [g_a | g_b] = sigmoid(H [Wv ; Ws]^T + [bv | bs])
  s1.a = a.h; s1.lda = a.d; s1.b = a.w_v; s1.b2 = a.w_s; s1.ldb = a.d; s1.n_split = a.c; s1.n_total = 2 * a.c;
  s1.k_total = a.d; s1.out = a.g_a; s1.out2 = a.g_b; s1.ldo = a.c; s1.bias = a.b_v; s1.bias2 = a.b_s; s1.epi = kEpiSigmoid;
  if (gate_sum) {
    P.cs[0] = ColItem{a.g_a, gate_sum, a.n, a.c, a.c, run_v, run_s, (float)a.n, step};
    P.n_cs = 1;
  }
  return launch_pipeline(P, f, kTagFusedFwd, st);
}

int launch_tile_bwd(const FusedBwdArgs& a, float* dz_flat, float* d_b_v, float* d_b_s, float* d_b_sq, void* ws,
                    size_t ws_bytes, cudaStream_t st) {
  TileCfg f;
  if (!make_tile_cfg(a.n, a.c, a.hw, a.d, true, &f)) return GML_E_UNSUPPORTED;
  const size_t wt = round_up((size_t)2 * a.c * a.d * 4, 256);
  if (!ws || ws_bytes < f.ctr_bytes + f.part_bytes + 2 * wt) return GML_E_WORKSPACE;
  const void* al[] = {a.go_a, a.go_b, a.a, a.b, a.d_a, a.d_b, a.w_sq, a.w_v, a.w_s, a.h, a.g_a, a.g_b, a.de_a, a.de_b,
                      a.dh, dz_flat, ws};
  for (const void* p : al)
    if (!aligned16(p)) return GML_E_UNSUPPORTED;
  TileParams P;
  fill_common(P, f, a.n, a.c, a.hw, a.d, true, ws);
  char* wp = static_cast<char*>(ws) + f.ctr_bytes + f.part_bytes;
  P.w_cat_t = reinterpret_cast<float*>(wp);
  P.w_sq_t = reinterpret_cast<float*>(wp + wt);
  P.w_v = a.w_v; P.w_s = a.w_s; P.w_sq = a.w_sq;
  float* dz_a = dz_flat;
  float* dz_b = dz_flat + (size_t)a.n * a.c;
  P.x[0] = a.go_a; P.x[1] = a.go_b; P.y[0] = a.a; P.y[1] = a.b;
  P.out[0] = a.d_a; P.out[1] = a.d_b;
  P.gate[0] = a.g_a; P.gate[1] = a.g_b;
  P.add[0] = dz_a; P.add[1] = dz_b;
  P.rout[0] = a.de_a; P.rout[1] = a.de_b; P.rout_ld = a.c;
  P.gate_scale = a.gate_scale;
  GemmStage& s0 = P.st[0];   // dH = ([dE_a | dE_b] [Wv ; Ws]) * [H > 0]
  s0.a = a.de_a; s0.a2 = a.de_b; s0.lda = a.c; s0.lda2 = a.c; s0.k_split = a.c;
  s0.b = P.w_cat_t; s0.ldb = 2 * a.c; s0.n_split = a.d; s0.n_total = a.d; s0.k_total = 2 * a.c;
  s0.out = a.dh; s0.ldo = a.d; s0.mask = a.h; s0.ldmask = a.d; s0.epi = kEpiMask;
  GemmStage& s1 = P.st[1];   // dZ = dH Wsq, stored per modality and already divided by HW (MeanBackward)
  // (w_sq_t is ONE [2C, D] matrix: b2 simply continues it at row C, where the output switches to dz_b)
  s1.a = a.dh; s1.lda = a.d; s1.b = P.w_sq_t; s1.b2 = P.w_sq_t + (size_t)a.c * a.d; s1.ldb = a.d; s1.n_split = a.c;
  s1.n_total = 2 * a.c; s1.k_total = a.d; s1.out = dz_a; s1.out2 = dz_b; s1.ldo = a.c; s1.epi = kEpiDiv;
  s1.div = (float)a.hw;
  int ncs = 0;
  if (d_b_v) P.cs[ncs++] = ColItem{a.de_a, d_b_v, a.n, a.c, a.c, nullptr, nullptr, 1.f, 0.f};
  if (d_b_s) P.cs[ncs++] = ColItem{a.de_b, d_b_s, a.n, a.c, a.c, nullptr, nullptr, 1.f, 0.f};
  if (d_b_sq) P.cs[ncs++] = ColItem{a.dh, d_b_sq, a.n, a.d, a.d, nullptr, nullptr, 1.f, 0.f};
  P.n_cs = ncs;
  return launch_pipeline(P, f, kTagFusedBwd, st);
}

}  // namespace gml
