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
#include <cuda.h>

#include <mutex>

#include "common.cuh"
#include "kernels.h"
#include "umma.cuh"

namespace gml {

int g_tile_kind = 0;        // tunable "tile_kind": 0 automatic, 1 whenever the shape is supported, 2 never
int g_tile_lag = 0;         // tunable "tile_lag": tiles between a tile's R and S stage in the stream queue (0 = automatic)
int g_tile_gemm_ctas = 0;   // tunable "tile_gemm_ctas": 0 = from the FLOP/byte estimate
int g_tile_m = 0;           // tunable "tile_m": samples per tile, 0 = automatic (~25 MB of feature map per tile)
int g_tile_chunk_kb = 28;   // tunable "tile_chunk_kb": upper bound of a backward chunk (one work item) in KB; a slot holds two
int g_tile_chunk_kb_fwd = 56;  // tunable "tile_chunk_kb_fwd": the same for the forward (one buffer per slot).  The loader warp
                            // spends ~1000 cycles per item whatever its size, so fewer, larger items stream faster:
                            // forward 128x28^2 0.434 -> 0.379 ms, 256x14^2 0.236 -> 0.213 ms (profiles/r2_sweep.md)
int g_tile_light_fwd = 0;    // tunable "tile_light_fwd": 1 = forward of blocks with < 1 MB of weights on the pipeline from tile_min_mb_light on
int g_tile_min_mb_light = 190;  // tunable "tile_min_mb_light": the same threshold for the forward of blocks with < 1 MB of weights
int g_tile_min_mb = 96;     // tunable "tile_min_mb": automatic mode takes the tile path from this many MB per modality (forward: x0.5 for >= 2 MB of FC weights, x1.5 below)
int g_tile_ksplit_tiles = 0;  // tunable "tile_ksplit_tiles": k-tiles per split-K item (0 = no split-K: the deep pipeline hides the
                            // chain latency, and partial planes + folds cost more GEMM-CTA time than they save)
int g_tile_switch = 1;      // tunable "tile_switch": GEMM CTAs join the stream role once the GEMM tickets are exhausted
int g_tile_trace_only = 0;  // tunable "tile_trace_only" (with tile_stats_ptr): timeline rows only
int g_tile_rpol = 0;        // tunable "tile_rpol": L2 policy of the R-stage reads (0 auto: forward evict_last, backward evict_first; 1 evict_first, 2 evict_normal, 3 evict_last)
int g_tile_max_slots = 0;   // tunable "tile_max_slots": cap on the slot ring (0 = as many as fit)
int g_tile_split_copies = 1;
int g_tile_draw = 4;        // tunable "tile_draw": stream tickets per draw (two draws are kept in flight)
int g_tile_wgrad = 1;       // tunable "tile_wgrad": weight-gradient GEMMs run inside the backward pipeline launch
int g_tile_nodeps = 0;      // debug/measurement ONLY (wrong results): S items do not wait for the gates
long long* g_tile_stats = nullptr;  // debug: per-CTA cycle breakdown (device buffer, 16 slots per CTA)

namespace {

constexpr int kThreads = U_PRODUCERS + 64;  // 320: warps 0-7 workers, warp 8 loader / MMA issuer, warp 9 TMA producer (GEMM role)
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
  // operands as TMA tensor maps (fp32 [rows, K], box 32 k x 32 rows, SWIZZLE_128B = the canonical K-major UMMA layout)
  alignas(64) CUtensorMap tm_a;    // A rows = samples
  alignas(64) CUtensorMap tm_a2;   // k >= k_split comes from here (k - k_split)
  alignas(64) CUtensorMap tm_b;    // B rows = output columns
  alignas(64) CUtensorMap tm_b2;   // rows >= n_split come from here (n - n_split)
  int k_split, n_split;
  int b_box_rows;                  // 128 when a 128-row B tile never straddles n_split (one TMA per tile), else 32
  int n_total, k_total;
  int n_tiles, splits, k_per_split;  // k_per_split is a multiple of UK
  float* out; float* out2;           // column < n_split -> out[row * ldo + col], else out2[row * ldo + col - n_split]
  int ldo;
  const float* bias; const float* bias2;
  const float* mask; int ldmask;     // kEpiMask: keep where mask > 0
  int epi; float div;
  // weight-gradient ("post") stages: A is a [m_rows, K] matrix of its own (K = batch), tiled in 128 rows; output rows
  // >= m_split go to out2 (row - m_split).  0 = a stage of the per-tile FC chain (A rows = the samples of the tile).
  int m_rows, m_split;
};

// transpose job: out[j * ld_out + i] = in[i * ld_in + j], i < rows, j < cols (32 x 32 tiles through shared memory)
struct TrJob { const float* in; float* out; int rows, cols, ld_in, ld_out; };

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
  GemmStage st[4];          // [0], [1]: the FC chain of a tile; [2], [3]: weight gradients (backward, optional)
  int n_post;               // 0 or 2 post stages
  // K-major operands of the weight-gradient GEMMs (K = batch): H^T and Z^T are transposed by the otherwise idle warp 9
  // of the stream CTAs at the start of the launch; dE^T is written by the R-stage workers next to dE, dH^T by the
  // epilogue of the dH GEMM next to dH.  ldt = row stride of all four ([*, ldt], ldt = N rounded up to 32).
  TrJob tr_pre[2];
  int n_tr_pre;
  float* de_t; float* dh_t; int ldt;
  ColItem cs[4];
  int n_cs;
  // backward prologue: transposed weights
  const float* w_v; const float* w_s; const float* w_sq;
  float* w_cat_t; float* w_sq_t;   // [D, 2C], [2C, D]; nullptr in the forward
  int n, c, hw, d, bwd;
  int m_tile, n_tiles, p, lanes, lanes_log2, lag, n_gemm;
  int slots_log2, cps;       // log2(slots); chunks per (sample, modality) = C / P
  uint32_t cps_magic;        // ceil(2^32 / cps): ch / cps == __umulhi(ch, cps_magic) for ch < 2^16
  int slots, wps, nbuf;      // slots = 2 * groups; a group of wps = 8 / groups worker warps owns two slots (one being
                             // processed while the other one loads)
  uint32_t chunk_bytes, slot_bytes, hw_magic;
  float gate_scale;
  unsigned* ctr;
  float* part;
  size_t part_tile_floats;   // partial planes of one tile: (sum over stages of n_tiles * splits) * 128 * 128
  int gemm_joins_stream;
  int draw;                  // stream tickets per draw
  int split_copies;          // measurement: forward R chunks are fetched as this many bulk copies
  int rpol;                  // L2 policy of the R-stage reads: 0 evict_last, 1 evict_first, 2 evict_normal
  int trace_only;            // with `stats`: only the cheap timeline rows, no per-CTA cycle accounting
  int nodeps;                // measurement only: S items skip their dependency wait (results are garbage)
  long long* stats;          // debug (tunable "tile_stats_ptr"): 16 clock64 sums per CTA, or nullptr
};
// Measurement hooks (cycle breakdowns, the timeline trace, the stage-isolation modes of scripts/sweep.py) are compiled in
// only with -DGML_TILE_TRACE (make EXTRA=-DGML_TILE_TRACE): the loader warp runs ~1000 dependent cycles per item, every
// instruction in its loop counts.
#ifdef GML_TILE_TRACE
#define TILE_STATS(P) ((P).stats)
#define TILE_NODEPS(P) ((P).nodeps)
#define TILE_TRACE_ONLY(P) ((P).trace_only)
#else
#define TILE_STATS(P) (static_cast<long long*>(nullptr))
#define TILE_NODEPS(P) 0
#define TILE_TRACE_ONLY(P) 0
#endif


// ---- small PTX helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// cycle counter the compiler may not move memory operations across (debug statistics)
__device__ __forceinline__ long long clk() {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
  return t;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// bounded wait on a monotonically increasing counter written with release semantics by other CTAs
// timeline trace (tunable tile_stats_ptr, measurement only): absolute globaltimer ns
constexpr int kTraceGemmRow = 400, kTraceTileRow = 700, kTraceStartRow = 399;
__device__ __forceinline__ void trace_min(long long* stats, int row, int col) {
  if (stats) atomicMin(reinterpret_cast<unsigned long long*>(stats) + (size_t)row * 16 + col, global_ns());
}
__device__ __forceinline__ void trace_max(long long* stats, int row, int col) {
  if (stats) atomicMax(reinterpret_cast<unsigned long long*>(stats) + (size_t)row * 16 + col, global_ns());
}

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

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {  // non-blocking
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(u_smem_addr(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

struct __align__(16) SlotMeta { int kind, mod, tile, q0, rofs, sample, c0, pad2; };  // rofs: offset of the chunk's first R result

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

// (all 32 lanes of a converged warp call these; one elected lane issues -- see umma.cuh for why)
__device__ __forceinline__ void bulk_g2s_elect(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n\t}"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_el(uint32_t bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}"
      ::"r"(bar), "r"(bytes)
      : "memory");
}

// The loader: the WHOLE control warp runs it converged (every value below is the same in all lanes); instructions with
// side effects are issued by one lane.  It has to sustain one ~25 KB item per few hundred cycles, so per item there is
// no integer division (segment data is recomputed only when the ticket crosses into a new segment; sample / channel of
// a chunk by multiplication), the S dependency is polled once per tile, and the TMA operands never sit in a divergent
// region.
// The loader also signals R-stage completion: a slot's `empty` barrier tells it that all worker warps of the slot are
// done with the item that occupied it, so it adds this CTA's finished R items of a tile to the tile's global counter
// with ONE release per (CTA, tile) instead of a gpu-scope fence per warp and item.
__device__ void stream_loader(const TileParams& P, const StreamSmem& sm, SlotMeta* metas, uint64_t* full, uint64_t* empty,
                              long long* st_cycles, int lane) {
  const uint64_t pol_drop = policy_evict_first();
  // R-stage reads: the forward asks L2 to keep them (a late S item of the same tile sometimes still hits); the backward's
  // S stage never does, and evict_last lines only push the GEMM operands (Z, dE, weights) out: 512x7^2 backward
  // 0.222 -> 0.208 ms with evict_first (profiles/r2_sweep.md).  rpol: 0 = this rule, 1 evict_first, 2 normal, 3 evict_last
  const uint64_t pol_keep = P.rpol == 0 ? (P.bwd ? pol_drop : policy_evict_last())
                                        : (P.rpol == 1 ? pol_drop : (P.rpol == 2 ? policy_evict_normal() : policy_evict_last()));
  const int T = P.n_tiles, nseg = 2 * (T + P.lag);
  const uint32_t smem0 = u_smem_addr(sm.base), full0 = u_smem_addr(full);
  const uint32_t vec_bytes = (uint32_t)P.p * 4u;
  // current segment (R or S stage of one tile): [seg_start, seg_end) in ticket space
  int seg = -1, seg_t = 0, seg_kind = kItemR;
  unsigned seg_start = 0, seg_end = 0, seg_chunks = 0;
  int seg_qbase = 0, seg_n0 = 0;
  auto next_segment = [&]() {
    ++seg;
    seg_start = seg_end;
    if (seg >= nseg) { seg_end = 0xffffffffu; return; }
    seg_kind = (seg & 1) ? kItemS : kItemR;
    seg_t = (seg & 1) ? (seg >> 1) - P.lag : (seg >> 1);
    if (seg_t < 0 || seg_t >= T) { seg_chunks = 0; return; }
    // measurement only: 5 = R stage alone, 6 = S stage alone (no GEMM work, no dependencies; results are garbage)
    if (((TILE_NODEPS(P) == 5 || TILE_NODEPS(P) == 7) && seg_kind == kItemS) || ((TILE_NODEPS(P) == 6 || TILE_NODEPS(P) == 8) && seg_kind == kItemR)) { seg_chunks = 0; return; }
    seg_chunks = (unsigned)(tile_rows(P, seg_t) * P.cps);
    seg_end = seg_start + 2u * seg_chunks;
    seg_n0 = seg_t * P.m_tile;
    seg_qbase = seg_n0 * P.c;
  };
  next_segment();
  unsigned uses = 0;       // items issued so far; item u lives in slot u % slots
  unsigned retired = 0;    // items [0, retired) are known to be finished by their worker warps
  int slot = 0;
  const unsigned need_f2 = (unsigned)P.st[1].n_tiles;
  int ready_tile = -1;     // highest tile whose gates are known to be complete
  // R bookkeeping, all in registers (every lane holds the same values).  `iss` / `fin` pack four 16-bit counters: R items
  // issued / finished for tiles flush_tile .. flush_tile + 3 (`flush_tile` = lowest tile whose count this CTA has not
  // published yet); `rec` packs one byte per slot: 0x80 | (tile & 0x7f) for an R item, 0 for an S item.
  unsigned long long iss = 0, fin = 0, rec = 0;
  int flush_tile = 0;
  long long c_empty = 0, c_dep = 0, c_head = 0, c_issue = 0, c_draw = 0, t_prev = 0;
  const bool timing = st_cycles != nullptr;
  if (timing) t_prev = clk();

  auto account = [&](int s_) {   // the item in slot s_ is finished
    const unsigned r = (unsigned)(rec >> (8 * s_)) & 0xffu;
    if (r & 0x80u) {
      const unsigned j = ((r & 0x7fu) - ((unsigned)flush_tile & 0x7fu)) & 0x7fu;  // tile - flush_tile (< 4 by construction)
      fin += 1ull << (16 * j);
    }
  };
  // items [retired, upto) -> wait for their workers, account finished R items
  auto retire = [&](unsigned upto) {
    while (retired < upto) {
      const int s_ = (int)(retired & (unsigned)(P.slots - 1));
      u_mbar_wait(&empty[s_], (retired >> P.slots_log2) & 1u);
      account(s_);
      ++retired;
    }
  };
  // the same without blocking: whatever the workers have finished by now
  auto retire_ready = [&]() {
    while (retired < uses) {
      const int s_ = (int)(retired & (unsigned)(P.slots - 1));
      if (!mbar_test(&empty[s_], (retired >> P.slots_log2) & 1u)) break;
      account(s_);
      ++retired;
    }
  };
  // publish every tile whose R items (of this CTA) are all issued and finished
  auto flush = [&](bool exhausted) {
    while (flush_tile < T && (exhausted || seg > 2 * flush_tile) && (fin & 0xffffull) == (iss & 0xffffull)) {
      const unsigned cnt = (unsigned)(iss & 0xffffull);
      if (cnt && lane == 0) red_release_add(tile_ctr(P, flush_tile) + 0, cnt);
      if (cnt && lane == 0 && flush_tile < 64) trace_max(TILE_STATS(P), kTraceTileRow + flush_tile, 1);
      iss >>= 16; fin >>= 16;
      ++flush_tile;
    }
    __syncwarp();
  };

  // Tickets are drawn kDraw at a time and TWO draws ahead.  The atomic's round trip is 2-4 us while the memory system
  // is saturated; with four tickets drawn one batch ahead the loader spent ~0.9 us per item waiting for its next
  // batch, which capped a CTA at one 28 KB chunk per 0.9 us whatever the chunk size (profiles/r2_tile_loader.md).
  const unsigned kDraw = (unsigned)P.draw;
  auto run_batch = [&](unsigned cur) -> bool {   // true when the queue is exhausted
    for (unsigned ticket = cur; ticket < cur + kDraw; ++ticket) {
      while (ticket >= seg_end) next_segment();
      if (seg >= nseg) {  // queue exhausted: finish the bookkeeping, then tell the workers
        retire(uses);
        flush(true);
        if (lane == 0) {
          for (int s_ = 0; s_ < P.slots; ++s_) {  // every slot group waits on its own barrier
            metas[s_].kind = kItemStop;
            u_mbar_arrive(&full[s_]);
          }
          if (timing) {
            st_cycles[2] = c_empty; st_cycles[3] = c_dep; st_cycles[1] = uses;
            st_cycles[8] = c_head; st_cycles[9] = c_issue; st_cycles[10] = c_draw;
          }
        }
        __syncwarp();
        return true;
      }
      const long long t0 = timing ? clk() : 0;
      if (uses >= (unsigned)P.slots) retire(uses - (unsigned)P.slots + 1u);  // frees this item's slot
      if (seg > 2 * flush_tile) {  // a finished R stage is waiting to be published (else nothing to do: skip the polls)
        retire_ready();
        flush(false);
      }
      if (seg_kind == kItemR && seg_t - flush_tile >= 4) {  // keep the four-tile bookkeeping window valid
        retire(uses);
        flush(false);
      }
      const long long t1 = timing ? clk() : 0;
      const unsigned idx = ticket - seg_start;
      const int mod = idx >= seg_chunks ? 1 : 0;
      const unsigned ch = idx - (mod ? seg_chunks : 0u);
      const int q0 = seg_qbase + (int)ch * P.p;
      if (seg_kind == kItemS && seg_t > ready_tile) {
        if (!TILE_NODEPS(P)) {
          bool ok = ld_acquire_u32(tile_ctr(P, seg_t) + 2) >= need_f2;
          ok = __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;
          if (!ok) {
            // about to block: everything this CTA still owes the R counters must be published first
            retire(uses);
            flush(false);
            wait_counter(tile_ctr(P, seg_t) + 2, need_f2);
          }
        }
        ready_tile = seg_t;
        if (lane == 0 && seg_t < 64) trace_max(TILE_STATS(P), kTraceTileRow + seg_t, 2);
        asm volatile("fence.proxy.async;" ::: "memory");  // the gates were written through the generic proxy
      }
      const long long t2 = timing ? clk() : 0;
      c_empty += t1 - t0; c_dep += t2 - t1;
      if (timing) c_head += t0 - t_prev;
      if (lane == 0) {
        // first plane of the chunk = sample n, channel c0 (a chunk never straddles samples: P divides C)
        const int nl = P.cps == 1 ? (int)ch : (int)__umulhi(ch, P.cps_magic);  // ch / cps
        SlotMeta m;
        m.kind = seg_kind; m.mod = mod; m.tile = seg_t; m.q0 = q0;
        m.sample = seg_n0 + nl; m.c0 = ((int)ch - nl * P.cps) * P.p;
        m.rofs = m.sample * P.rout_ld + m.c0;
        *reinterpret_cast<int4*>(&metas[slot]) = make_int4(m.kind, m.mod, m.tile, m.q0);
        *reinterpret_cast<int4*>(&metas[slot].rofs) = make_int4(m.rofs, m.sample, m.c0, 0);
      }
      __syncwarp();
      rec = (rec & ~(0xffull << (8 * slot))) |
            ((unsigned long long)(seg_kind == kItemR ? (0x80u | ((unsigned)seg_t & 0x7fu)) : 0u) << (8 * slot));
      if (seg_kind == kItemR) iss += 1ull << (16 * (seg_t - flush_tile));
      const size_t off = (size_t)q0 * P.hw;
      const uint32_t sbase = smem0 + (uint32_t)slot * P.slot_bytes, bar = full0 + (uint32_t)slot * 8u;
      const uint32_t sgate = sbase + (uint32_t)P.nbuf * P.chunk_bytes;
      if (seg_kind == kItemR) {
        if (P.bwd) {
          mbar_expect_tx_el(bar, 2 * P.chunk_bytes + vec_bytes);
          bulk_g2s_elect(sbase, P.x[mod] + off, P.chunk_bytes, bar, pol_keep);
          bulk_g2s_elect(sbase + P.chunk_bytes, P.y[mod] + off, P.chunk_bytes, bar, pol_drop);
          bulk_g2s_elect(sgate, P.gate[mod] + q0, vec_bytes, bar, pol_keep);
        } else {
          mbar_expect_tx_el(bar, P.chunk_bytes);
          if (P.split_copies > 1) {  // measurement: the same bytes as several smaller bulk copies
            const uint32_t part = P.chunk_bytes / (uint32_t)P.split_copies;
            for (int j = 0; j < P.split_copies; ++j)
              bulk_g2s_elect(sbase + j * part, reinterpret_cast<const char*>(P.x[mod] + off) + (size_t)j * part, part, bar, pol_keep);
          } else {
            bulk_g2s_elect(sbase, P.x[mod] + off, P.chunk_bytes, bar, pol_keep);
          }
        }
      } else {
        mbar_expect_tx_el(bar, P.chunk_bytes + vec_bytes * (P.bwd ? 2u : 1u));
        bulk_g2s_elect(sbase, P.x[mod] + off, P.chunk_bytes, bar, pol_drop);
        bulk_g2s_elect(sgate, P.gate[mod] + q0, vec_bytes, bar, pol_drop);
        if (P.bwd) bulk_g2s_elect(sgate + vec_bytes, P.add[mod] + q0, vec_bytes, bar, pol_drop);
      }
      if (lane == 0 && TILE_STATS(P) && seg_t < 64 && idx == 0) {  // first chunk of a segment only: keeps the trace cheap
        if (seg_kind == kItemR) trace_min(TILE_STATS(P), kTraceTileRow + seg_t, 0);
        else trace_min(TILE_STATS(P), kTraceTileRow + seg_t, 3);
      }
      if (lane == 0 && TILE_STATS(P) && seg_t < 64 && seg_kind == kItemS && ticket + 1 == seg_end)
        trace_max(TILE_STATS(P), kTraceTileRow + seg_t, 4);
      ++uses;
      if (++slot == P.slots) slot = 0;
      if (timing) { t_prev = clk(); c_issue += t_prev - t2; }
    }
    return false;
  };
  unsigned pend_a = 0, pend_b = 0;
  if (lane == 0) {
    pend_a = atomicAdd(&P.ctr[0], kDraw);
    pend_b = atomicAdd(&P.ctr[0], kDraw);
  }
  unsigned cur = __shfl_sync(0xffffffffu, pend_a, 0);
  for (;;) {
    // (two named registers alternate so that no instruction touches a draw before the batch it is needed for)
    if (lane == 0) pend_a = atomicAdd(&P.ctr[0], kDraw);
    if (run_batch(cur)) return;
    { const long long d0 = timing ? clk() : 0; cur = __shfl_sync(0xffffffffu, pend_b, 0); if (timing) { const long long d1 = clk(); c_draw += d1 - d0; t_prev = d1; } }
    if (lane == 0) pend_b = atomicAdd(&P.ctr[0], kDraw);
    if (run_batch(cur)) return;
    { const long long d0 = timing ? clk() : 0; cur = __shfl_sync(0xffffffffu, pend_a, 0); if (timing) { const long long d1 = clk(); c_draw += d1 - d0; t_prev = d1; } }
  }
}

// The eight worker warps form groups of `wps` warps; a group owns TWO slots and works on one of them while the TMA
// fills the other.  Different groups work on different items at the same time, so whatever one group waits for (its
// result stores before the releasing arrive, shared-memory latency) is hidden behind the others.  Within a warp the
// loops are written for instruction-level parallelism -- 8 independent 128-bit shared loads per lane before the first
// add, two planes' shuffle chains interleaved -- because one SM only has these eight warps to hide latency with.
//
// plane sums (forward) or <grad_out, input> dots (backward) of the P planes of a chunk; L lanes per plane
template <bool BWD>
__device__ __forceinline__ float plane_partial_vec(const float4* __restrict__ v, const float4* __restrict__ w, int hw4,
                                                   int lane_in, int L) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int i0 = lane_in; i0 < hw4; i0 += 8 * L) {
    float4 x[8], y[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * L;
      const bool ok = i < hw4;
      x[u] = ok ? v[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      if (BWD) y[u] = ok ? w[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (BWD) {
        a0 = fmaf(x[u].x, y[u].x, a0); a1 = fmaf(x[u].y, y[u].y, a1);
        a2 = fmaf(x[u].z, y[u].z, a2); a3 = fmaf(x[u].w, y[u].w, a3);
      } else {
        a0 += x[u].x; a1 += x[u].y; a2 += x[u].z; a3 += x[u].w;
      }
    }
  }
  return (a0 + a1) + (a2 + a3);
}
template <bool BWD>
__device__ __forceinline__ float plane_partial_scalar(const float* __restrict__ v, const float* __restrict__ w, int hw,
                                                      int lane_in, int L) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int i = lane_in;
  for (; i + 7 * L < hw; i += 8 * L) {
    float x[8], y[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { x[u] = v[i + u * L]; if (BWD) y[u] = w[i + u * L]; }
    if (BWD) {
      a0 = fmaf(x[0], y[0], a0); a1 = fmaf(x[1], y[1], a1); a2 = fmaf(x[2], y[2], a2); a3 = fmaf(x[3], y[3], a3);
      a0 = fmaf(x[4], y[4], a0); a1 = fmaf(x[5], y[5], a1); a2 = fmaf(x[6], y[6], a2); a3 = fmaf(x[7], y[7], a3);
    } else {
      a0 += x[0]; a1 += x[1]; a2 += x[2]; a3 += x[3]; a0 += x[4]; a1 += x[5]; a2 += x[6]; a3 += x[7];
    }
  }
  for (; i < hw; i += L) a0 = BWD ? fmaf(v[i], w[i], a0) : a0 + v[i];
  return (a0 + a1) + (a2 + a3);
}

template <bool BWD>
__device__ __forceinline__ void reduce_chunk_t(const TileParams& P, const SlotMeta& m, const float* b0, const float* b1,
                                               const float* sgate, int sw, int lane) {
  const int L = P.lanes, lane_in = lane & (L - 1), gpw = 32 >> P.lanes_log2, grp = lane >> P.lanes_log2;
  const int hw = P.hw;
  float* dst = P.rout[m.mod] + m.rofs;
  const int stride = P.wps * gpw;
  // two planes per iteration (independent load / shuffle chains); the trip count is uniform over the warp
  for (int base = sw * gpw; base < P.p; base += 2 * stride) {
    const int pl0 = base + grp, pl1 = pl0 + stride;
    const bool act0 = pl0 < P.p, act1 = pl1 < P.p;
    float t0 = 0.f, t1 = 0.f;
    if ((hw & 3) == 0) {
      const int hw4 = hw >> 2;
      const float4* v = reinterpret_cast<const float4*>(b0);
      const float4* w = reinterpret_cast<const float4*>(b1);
      if (act0) t0 = plane_partial_vec<BWD>(v + (size_t)pl0 * hw4, w + (size_t)pl0 * hw4, hw4, lane_in, L);
      if (act1) t1 = plane_partial_vec<BWD>(v + (size_t)pl1 * hw4, w + (size_t)pl1 * hw4, hw4, lane_in, L);
    } else {
      if (act0) t0 = plane_partial_scalar<BWD>(b0 + (size_t)pl0 * hw, b1 + (size_t)pl0 * hw, hw, lane_in, L);
      if (act1) t1 = plane_partial_scalar<BWD>(b0 + (size_t)pl1 * hw, b1 + (size_t)pl1 * hw, hw, lane_in, L);
    }
    for (int o = L >> 1; o > 0; o >>= 1) {
      t0 += __shfl_xor_sync(0xffffffffu, t0, o);
      t1 += __shfl_xor_sync(0xffffffffu, t1, o);
    }
    if (lane_in == 0 && (TILE_NODEPS(P) != 4 || t0 == 123.456f)) {
      if (BWD) {
        // dE = dg * g (1 - g); with the weight gradients folded in, also dE^T[channel][sample] (one 4-byte store per
        // plane: ~N*2C stores per launch, absorbed by L2)
        float* tdst = P.de_t ? P.de_t + (size_t)(m.mod * P.c + m.c0) * P.ldt + m.sample : nullptr;
        if (act0) {
          const float g = sgate[pl0]; const float v = t0 * P.gate_scale * g * (1.f - g);
          dst[pl0] = v;
          if (tdst) tdst[(size_t)pl0 * P.ldt] = v;
        }
        if (act1) {
          const float g = sgate[pl1]; const float v = t1 * P.gate_scale * g * (1.f - g);
          dst[pl1] = v;
          if (tdst) tdst[(size_t)pl1 * P.ldt] = v;
        }
      } else {
        if (act0) dst[pl0] = t0 / (float)hw;   // squeeze
        if (act1) dst[pl1] = t1 / (float)hw;
      }
    }
  }
}

// plane of element e of a chunk: e / HW by multiplication (exact for e < 2^20, 2 <= HW < 2^12; HW = 1 has no 32-bit magic)
__device__ __forceinline__ int plane_of(unsigned e, const TileParams& P) {
  return P.hw == 1 ? (int)e : (int)__umulhi(e, P.hw_magic);
}

// out = x * (gate * gate_scale) [+ add]: flat 128-bit walk over the chunk, the plane of an element from its index;
// four independent vectors per lane in flight
template <bool BWD>
__device__ __forceinline__ void scale_chunk_t(const TileParams& P, const SlotMeta& m, const float* b0, const float* sgate,
                                              const float* sadd, int sw, int lane) {
  const int hw = P.hw;
  const int nvec = (int)(P.chunk_bytes >> 4);
  const int step = 32 * P.wps;
  const float4* v = reinterpret_cast<const float4*>(b0);
  float4* o = reinterpret_cast<float4*>(P.out[m.mod] + (size_t)m.q0 * hw);
  const float gs = P.gate_scale;
  const uint64_t pol_out = policy_evict_first();
  const bool vec_planes = (hw & 3) == 0;
  for (int i0 = sw * 32 + lane; i0 < nvec; i0 += 4 * step) {
    float4 x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * step;
      x[u] = i < nvec ? v[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * step;
      if (i >= nvec) break;
      const unsigned e = 4u * (unsigned)i;
      const int p0 = plane_of(e, P);
      int p1 = p0, p2 = p0, p3 = p0;
      if (!vec_planes) { p1 = plane_of(e + 1, P); p2 = plane_of(e + 2, P); p3 = plane_of(e + 3, P); }
      if (BWD) {
        x[u].x = fmaf(x[u].x, sgate[p0] * gs, sadd[p0]); x[u].y = fmaf(x[u].y, sgate[p1] * gs, sadd[p1]);
        x[u].z = fmaf(x[u].z, sgate[p2] * gs, sadd[p2]); x[u].w = fmaf(x[u].w, sgate[p3] * gs, sadd[p3]);
      } else {
        x[u].x *= sgate[p0] * gs; x[u].y *= sgate[p1] * gs; x[u].z *= sgate[p2] * gs; x[u].w *= sgate[p3] * gs;
      }
      if (TILE_NODEPS(P) != 4) stg_hint(o + i, x[u], pol_out);
      else if (x[u].x == 123.456f) stg_stream(o + i, x[u]);
    }
  }
}

__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t* bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(u_smem_addr(bar)) : "memory");
}

__device__ void stream_workers(const TileParams& P, const StreamSmem& sm, const SlotMeta* metas, uint64_t* full,
                               uint64_t* empty, int tid, long long* st_cycles) {
  const int lane = tid & 31, warp = tid >> 5;
  const int groups = P.slots >> 1;
  const int grp = warp / P.wps, sw = warp - grp * P.wps;
  int slot = grp;            // this group's slots: grp and grp + groups, used alternately
  uint32_t phase = 0;        // both slots are on the same use number until the second one has been consumed
  long long c_wait = 0, c_red = 0, c_scale = 0;
  const bool timing = st_cycles != nullptr && tid == 0;
  for (;;) {
    const long long t0 = timing ? clk() : 0;
    u_mbar_wait(&full[slot], phase);
    const long long t1 = timing ? clk() : 0;
    const SlotMeta m = metas[slot];
    if (m.kind == kItemStop) break;
    if (TILE_NODEPS(P) == 2 || TILE_NODEPS(P) == 3 || TILE_NODEPS(P) >= 7) {  // measurement only (7 / 8: R / S stage alone, loads only): no compute, just recycle the slot
      if (TILE_NODEPS(P) == 3 && lane == 0) { volatile float sink = sm.chunk(slot, 0)[tid]; (void)sink; }
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(&empty[slot]);
    } else if (m.kind == kItemR) {
      if (P.bwd) reduce_chunk_t<true>(P, m, sm.chunk(slot, 0), sm.chunk(slot, 1), sm.gate(slot, P.nbuf), sw, lane);
      else reduce_chunk_t<false>(P, m, sm.chunk(slot, 0), sm.chunk(slot, 1), sm.gate(slot, P.nbuf), sw, lane);
      __syncwarp();
      if (lane == 0) u_mbar_arrive(&empty[slot]);  // release.cta: the loader publishes the R results after its wait
    } else {
      if (P.bwd) scale_chunk_t<true>(P, m, sm.chunk(slot, 0), sm.gate(slot, P.nbuf), sm.add(slot, P.nbuf), sw, lane);
      else scale_chunk_t<false>(P, m, sm.chunk(slot, 0), sm.gate(slot, P.nbuf), sm.add(slot, P.nbuf), sw, lane);
      __syncwarp();
      // nothing to publish: the arrival only says that this warp's shared-memory reads are done (their values
      // have been consumed by the stores issued above)
      if (lane == 0) mbar_arrive_relaxed(&empty[slot]);
    }
    if (timing) {
      const long long t2 = clk();
      c_wait += t1 - t0;
      if (m.kind == kItemR) c_red += t2 - t1; else c_scale += t2 - t1;
    }
    if (slot == grp) slot = grp + groups;
    else { slot = grp; phase ^= 1u; }
  }
  if (timing) { st_cycles[4] = c_wait; st_cycles[5] = c_red; st_cycles[6] = c_scale; }
}

// =================================================================================================================
// GEMM role
// =================================================================================================================
struct GemmBars { uint64_t* raw; uint64_t* full; uint64_t* empty; uint64_t* acc_full; uint64_t* acc_empty; };

struct GemmItem { int stage, tile, ntile, split; };

// 2-D TMA tile load: box (32 k x 32 rows) of an fp32 [rows, K] tensor -> shared memory, SWIZZLE_128B
// (called by all 32 lanes of a converged warp; one elected lane issues)
__device__ __forceinline__ void tma_load_2d_elect(uint32_t dst, const CUtensorMap* map, int k, int row, uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n\t}"
      ::"r"(dst), "l"(map), "r"(k), "r"(row), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_elect(uint32_t bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}"
      ::"r"(bar), "r"(bytes)
      : "memory");
}

// One 128 x 128 x [k0, k1) product on tcgen05 (3xTF32).  Workers leave with the fp32 result of their row x 64
// columns in `sum`.  kbase / gbase: k-tiles and accumulator groups this CTA has pushed through its barriers so far.
//
// Per k-tile of 32: the control thread (warp 8) fetches the raw fp32 operand tiles with TMA (tensor maps with
// SWIZZLE_128B write the canonical K-major UMMA layout directly, out-of-range rows / k arrive as zeros); the raw tile IS
// the "big" TF32 operand (the tensor core reads the top 19 bits of each word); the 256 workers only compute the
// "small" twin (x - trunc(x)) and hand the stage to the MMA issue.  No thread that fences (fence.proxy.async is a
// MEMBAR) has global loads in flight, so the TMA prefetch of the next stages is never serialised.
__device__ __forceinline__ void gemm_item_mainloop(const TileParams& P, const GemmStage& g, const GemmItem& it,
                                                   unsigned char* u_smem, const GemmBars& bars, uint32_t tmem,
                                                   uint32_t kbase, uint32_t gbase, int nk, float (&sum)[64], int warp,
                                                   int lane, int tid, long long* dbg) {  // (kbase, gbase, nk by value)
  const int m0 = g.m_rows ? it.tile * UM : it.tile * P.m_tile;
  const int n0 = it.ntile * UN;
  const int k_begin = it.split * g.k_per_split;
  if (warp >= U_PRODUCERS / 32) {
    // ===== control warps (all 32 lanes converged, see umma.cuh): warp 9 produces (TMA), warp 8 issues the MMAs =======
    const uint32_t ub = 0xffffffffu;
    kbase = __shfl_sync(ub, kbase, 0); gbase = __shfl_sync(ub, gbase, 0); nk = __shfl_sync(ub, nk, 0);
    const uint32_t smem_base = __shfl_sync(ub, u_smem_addr(u_smem), 0);
    if (warp == U_PRODUCERS / 32 + 1) {
      const int um0 = __shfl_sync(ub, m0, 0), un0 = __shfl_sync(ub, n0, 0), uk0 = __shfl_sync(ub, k_begin, 0);
      const uint32_t bar_raw = __shfl_sync(ub, u_smem_addr(bars.raw), 0);
      const int k_split = g.k_split, n_split = g.n_split;
      const bool b_one = g.b_box_rows == 128;
      long long c0 = dbg ? clk() : 0;
      for (int kt = 0; kt < nk; ++kt) {
        const uint32_t KT = kbase + (uint32_t)kt, stage = KT % USTAGES;
        // the stage was last read by the MMAs of k-tile KT - USTAGES
        if (KT >= (uint32_t)USTAGES) u_mbar_wait(&bars.empty[stage], ((KT / USTAGES) - 1u) & 1u);
        const uint32_t st = smem_base + stage * U_STAGE_BYTES, bar = bar_raw + stage * 8u;
        mbar_expect_tx_elect(bar, 2u * U_TILE_BYTES);
        const int k0 = uk0 + kt * UK;
        const bool second = k_split && k0 >= k_split;
        // A: one 128-row box
        tma_load_2d_elect(st, second ? &g.tm_a2 : &g.tm_a, second ? k0 - k_split : k0, um0, bar);
        if (b_one) {
          const bool hi = un0 >= n_split;
          tma_load_2d_elect(st + 2 * U_TILE_BYTES, hi ? &g.tm_b2 : &g.tm_b, k0, hi ? un0 - n_split : un0, bar);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int r = un0 + 32 * j;
            const bool hi = r >= n_split;
            tma_load_2d_elect(st + 2 * U_TILE_BYTES + j * 4096, hi ? &g.tm_b2 : &g.tm_b, k0, hi ? r - n_split : r, bar);
          }
        }
      }
      if (dbg && lane == 0) dbg[2] += clk() - c0;
      __syncwarp();
      return;
    }
    const uint32_t utmem = __shfl_sync(ub, tmem, 0);
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(UN >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);
    for (int kt = 0; kt < nk; ++kt) {
      const uint32_t KT = kbase + (uint32_t)kt, GI = gbase + (uint32_t)(kt / UGROUP);
      const uint32_t stage = KT % USTAGES, b = GI & 1u;
      const long long c0 = dbg ? clk() : 0;
      if (kt % UGROUP == 0 && GI >= 2) u_mbar_wait(&bars.acc_empty[b], ((GI >> 1) - 1u) & 1u);
      u_mbar_wait(&bars.full[stage], (KT / USTAGES) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const long long c1 = dbg ? clk() : 0;
      u_mma_stage_elect(smem_base, stage * U_STAGE_BYTES, utmem + b * (uint32_t)UN, kt % UGROUP == 0, idesc);
      u_commit_elect(&bars.empty[stage]);
      if (kt % UGROUP == UGROUP - 1 || kt == nk - 1) u_commit_elect(&bars.acc_full[b]);
      if (dbg && lane == 0) { const long long c2 = clk(); dbg[0] += c1 - c0; dbg[1] += c2 - c1; }
    }
    __syncwarp();
    return;
  }
  // ===== workers: operand split + accumulator drain ==============================================================
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
  const bool wdbg = dbg != nullptr && tid == 0;
  for (int kt = 0; kt < nk; ++kt) {
    const uint32_t KT = kbase + (uint32_t)kt, stage = KT % USTAGES;
    const long long w0 = wdbg ? clk() : 0;
    u_mbar_wait(&bars.raw[stage], (KT / USTAGES) & 1u);
    const long long w1 = wdbg ? clk() : 0;
    unsigned char* st = u_smem + stage * U_STAGE_BYTES;
    // small = x - (x with the low 13 mantissa bits cleared), same layout as the raw tile: a linear pass
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4* raw = reinterpret_cast<const float4*>(st + h * 2 * U_TILE_BYTES);
      float4* sml = reinterpret_cast<float4*>(st + h * 2 * U_TILE_BYTES + U_TILE_BYTES);
      float4 x[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = raw[tid + j * U_PRODUCERS];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4 r;
        r.x = x[j].x - __uint_as_float(__float_as_uint(x[j].x) & 0xffffe000u);
        r.y = x[j].y - __uint_as_float(__float_as_uint(x[j].y) & 0xffffe000u);
        r.z = x[j].z - __uint_as_float(__float_as_uint(x[j].z) & 0xffffe000u);
        r.w = x[j].w - __uint_as_float(__float_as_uint(x[j].w) & 0xffffe000u);
        sml[tid + j * U_PRODUCERS] = r;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA unit
    u_mbar_arrive(&bars.full[stage]);
    const long long w2 = wdbg ? clk() : 0;
    if (kt % UGROUP >= 1 && drained < kt / UGROUP) { drain(drained); ++drained; }
    if (wdbg) { const long long w3 = clk(); dbg[4] += w1 - w0; dbg[5] += w2 - w1; dbg[6] += w3 - w2; }
  }
  const long long w4 = wdbg ? clk() : 0;
  while (drained < ngroups) { drain(drained); ++drained; }
  if (wdbg) dbg[6] += clk() - w4;
}

// Epilogue.  The accumulators first go through shared memory (the operand stages are idle by then) so that
// every warp then handles whole ROWS: bias / mask reads and the output stores are coalesced 512-byte runs, and the
// code is one short loop (an epilogue unrolled over the 64 columns a thread holds is > 100 KB of instructions and
// thrashes the instruction cache).  Requires n_total, n_split and ldo to be multiples of 4.
constexpr int kTilePitch = UN + 4;  // floats per row of the staged tile: conflict-free 128-bit writes and reads

__device__ __forceinline__ void gemm_stage_tile(float* tile, const float (&sum)[64], int warp, int lane) {
  float* dst = tile + (size_t)(32 * (warp & 3) + lane) * kTilePitch + (warp >> 2) * 64;
#pragma unroll
  for (int c = 0; c < 16; ++c)
    *reinterpret_cast<float4*>(dst + 4 * c) = make_float4(sum[4 * c], sum[4 * c + 1], sum[4 * c + 2], sum[4 * c + 3]);
}

__device__ __noinline__ void gemm_store_rows(const TileParams& P, const GemmStage& g, int tile_idx, int ntile,
                                             float* tile, int warp, int lane, float* out_t, int ldt) {
  const int m0 = g.m_rows ? tile_idx * UM : tile_idx * P.m_tile;
  const int rows = g.m_rows ? min(UM, g.m_rows - m0) : tile_rows(P, tile_idx);
  const int col = ntile * UN + 4 * lane;
  if (col < g.n_total) {
    const bool hi = col >= g.n_split;
    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g.bias) bias = __ldg(reinterpret_cast<const float4*>(hi ? g.bias2 + (col - g.n_split) : g.bias + col));
    float* obase = hi ? g.out2 + (col - g.n_split) : g.out + col;
    const int epi = g.epi;
    const float div = g.div;
    for (int r = warp; r < rows; r += U_PRODUCERS / 32) {
      float4 v = *reinterpret_cast<const float4*>(tile + (size_t)r * kTilePitch + 4 * lane);
      v.x += bias.x; v.y += bias.y; v.z += bias.z; v.w += bias.w;
      if (epi == kEpiRelu) {
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      } else if (epi == kEpiSigmoid) {
        v.x = sigmoidf_fast(v.x); v.y = sigmoidf_fast(v.y); v.z = sigmoidf_fast(v.z); v.w = sigmoidf_fast(v.w);
      } else if (epi == kEpiMask) {
        const float4 mk = __ldg(reinterpret_cast<const float4*>(g.mask + (size_t)(m0 + r) * g.ldmask + col));
        v.x = mk.x > 0.f ? v.x : 0.f; v.y = mk.y > 0.f ? v.y : 0.f; v.z = mk.z > 0.f ? v.z : 0.f; v.w = mk.w > 0.f ? v.w : 0.f;
      } else {
        v.x = v.x / div; v.y = v.y / div; v.z = v.z / div; v.w = v.w / div;
      }
      const int gr = m0 + r;
      float* o = (g.m_split && gr >= g.m_split) ? g.out2 + (size_t)(gr - g.m_split) * g.ldo + col
                                                : obase + (size_t)gr * g.ldo;
      *reinterpret_cast<float4*>(o) = v;
      if (out_t) *reinterpret_cast<float4*>(tile + (size_t)r * kTilePitch + 4 * lane) = v;  // final value for the pass below
    }
  }
  if (out_t) {
    // K-major copy for the weight-gradient GEMMs: out_t[column][sample].  Column j of the staged tile goes out as one
    // run of `rows` consecutive floats (the eight worker warps take columns j = warp, warp + 8, ...).
    asm volatile("bar.sync 1, %0;" ::"n"(U_PRODUCERS) : "memory");
    const int ncols = min(UN, g.n_total - ntile * UN);
    for (int j = warp; j < ncols; j += U_PRODUCERS / 32) {
      float* dst = out_t + (size_t)(ntile * UN + j) * ldt + m0;
#pragma unroll
      for (int q = 0; q < UM / 32; ++q) {
        const int r = lane + 32 * q;
        if (r < rows) dst[r] = tile[(size_t)r * kTilePitch + j];
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

// Generic 32 x 32 transposes (any rows / cols) by ONE warp per party: the tiles of all jobs form one index space, party
// `part` of `nparts` takes every nparts-th tile.  All 32 row loads of a tile are issued before the first one is used:
// while the stream CTAs saturate HBM a load takes several microseconds, so the depth in flight is what counts.
// Used to give the weight-gradient GEMMs K-major operands: dW = X^T Y reduces over the batch, and a TF32 UMMA operand
// must have its reduction dimension contiguous.
__device__ void transpose_jobs_warp(const TrJob* jobs, int njobs, int part, int nparts, float (*s)[33], int lane) {
  int base = 0;
  for (int j = 0; j < njobs; ++j) {
    const TrJob& t = jobs[j];
    const int tr = (t.rows + 31) / 32, tc = (t.cols + 31) / 32, total = tr * tc;
    int first = (part - base) % nparts;
    if (first < 0) first += nparts;
    for (int tile = first; tile < total; tile += nparts) {
      const int r0 = (tile / tc) * 32, c0 = (tile % tc) * 32;
      const bool cin = c0 + lane < t.cols;
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i)
        v[i] = (cin && r0 + i < t.rows) ? __ldcg(t.in + (size_t)(r0 + i) * t.ld_in + c0 + lane) : 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) s[i][lane] = v[i];
      __syncwarp();
      const bool rout = r0 + lane < t.rows;
#pragma unroll 8
      for (int i = 0; i < 32; ++i)
        if (rout && c0 + i < t.cols) t.out[(size_t)(c0 + i) * t.ld_out + r0 + lane] = s[lane][i];
      __syncwarp();
    }
    base += total;
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

__device__ void gemm_role(const TileParams& P, int grank, unsigned char* u_smem, int tid, long long* st_cycles) {
  __shared__ __align__(8) uint64_t bar_raw[USTAGES], bar_full[USTAGES], bar_empty[USTAGES], bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t s_tmem;
  __shared__ unsigned s_ticket, s_flag;
  const int warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < USTAGES; ++s) {
      u_mbar_init(&bar_raw[s], 1); u_mbar_init(&bar_full[s], U_PRODUCERS); u_mbar_init(&bar_empty[s], 1);
    }
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
  const GemmBars bars{bar_raw, bar_full, bar_empty, bar_acc_full, bar_acc_empty};

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
  // ticket space: [FC chain items of every tile][column sums][weight-gradient GEMM items]
  const unsigned t_cs = per_tile * (unsigned)T, t_post = t_cs + (unsigned)cs_blocks[4];
  unsigned post_items[2] = {0u, 0u};
  if (P.n_post) {
    post_items[0] = (unsigned)(((P.st[2].m_rows + UM - 1) / UM) * P.st[2].n_tiles);
    post_items[1] = (unsigned)(((P.st[3].m_rows + UM - 1) / UM) * P.st[3].n_tiles);
  }
  const unsigned total = t_post + post_items[0] + post_items[1];
  uint32_t kbase = 0, gbase = 0;
  float sum[64];
  long long c_dep = 0, c_main = 0, c_epi = 0, n_items = 0, c_e[4] = {0, 0, 0, 0};
  const bool timing = st_cycles != nullptr && tid == 0;

  for (;;) {
    if (tid == 0) s_ticket = atomicAdd(&P.ctr[1], 1u);
    __syncthreads();
    const unsigned ticket = s_ticket;
    if (ticket >= total || TILE_NODEPS(P) >= 5) break;
    if (ticket >= t_cs && ticket < t_post) {
      // ---- column-sum item: needs every tile's R, F1 and F2 -----------------------------------------------------
      if (tid == 0) {
        for (int t = 0; t < T; ++t) {
          const unsigned rneed = 2u * (unsigned)((size_t)tile_rows(P, t) * P.c / P.p);
          wait_counter(tile_ctr(P, t) + 0, rneed);
          wait_counter(tile_ctr(P, t) + 1, (unsigned)P.st[0].n_tiles);
          wait_counter(tile_ctr(P, t) + 2, (unsigned)P.st[1].n_tiles);
        }
      }
      __syncthreads();
      int blk = (int)(ticket - t_cs), which = 0;
      while (which < 3 && blk >= cs_blocks[which + 1]) ++which;
      colsum_item(P.cs[which], blk - cs_blocks[which], reinterpret_cast<float*>(u_smem), tid);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // scratch is a later item's TMA destination
      continue;
    }
    GemmItem it;
    int r;
    if (ticket >= t_post) {   // weight-gradient item: (m-tile, n-tile) of post stage 2 or 3
      unsigned q = ticket - t_post;
      it.stage = q >= post_items[0] ? 3 : 2;
      if (it.stage == 3) q -= post_items[0];
      it.tile = (int)(q / (unsigned)P.st[it.stage].n_tiles);
      r = (int)(q - (unsigned)it.tile * (unsigned)P.st[it.stage].n_tiles);
    } else {
      it.tile = (int)(ticket / per_tile);
      r = (int)(ticket - (unsigned)it.tile * per_tile);
      it.stage = r >= i1 ? 1 : 0;
      if (it.stage) r -= i1;
    }
    const GemmStage& g = P.st[it.stage];
    const int rr = r;
    it.ntile = rr / g.splits;
    it.split = rr - it.ntile * g.splits;
    unsigned* tc = tile_ctr(P, it.stage < 2 ? it.tile : 0);
    const long long tg0 = timing ? clk() : 0;
    const unsigned long long t_item0 = (tid == 0 && TILE_STATS(P)) ? global_ns() : 0ull;
    if (tid == 0) {
      if (P.w_cat_t) wait_counter(&P.ctr[3], (unsigned)P.n_gemm);
      if (it.stage >= 2) {
        // dE^T is complete with every tile's R stage, dH^T with every tile's dH GEMM, H^T / Z^T when every original
        // stream CTA's warp 9 has reported
        for (int t = 0; t < T; ++t) {
          const unsigned rneed = 2u * (unsigned)((size_t)tile_rows(P, t) * P.c / P.p);
          wait_counter(tile_ctr(P, t) + 0, rneed);
          wait_counter(tile_ctr(P, t) + 1, (unsigned)P.st[0].n_tiles);
        }
        wait_counter(&P.ctr[4], gridDim.x - (unsigned)P.n_gemm);
      } else if (it.stage == 0) {
        const unsigned rneed = 2u * (unsigned)((size_t)tile_rows(P, it.tile) * P.c / P.p);  // one per R item
        wait_counter(tc + 0, rneed);
        // the partial planes of this ring position were last used by tile - kRingTiles: fully folded?
        // (only with split-K: without it this would hold every tile's chain behind the one four tiles ahead --
        // measured as a 25 us stall per tile at 512x7^2, profiles/r2_tile_timeline_512.txt)
        if (it.tile >= kRingTiles && (P.st[0].splits > 1 || P.st[1].splits > 1)) wait_counter(tile_ctr(P, it.tile - kRingTiles) + 2, (unsigned)P.st[1].n_tiles);
      } else {
        wait_counter(tc + 1, (unsigned)P.st[0].n_tiles);
      }
    }
    __syncthreads();
    const int k_begin = it.split * g.k_per_split;
    const int k_end = min(g.k_total, k_begin + g.k_per_split);
    const int nk = k_end > k_begin ? (k_end - k_begin + UK - 1) / UK : 0;
    const long long tg1 = timing ? clk() : 0;
    if (tid == 0 && TILE_STATS(P) && ticket < 256) {
      long long* row = TILE_STATS(P) + (size_t)(kTraceGemmRow + ticket) * 16;
      row[0] = it.tile; row[1] = it.stage; row[2] = it.ntile; row[3] = grank; row[4] = (long long)t_item0; row[5] = (long long)global_ns();
    }
    gemm_item_mainloop(P, g, it, u_smem, bars, tmem, kbase, gbase, nk, sum, warp, lane, tid,
                       st_cycles ? TILE_STATS(P) + (size_t)(gridDim.x + blockIdx.x) * 16 : nullptr);
    const long long tg2 = timing ? clk() : 0;
    if (tid == 0 && TILE_STATS(P) && ticket < 256) TILE_STATS(P)[(size_t)(kTraceGemmRow + ticket) * 16 + 6] = (long long)global_ns();
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
    const bool timing2 = st_cycles != nullptr && (tid == 128 || tid == 256);
    const long long tg3 = (timing || timing2) ? clk() : 0;
    long long tg4 = tg3, tg5 = tg3;
    if (finish) {
      if (warp < 8) gemm_stage_tile(reinterpret_cast<float*>(u_smem), sum, warp, lane);
      __syncthreads();
      if (warp < 8)
        gemm_store_rows(P, g, it.tile, it.ntile, reinterpret_cast<float*>(u_smem), warp, lane,
                        (it.stage == 0 && P.dh_t) ? P.dh_t : nullptr, P.ldt);
      // the staged tile lives in the operand stages: order these generic accesses before the next item's TMA writes
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tg4 = (timing || timing2) ? clk() : 0;
      __threadfence();
      tg5 = (timing || timing2) ? clk() : 0;
      __syncthreads();
      if (timing2) {
        const long long t6 = clk();
        if (tid == 128) { st_cycles[12] += tg4 - tg3; st_cycles[13] += tg5 - tg4; }
        else { st_cycles[14] += tg5 - tg4; st_cycles[15] += t6 - tg5; }
      }
      if (tid == 0 && it.stage < 2) red_release_add(tc + 1 + it.stage, 1u);
      if (tid == 0 && TILE_STATS(P) && ticket < 256) TILE_STATS(P)[(size_t)(kTraceGemmRow + ticket) * 16 + 7] = (long long)global_ns();
    }
    if (timing) {
      const long long tg6 = clk();
      c_dep += tg1 - tg0; c_main += tg2 - tg1; c_epi += tg6 - tg2; ++n_items;
      c_e[0] += tg3 - tg2; c_e[1] += tg4 - tg3; c_e[2] += tg5 - tg4; c_e[3] += tg6 - tg5;
    }
  }
  if (timing) {
    st_cycles[1] = n_items; st_cycles[2] = c_dep; st_cycles[3] = c_main; st_cycles[4] = c_epi;
    st_cycles[8] = c_e[0]; st_cycles[9] = c_e[1]; st_cycles[10] = c_e[2]; st_cycles[11] = c_e[3];
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
  __shared__ float s_tr[32][33];
  const int tid = threadIdx.x;
  if (tid == 0) s_role = atomicAdd(&P.ctr[2], 1u);
  __syncthreads();
  const int role = (int)s_role;
  long long* st_cycles = (TILE_STATS(P) && !TILE_TRACE_ONLY(P)) ? TILE_STATS(P) + (size_t)blockIdx.x * 16 : nullptr;
  const long long t_begin = st_cycles ? clk() : 0;
  if (tid == 0) trace_min(TILE_STATS(P), kTraceStartRow, 0);
  if (role < P.n_gemm) {
    gemm_role(P, role, smem, tid, st_cycles);
    if (st_cycles && tid == 0) { st_cycles[0] = 1; st_cycles[7] = clk() - t_begin; }
    // The last GEMM ticket is drawn when the last tile's reduce stage is complete, i.e. when the second half of the
    // stream queue (S items of the last `lag` tiles) is still ahead.  The CTA then joins the stream role: tickets are
    // drawn dynamically, so a late joiner simply takes the next ones.  (512x7^2: 48 of 148 SMs would otherwise idle
    // for the last ~60 % of the launch.)
    if (!P.gemm_joins_stream) return;
    st_cycles = nullptr;
    __syncthreads();
  }
  if (tid == 0) {
    for (int s = 0; s < P.slots; ++s) { u_mbar_init(&s_full[s], 1); u_mbar_init(&s_empty[s], P.wps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  StreamSmem sm{smem, P.slot_bytes, P.chunk_bytes, P.p};
  if (tid >= U_PRODUCERS + 32) {
    // warp 9 has no part in streaming: in the CTAs that started in the stream role it transposes H and Z for the
    // weight-gradient GEMMs (no dependencies: forward results), one 32 x 32 tile at a time, then reports
    if (P.n_tr_pre && role >= P.n_gemm) {
      transpose_jobs_warp(P.tr_pre, P.n_tr_pre, role - P.n_gemm, (int)gridDim.x - P.n_gemm, s_tr, tid & 31);
      __threadfence();
      __syncwarp();
      if ((tid & 31) == 0) red_release_add(&P.ctr[4], 1u);
    }
  } else if (tid >= U_PRODUCERS) {
    stream_loader(P, sm, s_meta, s_full, s_empty, st_cycles, tid & 31);
  } else {
    stream_workers(P, sm, s_meta, s_full, s_empty, tid, st_cycles);
    if (st_cycles && tid == 0) { st_cycles[0] = 0; st_cycles[7] = clk() - t_begin; }
  }
}

// ---- host side -------------------------------------------------------------------------------------------------
struct TileCfg {
  int m_tile, n_tiles, p, lanes, lanes_log2, n_gemm, slots, wps, nbuf, lag;
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
  // forward chunks are doubled once a modality has >= ~190 MB: with fewer items in total the finer interleave of R and
  // S items matters more than the per-item cost (256x14^2 at batch 512: 0.129 ms with 28 KB vs 0.137 ms with 56 KB)
  const bool big_fwd = !bwd && (size_t)n * c * hw * 4 >= ((size_t)g_tile_min_mb_light << 20);
  const size_t bound = (size_t)(big_fwd ? g_tile_chunk_kb_fwd : g_tile_chunk_kb) * 1024;
  f.p = 0;
  for (int p = c; p >= 4; --p) {
    if (c % p || p % 4) continue;
    if ((size_t)p * hw * 4 <= bound) { f.p = p; break; }
  }
  if (!f.p) return false;
  f.chunk_bytes = (uint32_t)((size_t)f.p * hw * 4);
  // lanes per plane inside a worker warp: 128-bit walks want >= 8 lanes on a plane (a quarter warp then reads 128
  // contiguous bytes: conflict-free) and a few vectors per lane; scalar walks (HW % 4 != 0) of short planes use one
  // lane per plane (an odd HW makes the 32 planes of a warp hit 32 different banks) and need no shuffles
  int lanes = 1, lg = 0;
  if (hw % 4 == 0) {
    const int hw4 = hw / 4;
    while (lanes < 32 && lanes * 8 < hw4) { lanes *= 2; ++lg; }   // <= 8 vectors per lane: one batch of loads
    if (lanes < 8 && hw4 >= 8) { lanes = 8; lg = 3; }
  } else {
    while (lanes < 32 && lanes * 64 < hw) { lanes *= 2; ++lg; }
  }
  f.lanes = lanes;
  f.lanes_log2 = lg;
  f.nbuf = bwd ? 2 : 1;
  f.slot_bytes = (uint32_t)round_up((size_t)f.nbuf * f.chunk_bytes + 2 * (size_t)f.p * 4, 128);
  const size_t gemm_smem = (size_t)USTAGES * U_STAGE_BYTES;
  const size_t budget = 218 * 1024;  // 227 KB - static shared memory (5 KB) - alignment slack - 1 KB header
  int slots = (int)(budget / f.slot_bytes);
  if (slots < 2) return false;
  slots = slots >= 8 ? 8 : (slots >= 4 ? 4 : 2);  // a power of two; a group of warps owns two slots
  if (g_tile_max_slots >= 2 && slots > g_tile_max_slots) slots = g_tile_max_slots >= 8 ? 8 : (g_tile_max_slots >= 4 ? 4 : 2);
  f.slots = slots;
  f.wps = (U_PRODUCERS / 32) / (slots / 2);
  size_t sm = (size_t)slots * f.slot_bytes;
  if (sm < gemm_smem) sm = gemm_smem;
  f.smem_bytes = sm + 1024;
  // tile: ~25 MB of resident feature map (both modalities), at least ~4 tiles for the pipeline, <= 128 samples
  const size_t per_sample = (size_t)2 * c * hw * 4;
  // tile: up to 128 samples (a full MMA), at least ~4 tiles for the pipeline
  long m = g_tile_m > 0 ? g_tile_m : UM;
  if (m > UM) m = UM;
  if (g_tile_m <= 0 && m > (n + 3) / 4) m = (n + 3) / 4;
  if (g_tile_m <= 0 && m > 8) m &= ~7L;
  if (m < 1) m = 1;
  f.m_tile = (int)m;
  f.n_tiles = (n + f.m_tile - 1) / f.m_tile;
  f.lag = g_tile_lag > 0 ? g_tile_lag : 8;  // all R stages first when there are <= 8 tiles: measured best (profiles/r2_sweep.md)
  if (f.lag > f.n_tiles) f.lag = f.n_tiles;
  // GEMM stages: forward (K = 2C -> N = D), (K = D -> N = 2C); backward (K = 2C -> N = D), (K = D -> N = 2C)
  const int kk[2] = {2 * c, d}, nn[2] = {d, 2 * c};
  size_t items = 0, units = 0;
  for (int s = 0; s < 2; ++s) {
    f.nt[s] = ceil_div(nn[s], UN);
    const int nk = ceil_div(kk[s], UK);
    int sp = g_tile_ksplit_tiles > 0 ? ceil_div(nk, g_tile_ksplit_tiles) : 1;
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
    // The 3xTF32 FC work of a tile grows with C^2 while its bytes grow with C.  Measured best on B200 with GEMM CTAs
    // that join the stream role afterwards (profiles/r2_sweep.md): forward 16 CTAs for C = 256, 48-56 for 512, i.e.
    // ~C*D / 4096; the backward also runs the weight-gradient GEMMs and wants ~C*D / 3277 (20 and 64).
    (void)units; (void)items;
    const bool fold = bwd && g_tile_wgrad;
    g = (int)((double)c * d / (fold ? 3277.0 : 4096.0) + 0.5);
    if (g < 3) g = 3;
    if (g > (fold ? 64 : 56)) g = fold ? 64 : 56;
  }
  const int sms = sm_count();
  if (g > sms / 2) g = sms / 2;
  if (g < 1) g = 1;
  f.n_gemm = g;
  *o = f;
  return true;
}

// ---- TMA tensor maps (driver entry point resolved at run time: no link-time dependency on libcuda) ---------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// fp32 [rows, k] with leading dimension ld (elements): box 32 k x 32 rows, 128-byte swizzle, zeros out of range
int make_map(CUtensorMap* m, const float* base, int rows, int k, int ld, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return GML_E_UNSUPPORTED;
  const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)UK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? GML_OK : GML_E_UNSUPPORTED;
}

void fill_common(TileParams& P, const TileCfg& f, int n, int c, int hw, int d, bool bwd, void* ws) {
  P.n = n; P.c = c; P.hw = hw; P.d = d; P.bwd = bwd ? 1 : 0;
  P.m_tile = f.m_tile; P.n_tiles = f.n_tiles; P.p = f.p; P.lanes = f.lanes; P.lanes_log2 = f.lanes_log2; P.lag = f.lag;
  P.n_gemm = f.n_gemm; P.slots = f.slots; P.wps = f.wps; P.nbuf = f.nbuf; P.chunk_bytes = f.chunk_bytes; P.slot_bytes = f.slot_bytes;
  P.hw_magic = (uint32_t)(((1ull << 32) + (unsigned)hw - 1) / (unsigned)hw);
  P.cps = c / f.p;
  P.cps_magic = P.cps == 1 ? 0u : (uint32_t)(((1ull << 32) + (unsigned)P.cps - 1) / (unsigned)P.cps);
  P.slots_log2 = f.slots == 8 ? 3 : (f.slots == 4 ? 2 : 1);
  P.ctr = reinterpret_cast<unsigned*>(ws);
  P.part = reinterpret_cast<float*>(static_cast<char*>(ws) + f.ctr_bytes);
  P.part_tile_floats = f.part_tile_floats;
  for (int s = 0; s < 4; ++s) {
    P.st[s].n_tiles = s < 2 ? f.nt[s] : 0; P.st[s].splits = s < 2 ? f.splits[s] : 1; P.st[s].k_per_split = s < 2 ? f.kps[s] : 0;
    P.st[s].k_split = 0; P.st[s].m_rows = 0; P.st[s].m_split = 0;
    P.st[s].b_box_rows = 128;
    P.st[s].out2 = nullptr; P.st[s].bias = nullptr; P.st[s].bias2 = nullptr; P.st[s].mask = nullptr; P.st[s].ldmask = 0;
    P.st[s].div = 1.f;
  }
  P.n_cs = 0;
  P.n_post = 0; P.n_tr_pre = 0; P.de_t = nullptr; P.dh_t = nullptr; P.ldt = 0;
  P.stats = g_tile_stats;
  P.nodeps = g_tile_nodeps;
  P.gemm_joins_stream = g_tile_switch;
  P.draw = g_tile_draw;
  P.split_copies = g_tile_split_copies;
  P.rpol = g_tile_rpol;
  P.trace_only = g_tile_trace_only;
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

// Automatic selection (measured on B200, profiles/r2_sweep.md): the pipeline streams at ~6 TB/s of real traffic but its S
// stage re-reads from HBM (the FC chain of a tile takes longer than L2 can hold the tile), i.e. it moves 6u / 8u.  That
// beats the streaming multi-kernel path and the cluster kernels' per-group weight re-reads for the weight-heavy blocks
// (C >= 256) once the batch is large; the 128-channel block stays on the cluster kernels (4u / 6u from L2-resident groups).
bool tile_preferred(int n, int c, int hw, int d, bool bwd) {
  if (g_tile_kind == 1) return true;
  const size_t u = (size_t)n * c * hw * 4;
  const size_t w_bytes = (size_t)16 * c * d;
  // Forward: 512x7^2 (4 MB of FC weights) from half the threshold (batch 512: 0.123 vs 0.126 ms), 256x14^2 (1 MB) only
  // from 1.5x of it -- its cluster kernel still wins at batch 512 (0.126 vs 0.133 ms) and loses at 1024 (0.290 vs 0.205)
  if (w_bytes >= (1u << 20)) {
    const size_t base = (size_t)g_tile_min_mb << 20;
    // backward: from half the threshold for both (256x14^2 at batch 256: 0.120 vs 0.126 ms, 512x7^2 at 512: 0.155 vs 0.168)
    return u >= (bwd ? base / 2 : (w_bytes >= (2u << 20) ? base / 2 : base * 3 / 2));
  }
  // light-weight blocks (128 channels): the cluster kernels move 4u / 6u and, with their weight slices in shared memory,
  // win both directions (128x28^2 forward at batch 1024: 0.357 ms vs 0.381 ms through the pipeline's 56 KB items)
  return !bwd && g_tile_light_fwd && u >= ((size_t)g_tile_min_mb_light << 20);
}

// K-major copies for the weight-gradient GEMMs: H^T, dH^T [D, ldT] and Z^T, dE^T [2C, ldT], ldT = N rounded up to 32
static size_t tile_wgrad_bytes(int n, int c, int d) {
  const size_t ldt = round_up((size_t)n, 32);
  return 2 * round_up((size_t)d * ldt * 4, 256) + 2 * round_up((size_t)2 * c * ldt * 4, 256);
}

size_t tile_fwd_workspace_bytes(int n, int c, int hw, int d) {
  TileCfg f;
  if (!make_tile_cfg(n, c, hw, d, false, &f)) return 0;
  return f.ctr_bytes + f.part_bytes + 256;
}

size_t tile_bwd_workspace_bytes(int n, int c, int hw, int d) {
  TileCfg f;
  if (!make_tile_cfg(n, c, hw, d, true, &f)) return 0;
  return f.ctr_bytes + f.part_bytes + 2 * round_up((size_t)2 * c * d * 4, 256) + tile_wgrad_bytes(n, c, d) + 256;
}

int launch_tile_fwd(const FusedFwdArgs& a, float* gate_sum, float* run_v, float* run_s, float step, void* ws,
                    size_t ws_bytes, cudaStream_t st) {
  TileCfg f;
  if (!make_tile_cfg(a.n, a.c, a.hw, a.d, false, &f)) return GML_E_UNSUPPORTED;
  if (!ws || ws_bytes < f.ctr_bytes + f.part_bytes) return GML_E_WORKSPACE;
  const void* al[] = {a.a, a.b, a.a_out, a.b_out, a.w_sq, a.w_v, a.w_s, a.b_sq, a.b_v, a.b_s, a.z, a.h, a.g_a, a.g_b, ws};
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
  const int bbox = a.c % 128 == 0 ? 128 : 32;  // B tiles of stage 1 must not straddle the W_v / W_s boundary
  GemmStage& s0 = P.st[0];   // H = relu(Z Wsq^T + bsq)
  GML_TRY(make_map(&s0.tm_a, a.z, a.n, 2 * a.c, 2 * a.c, 128));
  GML_TRY(make_map(&s0.tm_b, a.w_sq, a.d, 2 * a.c, 2 * a.c, 128));
  s0.tm_a2 = s0.tm_a; s0.tm_b2 = s0.tm_b;
  s0.n_split = a.d; s0.n_total = a.d; s0.k_total = 2 * a.c;
  s0.out = a.h; s0.ldo = a.d; s0.bias = a.b_sq; s0.epi = kEpiRelu;
  GemmStage& s1 = P.st[1];   // [g_a | g_b] = sigmoid(H [Wv ; Ws]^T + [bv | bs])
  GML_TRY(make_map(&s1.tm_a, a.h, a.n, a.d, a.d, 128));
  GML_TRY(make_map(&s1.tm_b, a.w_v, a.c, a.d, a.d, bbox));
  GML_TRY(make_map(&s1.tm_b2, a.w_s, a.c, a.d, a.d, bbox));
  s1.tm_a2 = s1.tm_a;
  s1.n_split = a.c; s1.n_total = 2 * a.c; s1.b_box_rows = bbox;
  s1.k_total = a.d; s1.out = a.g_a; s1.out2 = a.g_b; s1.ldo = a.c; s1.bias = a.b_v; s1.bias2 = a.b_s; s1.epi = kEpiSigmoid;
  if (gate_sum) {
    P.cs[0] = ColItem{a.g_a, gate_sum, a.n, a.c, a.c, run_v, run_s, (float)a.n, step};
    P.n_cs = 1;
  }
  return launch_pipeline(P, f, kTagFusedFwd, st);
}

int launch_tile_bwd(const FusedBwdArgs& a, float* dz_flat, float* d_b_v, float* d_b_s, float* d_b_sq, float* d_w_v,
                    float* d_w_s, float* d_w_sq, bool* wgrad_done, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (wgrad_done) *wgrad_done = false;
  TileCfg f;
  if (!make_tile_cfg(a.n, a.c, a.hw, a.d, true, &f)) return GML_E_UNSUPPORTED;
  const size_t wt = round_up((size_t)2 * a.c * a.d * 4, 256);
  if (!ws || ws_bytes < f.ctr_bytes + f.part_bytes + 2 * wt + tile_wgrad_bytes(a.n, a.c, a.d)) return GML_E_WORKSPACE;
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
  const int bbox = a.c % 128 == 0 ? 128 : 32;
  GemmStage& s0 = P.st[0];   // dH = ([dE_a | dE_b] [Wv ; Ws]) * [H > 0]
  GML_TRY(make_map(&s0.tm_a, a.de_a, a.n, a.c, a.c, 128));
  GML_TRY(make_map(&s0.tm_a2, a.de_b, a.n, a.c, a.c, 128));
  GML_TRY(make_map(&s0.tm_b, P.w_cat_t, a.d, 2 * a.c, 2 * a.c, 128));
  s0.tm_b2 = s0.tm_b;
  s0.k_split = a.c; s0.n_split = a.d; s0.n_total = a.d; s0.k_total = 2 * a.c;
  s0.out = a.dh; s0.ldo = a.d; s0.mask = a.h; s0.ldmask = a.d; s0.epi = kEpiMask;
  GemmStage& s1 = P.st[1];   // dZ = dH Wsq, stored per modality and already divided by HW (MeanBackward)
  // (w_sq_t is ONE [2C, D] matrix: the second map simply continues it at row C, where the output switches to dz_b)
  GML_TRY(make_map(&s1.tm_a, a.dh, a.n, a.d, a.d, 128));
  GML_TRY(make_map(&s1.tm_b, P.w_sq_t, a.c, a.d, a.d, bbox));
  GML_TRY(make_map(&s1.tm_b2, P.w_sq_t + (size_t)a.c * a.d, a.c, a.d, a.d, bbox));
  s1.tm_a2 = s1.tm_a;
  s1.n_split = a.c; s1.n_total = 2 * a.c; s1.b_box_rows = bbox; s1.k_total = a.d; s1.out = dz_a; s1.out2 = dz_b; s1.ldo = a.c; s1.epi = kEpiDiv;
  s1.div = (float)a.hw;
  int ncs = 0;
  if (d_b_v) P.cs[ncs++] = ColItem{a.de_a, d_b_v, a.n, a.c, a.c, nullptr, nullptr, 1.f, 0.f};
  if (d_b_s) P.cs[ncs++] = ColItem{a.de_b, d_b_s, a.n, a.c, a.c, nullptr, nullptr, 1.f, 0.f};
  if (d_b_sq) P.cs[ncs++] = ColItem{a.dh, d_b_sq, a.n, a.d, a.d, nullptr, nullptr, 1.f, 0.f};
  P.n_cs = ncs;
  // Weight gradients inside the same launch: dW_v | dW_s = dE^T H and dW_sq = dH^T Z reduce over the batch, so their
  // operands need K-major copies (see TileParams); the GEMM CTAs then run them as ordinary 128 x 128 items once the last
  // tile's chain is done, while the stream CTAs are still busy with the S stage.
  const bool fold = g_tile_wgrad && d_w_v && d_w_s && d_w_sq && a.z && aligned16(d_w_v) && aligned16(d_w_s) &&
                    aligned16(d_w_sq) && aligned16(a.z);
  if (fold) {
    const int ldt = (int)round_up((size_t)a.n, 32);
    char* tp = wp + 2 * wt;
    float* h_t = reinterpret_cast<float*>(tp);  tp += round_up((size_t)a.d * ldt * 4, 256);
    float* dh_t = reinterpret_cast<float*>(tp); tp += round_up((size_t)a.d * ldt * 4, 256);
    float* z_t = reinterpret_cast<float*>(tp);  tp += round_up((size_t)2 * a.c * ldt * 4, 256);
    float* de_t = reinterpret_cast<float*>(tp);
    P.tr_pre[0] = TrJob{a.h, h_t, a.n, a.d, a.d, ldt};
    P.tr_pre[1] = TrJob{a.z, z_t, a.n, 2 * a.c, 2 * a.c, ldt};
    P.n_tr_pre = 2;
    P.de_t = de_t; P.dh_t = dh_t; P.ldt = ldt;
    const int kps = (int)round_up((size_t)a.n, UK);
    GemmStage& s2 = P.st[2];   // [dW_v ; dW_s] = dE^T H      ([2C, N] x [N, D])
    GML_TRY(make_map(&s2.tm_a, de_t, 2 * a.c, a.n, ldt, 128));
    GML_TRY(make_map(&s2.tm_b, h_t, a.d, a.n, ldt, a.d % 128 == 0 ? 128 : 32));
    s2.tm_a2 = s2.tm_a; s2.tm_b2 = s2.tm_b;
    s2.m_rows = 2 * a.c; s2.m_split = a.c; s2.n_split = a.d; s2.n_total = a.d; s2.k_total = a.n;
    s2.b_box_rows = a.d % 128 == 0 ? 128 : 32;
    s2.n_tiles = (a.d + UN - 1) / UN; s2.splits = 1; s2.k_per_split = kps;
    s2.out = d_w_v; s2.out2 = d_w_s; s2.ldo = a.d; s2.epi = kEpiDiv; s2.div = 1.f;
    GemmStage& s3 = P.st[3];   // dW_sq = dH^T Z             ([D, N] x [N, 2C])
    GML_TRY(make_map(&s3.tm_a, dh_t, a.d, a.n, ldt, 128));
    GML_TRY(make_map(&s3.tm_b, z_t, 2 * a.c, a.n, ldt, (2 * a.c) % 128 == 0 ? 128 : 32));
    s3.tm_a2 = s3.tm_a; s3.tm_b2 = s3.tm_b;
    s3.m_rows = a.d; s3.m_split = 0; s3.n_split = 2 * a.c; s3.n_total = 2 * a.c; s3.k_total = a.n;
    s3.b_box_rows = (2 * a.c) % 128 == 0 ? 128 : 32;
    s3.n_tiles = (2 * a.c + UN - 1) / UN; s3.splits = 1; s3.k_per_split = kps;
    s3.out = d_w_sq; s3.out2 = nullptr; s3.ldo = 2 * a.c; s3.epi = kEpiDiv; s3.div = 1.f;
    P.n_post = 2;
  }
  const int rc = launch_pipeline(P, f, kTagFusedBwd, st);
  if (rc == GML_OK && wgrad_done) *wgrad_done = fold;
  return rc;
}

}  // namespace gml
