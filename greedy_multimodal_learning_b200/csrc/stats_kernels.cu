// Analysis statistics of the hot path:
//   * multi-tensor sum of squares for the conditional learning speed (one launch for all
//     parameters and gradients; replaces src/callbacks.py:203-205 -- 2 reductions and 2 .item()
//     host syncs PER TENSOR in the reference),
//   * selected-row squeeze accumulation for the conditional utilization rate
//     (src/balanced_mmtm.py:186-201),
//   * argmax/equality accuracy counts (train.py:32-40).
// All three are HBM-bound scans; no tensor cores, no atomics on the data path, fixed reduction
// order (bit-reproducible).
#include <cooperative_groups.h>

#include "common.cuh"

namespace gml {

// tunable "sq_variant" (measurement): bits 0-1 chunk {4096, 2048, 1024, 8192} elements, bit 2 = no programmatic
// dependent launch of the fold, bits 3-6 blocks per SM (0 = 8)
int g_sq_variant = 0;

namespace {

constexpr int kMaxTensors = 1024;  // per launch; 14.3 KB of kernel parameters (CUDA >= 12.1 allows 32 KB)
constexpr int kMinChunk = 1024;    // smallest chunk (elements) a variant may use: sizes the workspace
constexpr int kSqThreads = 256;

struct SqnormTable {
  const float* ptr[kMaxTensors];
  int chunk_start[kMaxTensors + 1];  // prefix sum of per-tensor chunk counts
  unsigned char mask[kMaxTensors];   // GML_BUCKET_* bits
  unsigned char kind[kMaxTensors];   // 0 weight, 1 gradient
  long long numel[kMaxTensors];
  int n_tensors;
};

__device__ __forceinline__ double block_sum_to_double(float v, float* smem) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) smem[w] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kSqThreads / 32; ++i) r += (double)smem[i];
  }
  __syncthreads();
  return r;  // valid in thread 0
}

// Reduction tree (fixed order at every level, so the result is bit-reproducible):
//   chunk   : fp32 in the lane, fixed shuffle tree                               (scan kernel, one warp per chunk)
//   tensor  : lane l adds chunks l, l + 32, ... in fp64, then a fixed shuffle tree (fold kernel, one warp per tensor)
//   bucket  : lane l adds tensors l, l + 32, ... of the bucket, shuffle tree       (fold kernel, one warp per output)
// Two launches, the second one a single 1024-thread block started with programmatic dependent launch so that its
// launch latency hides behind the scan.  What was measured on the way here (profiles/r2_sqnorm.md):
//   * one "last block" folding all tensors after the scan: the SMs were active 64 k of the launch's 132 k cycles;
//   * per-tensor arrival counters (fold by the warp that completes a tensor): a fence + atomic round trip per
//     chunk, 50 us under ncu, and the counters cost a memset launch per call;
//   * ticket scheduling of chunks: slower than a static stride (one more round trip per chunk).
template <int kChunk>
__global__ void __launch_bounds__(kSqThreads)
    sqnorm_scan_kernel(const __grid_constant__ SqnormTable tab, double* __restrict__ partial) {
  // One WARP per chunk, no block barrier: 8 independent 128-bit loads per lane.  The host sizes the grid so that
  // every warp gets the same number of chunks (+-1): with ~1.2 chunks per warp a third of the launch was a tail.
  const int n_chunks = tab.chunk_start[tab.n_tensors];
  const int lane = threadIdx.x & 31;
  const int gwarp = blockIdx.x * (kSqThreads / 32) + (threadIdx.x >> 5), nwarps = gridDim.x * (kSqThreads / 32);
  for (int chunk = gwarp; chunk < n_chunks; chunk += nwarps) {
    // binary search: largest t with chunk_start[t] <= chunk
    int lo = 0, hi = tab.n_tensors;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (tab.chunk_start[mid] <= chunk) lo = mid; else hi = mid;
    }
    const long long off = (long long)(chunk - tab.chunk_start[lo]) * kChunk;
    const long long rem = tab.numel[lo] - off;
    const int cnt = rem < kChunk ? (int)rem : kChunk;
    const float* p = tab.ptr[lo] + off;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
      const float4* p4 = reinterpret_cast<const float4*>(p);
      const int n4 = cnt >> 2;
      for (int j0 = lane; j0 < n4; j0 += 8 * 32) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = j0 + u * 32;
          v[u] = j < n4 ? ldg_stream(p4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          a0 = fmaf(v[u].x, v[u].x, a0); a1 = fmaf(v[u].y, v[u].y, a1);
          a2 = fmaf(v[u].z, v[u].z, a2); a3 = fmaf(v[u].w, v[u].w, a3);
        }
      }
      for (int j = (n4 << 2) + lane; j < cnt; j += 32) a0 = fmaf(p[j], p[j], a0);
    } else {
      for (int j = lane; j < cnt; j += 32) {
        const float x = __ldg(p + j);
        a0 = fmaf(x, x, a0);
      }
    }
    const float s = warp_sum((a0 + a1) + (a2 + a3));  // fp32 inside the chunk, fixed shuffle tree
    if (lane == 0) partial[chunk] = (double)s;
  }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

constexpr int kFoldThreads = 1024;
constexpr int kFoldCluster = 8;  // one cluster of 8 x 32 warps: every tensor of the real model gets its own warp

__global__ void __launch_bounds__(kFoldThreads)
    sqnorm_fold_kernel(const __grid_constant__ SqnormTable tab, const double* __restrict__ partial,
                       double* __restrict__ out8, double* __restrict__ per_tensor, int accumulate_out) {
  // A single block walking ~9 tensors per warp one after the other took 18 us (r2_sqnorm3.ncu-rep): each tensor is a
  // chain of constant-bank misses, L2 loads and five fp64 shuffles.  A cluster gives 256 warps and a hardware barrier;
  // the per-tensor sums and bucket codes travel to CTA 0 through distributed shared memory, so the last stage reads
  // shared memory only (indexing the constant bank per lane would serialise).
  namespace cg = cooperative_groups;
  __shared__ double tsum[kMaxTensors];
  __shared__ unsigned char code[kMaxTensors];
  cg::cluster_group cluster = cg::this_cluster();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rank = (int)cluster.block_rank(), wpc = kFoldThreads / 32;
  double* tsum0 = cluster.map_shared_rank(tsum, 0);
  unsigned char* code0 = cluster.map_shared_rank(code, 0);
  // arrive now, wait just before the first remote write: CTA 0 must be resident before anyone writes into its
  // shared memory, but nobody has to wait for that while folding
  auto token = cluster.barrier_arrive();
  bool synced = false;
  asm volatile("griddepcontrol.wait;" ::: "memory");  // the scan's partials are complete and visible after this
  for (int t = rank * wpc + warp; t < tab.n_tensors; t += (int)cluster.num_blocks() * wpc) {
    const int c0 = tab.chunk_start[t], c1 = tab.chunk_start[t + 1];
    const unsigned char cd = (unsigned char)(tab.kind[t] * 16 + tab.mask[t]);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int c = c0 + lane;
    for (; c + 224 < c1; c += 256) {  // eight loads in flight; the order of the additions does not depend on it
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcg(partial + c + 32 * u);
      s0 += v[0]; s1 += v[1]; s2 += v[2]; s3 += v[3];
      s0 += v[4]; s1 += v[5]; s2 += v[6]; s3 += v[7];
    }
    for (; c + 96 < c1; c += 128) {
      const double v0 = __ldcg(partial + c), v1 = __ldcg(partial + c + 32), v2 = __ldcg(partial + c + 64),
                   v3 = __ldcg(partial + c + 96);
      s0 += v0; s1 += v1; s2 += v2; s3 += v3;
    }
    for (; c < c1; c += 32) s0 += __ldcg(partial + c);
    double s = (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (!synced) {
      cluster.barrier_wait(std::move(token));
      synced = true;
    }
    if (lane == 0) {
      tsum0[t] = s;
      code0[t] = cd;
      if (per_tensor) per_tensor[t] = s;
    }
  }
  if (!synced) cluster.barrier_wait(std::move(token));
  cluster.sync();
  if (rank == 0 && warp < 8) {  // out8 index: kind * 4 + bucket bit
    const int bit = 1 << (warp & 3), kind = warp >> 2;
    double s = 0.0;
    for (int t = lane; t < tab.n_tensors; t += 32) {
      const int cd = code[t];
      if ((cd >> 4) == kind && (cd & bit)) s += tsum[t];
    }
#pragma unroll
    for (int w = 16; w > 0; w >>= 1) s += __shfl_xor_sync(0xffffffffu, s, w);
    if (lane == 0) out8[warp] = (accumulate_out ? out8[warp] : 0.0) + s;
  }
}

__global__ void __launch_bounds__(256)
    squeeze_accumulate_kernel(const float* __restrict__ s, const uint8_t* __restrict__ select, int n, int c,
                              double* sum, long long* count) {
  // one thread per channel, rows in order: fixed summation order, coalesced across channels
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < c) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i)
      if (!select || select[i]) acc += (double)s[(size_t)i * c + j];
    sum[j] += acc;
  }
  if (j == 0 && count) {
    long long k = 0;
    for (int i = 0; i < n; ++i) k += (!select || select[i]) ? 1 : 0;
    *count += k;
  }
}

__global__ void __launch_bounds__(256)
    accuracy_counts_kernel(const float* __restrict__ l0, const float* __restrict__ l1,
                           const long long* __restrict__ labels, int n, int k, int* counts3) {
  // single block: n is a batch size; one thread per sample, argmax = first maximal index
  __shared__ int cnt[3];
  if (threadIdx.x < 3) cnt[threadIdx.x] = 0;
  __syncthreads();
  int c0 = 0, c1 = 0, cf = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float* a = l0 + (size_t)i * k;
    const float* b = l1 + (size_t)i * k;
    int i0 = 0, i1 = 0, if_ = 0;
    float m0 = a[0], m1 = b[0], mf = (a[0] + b[0]) / 2.f;  // model.py:108 (x_0 + x_1) / 2
    for (int j = 1; j < k; ++j) {
      const float f = (a[j] + b[j]) / 2.f;
      if (a[j] > m0) { m0 = a[j]; i0 = j; }
      if (b[j] > m1) { m1 = b[j]; i1 = j; }
      if (f > mf) { mf = f; if_ = j; }
    }
    const long long y = (n == 2) ? labels[0] : labels[i];  // train.py:36-37 batch-size-2 quirk
    cf += (if_ == y); c0 += (i0 == y); c1 += (i1 == y);
  }
  atomicAdd(&cnt[0], cf); atomicAdd(&cnt[1], c0); atomicAdd(&cnt[2], c1);  // integer: order-free
  __syncthreads();
  if (threadIdx.x < 3) counts3[threadIdx.x] = cnt[threadIdx.x];
}

inline long long chunks_of(long long numel, int chunk) { return (numel + chunk - 1) / chunk; }

}  // namespace

}  // namespace gml

using namespace gml;

extern "C" size_t gml_sqnorm_workspace_bytes(const int64_t* numel_host, int32_t n_tensors) {
  if (!numel_host || n_tensors <= 0) return 0;
  long long chunks = 0;
  for (int i = 0; i < n_tensors; ++i) chunks += chunks_of(numel_host[i], kMinChunk);
  return 256 + (size_t)chunks * sizeof(double);  // per-chunk partials (sized for the smallest chunk a variant uses)
}

extern "C" int gml_multi_tensor_sqnorm(const void* const* tensors_host, const int64_t* numel_host,
                                       const int32_t* bucket_mask_host, const int32_t* kind_host, int32_t n_tensors,
                                       double* out8, double* per_tensor, void* workspace, size_t workspace_bytes,
                                       void* stream) {
  if (!tensors_host || !numel_host || !bucket_mask_host || !kind_host || !out8 || !workspace || n_tensors <= 0)
    return GML_E_BADARG;
  if (workspace_bytes < gml_sqnorm_workspace_bytes(numel_host, n_tensors)) return GML_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* partial = reinterpret_cast<double*>(static_cast<char*>(workspace) + 256);
  static thread_local SqnormTable tab;  // 14 KB: keep it off the stack
  static const int kChunkOf[4] = {4096, 2048, 1024, 8192};
  const int chunk = kChunkOf[g_sq_variant & 3];
  const bool pdl = !(g_sq_variant & 4);
  const int bps = (g_sq_variant >> 3) & 15;  // blocks per SM (0 = 8)
  for (int base = 0; base < n_tensors; base += kMaxTensors) {
    const int cnt = (n_tensors - base) < kMaxTensors ? (n_tensors - base) : kMaxTensors;
    long long chunks = 0;
    for (int i = 0; i < cnt; ++i) {
      const int64_t ne = numel_host[base + i];
      if (ne < 0 || (ne > 0 && !tensors_host[base + i])) return GML_E_BADARG;
      if (kind_host[base + i] != 0 && kind_host[base + i] != 1) return GML_E_BADARG;
      tab.ptr[i] = static_cast<const float*>(tensors_host[base + i]);
      tab.numel[i] = ne;
      tab.mask[i] = (unsigned char)(bucket_mask_host[base + i] & 15);
      tab.kind[i] = (unsigned char)kind_host[base + i];
      tab.chunk_start[i] = (int)chunks;
      chunks += chunks_of(ne, chunk);
      if (chunks > 0x7fffffffLL) return GML_E_UNSUPPORTED;
    }
    tab.chunk_start[cnt] = (int)chunks;
    tab.n_tensors = cnt;
    // every warp the same number of chunks: rounds = ceil(chunks / resident warps), warps = ceil(chunks / rounds)
    const int wpb = kSqThreads / 32;
    const long long max_warps = (long long)kNumSMs * (bps ? bps : 8) * wpb;
    const long long rounds = chunks > 0 ? (chunks + max_warps - 1) / max_warps : 1;
    const long long warps = chunks > 0 ? (chunks + rounds - 1) / rounds : 1;
    const int grid = (int)((warps + wpb - 1) / wpb);
    double* pt = per_tensor ? per_tensor + base : nullptr;
    const int acc = base > 0 ? 1 : 0;
    {
      LaunchScope ls(kTagSqnorm, st);
      switch (g_sq_variant & 3) {
        case 0: sqnorm_scan_kernel<4096><<<grid, kSqThreads, 0, st>>>(tab, partial); break;
        case 1: sqnorm_scan_kernel<2048><<<grid, kSqThreads, 0, st>>>(tab, partial); break;
        case 2: sqnorm_scan_kernel<1024><<<grid, kSqThreads, 0, st>>>(tab, partial); break;
        default: sqnorm_scan_kernel<8192><<<grid, kSqThreads, 0, st>>>(tab, partial); break;
      }
      GML_LAUNCH_CHECK();
    }
    {
      LaunchScope ls(kTagSqnorm, st);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(kFoldCluster);
      cfg.blockDim = dim3(kFoldThreads);
      cfg.stream = st;
      cudaLaunchAttribute attr[2];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = kFoldCluster;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[1].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = pdl ? 2 : 1;
      GML_CUDA_TRY(cudaLaunchKernelEx(&cfg, sqnorm_fold_kernel, tab, (const double*)partial, out8, pt, acc));
    }
  }
  return GML_OK;
}

extern "C" int gml_squeeze_accumulate(const float* s, const uint8_t* select, int32_t n, int32_t c, double* sum,
                                      int64_t* count, void* stream) {
  if (!s || !sum || n < 0 || c <= 0) return GML_E_BADARG;
  if (n == 0) return GML_OK;
  {
    LaunchScope ls(kTagStats, static_cast<cudaStream_t>(stream));
    squeeze_accumulate_kernel<<<ceil_div(c, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        s, select, n, c, sum, reinterpret_cast<long long*>(count));
  }
  GML_LAUNCH_CHECK();
  return GML_OK;
}

extern "C" int gml_accuracy_counts(const float* logits0, const float* logits1, const int64_t* labels, int32_t n,
                                   int32_t k, int32_t* counts3, void* stream) {
  if (!logits0 || !logits1 || !labels || !counts3 || n <= 0 || k <= 0) return GML_E_BADARG;
  {
    LaunchScope ls(kTagStats, static_cast<cudaStream_t>(stream));
    accuracy_counts_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        logits0, logits1, reinterpret_cast<const long long*>(labels), n, k, counts3);
  }
  GML_LAUNCH_CHECK();
  return GML_OK;
}
