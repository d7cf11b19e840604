// Analysis statistics of the hot path:
//   * multi-tensor sum of squares for the conditional learning speed (one launch for all
//     parameters and gradients; replaces src/callbacks.py:203-205 -- 2 reductions and 2 .item()
//     host syncs PER TENSOR in the reference),
//   * selected-row squeeze accumulation for the conditional utilization rate
//     (src/balanced_mmtm.py:186-201),
//   * argmax/equality accuracy counts (train.py:32-40).
// All three are HBM-bound scans; no tensor cores, no atomics on the data path, fixed reduction
// order (bit-reproducible).
#include "common.cuh"

namespace gml {

namespace {

constexpr int kMaxTensors = 1024;  // per launch; 14.3 KB of kernel parameters (CUDA >= 12.1 allows 32 KB)
constexpr int kChunk = 4096;       // elements per chunk = 256 threads x 4 x float4
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

__global__ void __launch_bounds__(kSqThreads)
    sqnorm_kernel(const __grid_constant__ SqnormTable tab, double* __restrict__ partial, unsigned int* counter,
                  double* __restrict__ out8, double* __restrict__ per_tensor, int accumulate_out) {
  __shared__ bool is_last;
  // One WARP per chunk, no block barrier in the scan: 64 warps per SM with 8 independent 128-bit loads per lane keep
  // ~256 KB in flight per SM, enough to cover HBM latency at full bandwidth.
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
  // last block to finish folds the partials in a fixed order
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(counter, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // one warp per tensor: lane l adds chunks l, l + 32, ... (four loads in flight), then a fixed shuffle tree --
  // a fixed order, so the result is bit-reproducible; the big convolution weights have ~600 chunks each and a
  // single thread walking them was the tail of the whole launch
  __shared__ double tsum[kMaxTensors];
  const int warp = threadIdx.x >> 5;
  for (int t = warp; t < tab.n_tensors; t += kSqThreads / 32) {
    const int c0 = tab.chunk_start[t], c1 = tab.chunk_start[t + 1];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int c = c0 + lane;
    for (; c + 96 < c1; c += 128) {
      const double v0 = __ldcg(partial + c), v1 = __ldcg(partial + c + 32), v2 = __ldcg(partial + c + 64),
                   v3 = __ldcg(partial + c + 96);
      s0 += v0; s1 += v1; s2 += v2; s3 += v3;
    }
    for (; c < c1; c += 32) s0 += __ldcg(partial + c);
    double s = (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      tsum[t] = s;
      if (per_tensor) per_tensor[t] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    const int bit = 1 << (threadIdx.x & 3);
    const int kind = threadIdx.x >> 2;
    double s = accumulate_out ? out8[threadIdx.x] : 0.0;
    for (int t = 0; t < tab.n_tensors; ++t)
      if (tab.kind[t] == kind && (tab.mask[t] & bit)) s += tsum[t];
    out8[threadIdx.x] = s;
  }
  if (threadIdx.x == 0) *counter = 0u;  // workspace is reusable without a memset
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

inline long long chunks_of(long long numel) { return (numel + kChunk - 1) / kChunk; }

}  // namespace

}  // namespace gml

using namespace gml;

extern "C" size_t gml_sqnorm_workspace_bytes(const int64_t* numel_host, int32_t n_tensors) {
  if (!numel_host || n_tensors <= 0) return 0;
  long long chunks = 0;
  for (int i = 0; i < n_tensors; ++i) chunks += chunks_of(numel_host[i]);
  return 256 + (size_t)chunks * sizeof(double);  // [counter | pad][partials]
}

extern "C" int gml_multi_tensor_sqnorm(const void* const* tensors_host, const int64_t* numel_host,
                                       const int32_t* bucket_mask_host, const int32_t* kind_host, int32_t n_tensors,
                                       double* out8, double* per_tensor, void* workspace, size_t workspace_bytes,
                                       void* stream) {
  if (!tensors_host || !numel_host || !bucket_mask_host || !kind_host || !out8 || !workspace || n_tensors <= 0)
    return GML_E_BADARG;
  if (workspace_bytes < gml_sqnorm_workspace_bytes(numel_host, n_tensors)) return GML_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned int* counter = static_cast<unsigned int*>(workspace);
  double* partial = reinterpret_cast<double*>(static_cast<char*>(workspace) + 256);
  // The counter must start at zero: the kernel restores it, so zero it only on first use is not
  // knowable here -> a 4-byte async memset per call (stream ordered, negligible).
  GML_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
  static thread_local SqnormTable tab;  // 14 KB: keep it off the stack
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
      chunks += chunks_of(ne);
      if (chunks > 0x7fffffffLL) return GML_E_UNSUPPORTED;
    }
    tab.chunk_start[cnt] = (int)chunks;
    tab.n_tensors = cnt;
    int grid = kNumSMs * 8;
    if (chunks < grid) grid = chunks > 0 ? (int)chunks : 1;
    LaunchScope ls(kTagSqnorm, st);
    sqnorm_kernel<<<grid, kSqThreads, 0, st>>>(tab, partial, counter, out8, per_tensor ? per_tensor + base : nullptr,
                                               base > 0 ? 1 : 0);
    GML_LAUNCH_CHECK();
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
