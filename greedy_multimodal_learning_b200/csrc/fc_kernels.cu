// Small dense kernels for the squeeze/excitation FCs of the MMTM block and their gradients.
//
// fp32 on CUDA cores by design: the contract is 1e-5 relative parity with the reference's fp32
// Linear layers (BASELINE.json north_star), the problems are tiny next to the feature-map
// traffic, and they are not worth reshaping for tensor cores.  One register-tiled SGEMM covers
// every layout the forward/backward needs (see GemmDesc in kernels.h); reductions run in a fixed
// order, so results are bit-reproducible.
#include "common.cuh"
#include "kernels.h"

namespace gml {

namespace {

constexpr int BK = 16;

struct GemmBatch {
  GemmDesc d[2];
};

__device__ __forceinline__ float4 load4_guard(const float* base, size_t idx, int valid, bool vec) {
  // valid = number of in-range elements starting at idx (<= 0: none)
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid >= 4 && vec) {
    r = *reinterpret_cast<const float4*>(base + idx);
  } else if (valid > 0) {
    r.x = base[idx];
    if (valid > 1) r.y = base[idx + 1];
    if (valid > 2) r.z = base[idx + 2];
    if (valid > 3) r.w = base[idx + 3];
  }
  return r;
}

// One operand tile [BK x BT] (k-major in shared memory) from either layout.
template <int BT, int T>
struct TileLoader {
  static constexpr int kVecs = BT * BK / 4;          // float4 per tile
  static constexpr int kPerThread = (kVecs + T - 1) / T;
  float4 v[kPerThread];

  __device__ __forceinline__ void load(const float* src, int ld, bool kc, bool vec, int t0, int k0, int tmax,
                                       int kmax, int tid) {
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
      const int f = tid + i * T;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < kVecs) {
        if (kc) {  // src[t * ld + k], 4 consecutive k
          const int t = t0 + f / (BK / 4), k = k0 + (f % (BK / 4)) * 4;
          if (t < tmax) v[i] = load4_guard(src, (size_t)t * ld + k, kmax - k, vec);
        } else {   // src[k * ld + t], 4 consecutive t
          const int k = k0 + f / (BT / 4), t = t0 + (f % (BT / 4)) * 4;
          if (k < kmax) v[i] = load4_guard(src, (size_t)k * ld + t, tmax - t, vec);
        }
      }
    }
  }
  __device__ __forceinline__ void store(float (*s)[BT + 4], bool kc, int tid) const {
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
      const int f = tid + i * T;
      if (f < kVecs) {
        if (kc) {
          const int t = f / (BK / 4), kq = (f % (BK / 4)) * 4;
          s[kq + 0][t] = v[i].x; s[kq + 1][t] = v[i].y; s[kq + 2][t] = v[i].z; s[kq + 3][t] = v[i].w;
        } else {
          const int k = f / (BT / 4), tq = (f % (BT / 4)) * 4;
          *reinterpret_cast<float4*>(&s[k][tq]) = v[i];
        }
      }
    }
  }
};

template <int BM, int BN>
__global__ void __launch_bounds__((BM / 4) * (BN / 4)) gemm_kernel(GemmBatch batch) {
  constexpr int T = (BM / 4) * (BN / 4);
  const GemmDesc d = blockIdx.z ? batch.d[1] : batch.d[0];  // field-wise select, no local copy
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (m0 >= d.m || n0 >= d.n) return;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / 4), ty = tid / (BN / 4);
  const bool a_vec = (d.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(d.a) & 15u) == 0);
  const bool b_vec = (d.ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(d.b) & 15u) == 0);
  TileLoader<BM, T> la;
  TileLoader<BN, T> lb;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = (d.k + BK - 1) / BK;
  la.load(d.a, d.lda, d.a_kc, a_vec, m0, 0, d.m, d.k, tid);
  lb.load(d.b, d.ldb, d.b_kc, b_vec, n0, 0, d.n, d.k, tid);
  la.store(As[0], d.a_kc, tid);
  lb.store(Bs[0], d.b_kc, tid);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) {
      la.load(d.a, d.lda, d.a_kc, a_vec, m0, (kt + 1) * BK, d.m, d.k, tid);
      lb.load(d.b, d.ldb, d.b_kc, b_vec, n0, (kt + 1) * BK, d.n, d.k, tid);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w};
      const float b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      la.store(As[buf ^ 1], d.a_kc, tid);
      lb.store(Bs[buf ^ 1], d.b_kc, tid);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= d.m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= d.n) continue;
      float v = acc[i][j];
      float* cp = d.c + (size_t)m * d.ldc + n;
      if (d.beta) v += *cp;
      if (d.bias) v += d.bias[n];
      if (d.act == kActRelu) v = fmaxf(v, 0.f);
      else if (d.act == kActSigmoid) v = sigmoidf_ref(v);
      else if (d.act == kActReluMask) v = d.mask[(size_t)m * d.ldmask + n] > 0.f ? v : 0.f;
      *cp = v;
    }
  }
}

struct ColsumBatch {
  ColsumSeg seg[4];
  int block_start[5];
};

// 32 columns x 32 row groups per block: a warp reads 128 contiguous bytes of one row, every thread keeps up to 16
// loads in flight (the kernel is pure L2/HBM latency: a [1024, 512] input is 2 MB).  Fixed summation order.
constexpr int COLSUM_THREADS = 1024;
__global__ void __launch_bounds__(COLSUM_THREADS) colsum_kernel(const ColsumBatch b) {
  __shared__ float part[32][33];
  int sid = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if ((int)blockIdx.x >= b.block_start[i]) sid = i;
  const ColsumSeg s = sid == 0 ? b.seg[0] : (sid == 1 ? b.seg[1] : (sid == 2 ? b.seg[2] : b.seg[3]));
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = (blockIdx.x - b.block_start[sid]) * 32 + tx;
  float acc = 0.f;
  if (j < s.cols) {
    const float* col = s.x + j;
    int i = ty;
    for (; i + 15 * 32 < s.rows; i += 16 * 32) {
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = col[(size_t)(i + 32 * u) * s.ld];
#pragma unroll
      for (int u = 0; u < 16; ++u) acc += v[u];
    }
    for (; i + 3 * 32 < s.rows; i += 4 * 32) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = col[(size_t)(i + 32 * u) * s.ld];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc += v[u];
    }
    for (; i < s.rows; i += 32) acc += col[(size_t)i * s.ld];
  }
  part[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && j < s.cols) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) t += part[r][tx];
    s.out[j] = t;
    if (s.run_v) {
      const float mean = t / s.n_total;
      s.run_v[j] = (mean + s.run_v[j] * s.step) / (s.step + 1.f);
      if (s.run_s) s.run_s[j] = (mean + s.run_s[j] * s.step) / (s.step + 1.f);
    }
  }
}

__global__ void fill_rows_kernel(float* z, int rows, int ld, int off, const float* __restrict__ v, int cols) {
  const int total = rows * cols;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / cols, j = i - r * cols;
    z[(size_t)r * ld + off + j] = v[j];
  }
}

// balanced_mmtm.py:113-114 -- run <- (mean_n g_a + run * step) / (step + 1), fp32 like the reference
__global__ void running_update_kernel(float* run_v, float* run_s, const float* __restrict__ gate_sum, int c,
                                      float n_total, float step) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= c) return;
  const float mean = gate_sum[j] / n_total;
  run_v[j] = (mean + run_v[j] * step) / (step + 1.f);
  if (run_s) run_s[j] = (mean + run_s[j] * step) / (step + 1.f);
}

__global__ void fill_zero_kernel(float* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = 0.f;
}

}  // namespace

int launch_gemm(const GemmDesc* descs, int count, cudaStream_t st, void* ws, size_t ws_bytes) {
  if (count < 1 || count > 3) return GML_E_BADARG;
  {
    const int rc = launch_gemm_pipelined(descs, count, ws, ws_bytes, st);
    if (rc != GML_E_UNSUPPORTED) return rc;
  }
  if (count == 3) {  // the generic kernel batches two problems
    const int rc = launch_gemm(descs, 2, st, ws, ws_bytes);
    return rc != GML_OK ? rc : launch_gemm(descs + 2, 1, st, ws, ws_bytes);
  }
  for (int i = 0; i < count; ++i)
    if (descs[i].k_split) return GML_E_UNSUPPORTED;  // the generic kernel knows one K segment only
  GemmBatch batch;
  int max_m = 0, max_n = 0;
  for (int i = 0; i < count; ++i) {
    batch.d[i] = descs[i];
    if (descs[i].m <= 0 || descs[i].n <= 0 || descs[i].k <= 0) return GML_E_BADARG;
    max_m = descs[i].m > max_m ? descs[i].m : max_m;
    max_n = descs[i].n > max_n ? descs[i].n : max_n;
  }
  if (count == 1) batch.d[1] = descs[0];
  const long tiles64 = (long)ceil_div(max_m, 64) * ceil_div(max_n, 64) * count;
  if (tiles64 >= 120) {
    dim3 grid(ceil_div(max_n, 64), ceil_div(max_m, 64), count);
    LaunchScope ls(kTagGemm, st);
    gemm_kernel<64, 64><<<grid, 256, 0, st>>>(batch);
  } else {
    dim3 grid(ceil_div(max_n, 32), ceil_div(max_m, 32), count);
    LaunchScope ls(kTagGemm, st);
    gemm_kernel<32, 32><<<grid, 64, 0, st>>>(batch);
  }
  GML_LAUNCH_CHECK();
  return GML_OK;
}

int launch_colsums(const ColsumSeg* segs, int count, cudaStream_t st) {
  if (count < 1 || count > 4) return GML_E_BADARG;
  ColsumBatch b;
  int blocks = 0;
  for (int i = 0; i < 4; ++i) {
    b.block_start[i] = blocks;
    if (i < count) {
      b.seg[i] = segs[i];
      blocks += ceil_div(segs[i].cols, 32);
    } else {
      b.seg[i] = segs[0];
    }
  }
  b.block_start[4] = blocks;
  for (int i = count; i < 4; ++i) b.block_start[i] = blocks;  // unreachable segments
  { LaunchScope ls(kTagSmall, st);
  colsum_kernel<<<blocks, COLSUM_THREADS, 0, st>>>(b); }
  GML_LAUNCH_CHECK();
  return GML_OK;
}

int launch_colsum(const float* x, int rows, int cols, int ld, float* out, cudaStream_t st) {
  ColsumSeg s{x, out, rows, cols, ld, nullptr, nullptr, 1.f, 0.f};
  return launch_colsums(&s, 1, st);
}

int launch_fill_rows(float* z, int rows, int ld, int off, const float* v, int cols, cudaStream_t st) {
  const int total = rows * cols;
  { LaunchScope ls(kTagSmall, st);
  fill_rows_kernel<<<ceil_div(total, 256) < 1024 ? ceil_div(total, 256) : 1024, 256, 0, st>>>(z, rows, ld, off, v,
                                                                                            cols); }
  GML_LAUNCH_CHECK();
  return GML_OK;
}

int launch_running_update(float* run_v, float* run_s, const float* gate_sum, int c, double n_total, double step,
                          cudaStream_t st) {
  { LaunchScope ls(kTagSmall, st);
  running_update_kernel<<<ceil_div(c, 128), 128, 0, st>>>(run_v, run_s, gate_sum, c, (float)n_total, (float)step); }
  GML_LAUNCH_CHECK();
  return GML_OK;
}

int launch_fill_zero(float* p, size_t n, cudaStream_t st) {
  if (n == 0) return GML_OK;
  size_t blocks = (n + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  { LaunchScope ls(kTagSmall, st);
  fill_zero_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, n); }
  GML_LAUNCH_CHECK();
  return GML_OK;
}

}  // namespace gml
