// Internal launch interfaces between the translation units of libgml_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gml {

// ---- streaming plane kernels (stream_kernels.cu) -----------------------------------------
constexpr int kEpiMean = 0;   // out = sum / HW
constexpr int kEpiDGate = 1;  // out = mul * sum * g * (1 - g)

struct ReduceSeg {
  const float* x;     // [rows, hw]
  const float* y;     // [rows, hw] second operand of the dot (backward) or nullptr
  float* out;         // out[n * out_ld + out_off + c]
  const float* gate;  // [rows] (kEpiDGate)
  int rows, hw, c;    // rows = N * c
  int out_ld, out_off;
  float mul;
};
struct ReduceLaunch {
  ReduceSeg seg[2];
  int seg_blocks0;
  int keep_in_l2;
};

struct ScaleSeg {
  const float* x;      // [rows, hw]
  float* out;          // [rows, hw]
  const float* scale;  // [rows] or [c] when scale_bcast
  const float* add;    // add[n * add_ld + add_off + c], divided by hw (or nullptr)
  int rows, hw, c;
  int scale_bcast;
  int add_ld, add_off;
  float mul;
};
struct ScaleLaunch {
  ScaleSeg seg[2];
  int seg_blocks0;
};

int launch_plane_mean(const ReduceSeg& a, const ReduceSeg& b, bool keep_in_l2, cudaStream_t st);
int launch_plane_dgate(const ReduceSeg& a, const ReduceSeg& b, bool keep_in_l2, cudaStream_t st);
int launch_plane_scale(const ScaleSeg& a, const ScaleSeg& b, bool has_add, cudaStream_t st);

// ---- small dense kernels (fc_kernels.cu) ---------------------------------------------------
constexpr int kActNone = 0;
constexpr int kActRelu = 1;      // C = max(acc + bias, 0)
constexpr int kActSigmoid = 2;   // C = sigmoid(acc + bias)
constexpr int kActReluMask = 3;  // C = (beta*C + acc) * [mask > 0]

// C[M,N] (ldc) = beta * C + sum_k A(m,k) * B(k,n), fp32 on CUDA cores.
//   a_kc: A(m,k) = a[m*lda + k]  else a[k*lda + m]
//   b_kc: B(k,n) = b[n*ldb + k]  else b[k*ldb + n]
struct GemmDesc {
  const float* a; const float* b; float* c;
  const float* bias;  // [N] or nullptr
  const float* mask;  // [M, ldmask] for kActReluMask
  int m, n, k;
  int lda, ldb, ldc, ldmask;
  int a_kc, b_kc;
  int act;
  int beta;           // 0 or 1
  // optional second K segment: for k >= k_split the operands come from a2 / b2 (addressed with k - k_split and
  // lda2 / ldb2), i.e. C = [A | A2] [B | B2]^T in one launch.  k_split == 0: none.  Only the pipelined kernels
  // support it (gemm_ksplit_ok); k_split must be a multiple of 32.
  const float* a2; const float* b2;
  int lda2, ldb2, k_split;
};
bool gemm_ksplit_ok(const GemmDesc& d);
// up to three independent problems in one launch (blockIdx.z; equal K and operand layouts).  With a workspace of
// gemm_workspace_bytes() the pipelined split-K kernel (gemm_kernels.cu) is used whenever the operands
// are 16-byte aligned; otherwise the generic register-staged kernel (fc_kernels.cu).
size_t gemm_workspace_bytes();
int prepare_gemm_workspace(void* ws, size_t ws_bytes, cudaStream_t st);
int launch_gemm_pipelined(const GemmDesc* descs, int count, void* ws, size_t ws_bytes, cudaStream_t st);
int launch_gemm(const GemmDesc* descs, int count, cudaStream_t st, void* ws = nullptr, size_t ws_bytes = 0);

// out[j] = sum_i x[i*ld + j], i < rows, j < cols  (fixed order); up to 4 independent sums per launch.
// A segment with run_v != nullptr also applies the running-mean update (balanced_mmtm.py:113-114) to
// run_v / run_s with its column sums (n_total, step as floats).
struct ColsumSeg {
  const float* x; float* out;
  int rows, cols, ld;
  float* run_v; float* run_s;
  float n_total, step;
};
int launch_colsums(const ColsumSeg* segs, int count, cudaStream_t st);
int launch_colsum(const float* x, int rows, int cols, int ld, float* out, cudaStream_t st);
// z rows for mode 3: z[n, off + j] = v[j]
int launch_fill_rows(float* z, int rows, int ld, int off, const float* v, int cols, cudaStream_t st);
int launch_running_update(float* run_v, float* run_s, const float* gate_sum, int c, double n_total, double step,
                          cudaStream_t st);
int launch_fill_zero(float* p, size_t n, cudaStream_t st);

// ---- fused cluster kernels (fused_kernels.cu) ------------------------------------------------
struct FusedFwdArgs {
  const float* a; const float* b; float* a_out; float* b_out;
  const float* w_sq; const float* b_sq; const float* w_v; const float* b_v; const float* w_s; const float* b_s;
  float* z; float* h; float* g_a; float* g_b;
  const float* run_v; const float* run_s;
  int n, c, hw, d, mode;
  float gate_scale;
};
struct FusedBwdArgs {
  const float* go_a; const float* go_b; const float* a; const float* b;
  const float* w_sq; const float* w_v; const float* w_s;
  const float* z; const float* h; const float* g_a; const float* g_b;
  const float* run_v; const float* run_s;
  float* d_a; float* d_b;
  float* de_a; float* de_b; float* dh;  // [N,c],[N,c],[N,d] for the weight-gradient GEMMs
  int n, c, hw, d, mode;
  float gate_scale;
};
bool fused_supported(int n, int c_v, int c_s, int hw_v, int hw_s, int d, int mode);
bool fused_fwd_preferred(int n, int c, int hw, int d);
int launch_fused_fwd(const FusedFwdArgs& args, cudaStream_t st);
int launch_fused_bwd(const FusedBwdArgs& args, cudaStream_t st);


// ---- tile pipeline (tile_kernels.cu): one persistent kernel per block and direction, normal mode -----------------
bool tile_supported(int n, int c_v, int c_s, int hw_v, int hw_s, int d, int mode);
bool tile_preferred(int n, int c, int hw, int d, bool bwd);
size_t tile_fwd_workspace_bytes(int n, int c, int hw, int d);
size_t tile_bwd_workspace_bytes(int n, int c, int hw, int d);
// gate_sum / run_v / run_s may be nullptr (no column sum / no running-mean update inside the kernel)
int launch_tile_fwd(const FusedFwdArgs& args, float* gate_sum, float* run_v, float* run_s, float step, void* ws,
                    size_t ws_bytes, cudaStream_t st);
// dz_flat: [2, N, C] scratch; d_b_*: bias gradients (column sums of dE_a, dE_b, dH) or nullptr
int launch_tile_bwd(const FusedBwdArgs& a, float* dz_flat, float* d_b_v, float* d_b_s, float* d_b_sq, float* d_w_v,
                    float* d_w_s, float* d_w_sq, bool* wgrad_done, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace gml
