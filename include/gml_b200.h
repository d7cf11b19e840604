/*
 * gml_b200.h -- C ABI of the B200-native (sm_100a) hot path of greedy_multimodal_learning.
 *
 * The reference (SebastianHafner/greedy_multimodal_learning) is pure Python and has no FFI
 * of its own; the boundary it exposes for this path is a Python surface.
 * Each entry point below names the reference code it replaces (file:line relative to the
 * reference root).  The Python mirror of that surface lives in
 * greedy_multimodal_learning_b200/ and binds these symbols with ctypes
 * (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory unless the name ends in
 *     `_host`.  The caller owns all buffers (PyTorch caching allocator in practice); the
 *     library allocates no device memory and keeps no DATA state between calls.  What it does
 *     keep, per process: the tunables of gml_set_tunable (plain process-wide settings: set
 *     them before launching work, not concurrently with it), the launch counters, one lazily
 *     created helper stream + two events per host thread and device, and cached
 *     occupancy / function-attribute queries (mutex-protected).
 *   - every call is asynchronous and stream-ordered on `stream` (a cudaStream_t passed as
 *     void*; NULL = legacy default stream).  Calls from several host threads are safe as long
 *     as nobody changes a tunable meanwhile.
 *   - return value: 0 on success, a negative GML_E_* code otherwise; never throws.
 *   - tensors are fp32, dense, row-major.  Feature maps are NCHW-contiguous, seen as
 *     [N, C, HW] (the reference's `.view(shape[:2] + (-1,))`, src/balanced_mmtm.py:96).
 *   - "visual" = modality/view 0 = A, "skeleton" = modality/view 1 = B (reference naming).
 */
#ifndef GML_B200_H_
#define GML_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GML_ABI_VERSION 1

/* error codes */
#define GML_OK 0
#define GML_E_BADARG (-1)     /* null pointer, non-positive size, unknown mode      */
#define GML_E_ALIGN (-2)      /* reserved: all paths accept any 4-byte aligned data */
#define GML_E_WORKSPACE (-3)  /* workspace_bytes smaller than gml_*_workspace_bytes */
#define GML_E_CUDA (-4)       /* a CUDA runtime call failed; see gml_last_cuda_error */
#define GML_E_UNSUPPORTED (-5)

/* MMTM forward modes (src/balanced_mmtm.py) */
#define GML_MODE_NORMAL 0          /* :93-111,128-133                                         */
#define GML_MODE_CURATE_VISUAL 1   /* :135-143  curation_mode, caring_modality == 0           */
#define GML_MODE_CURATE_SKELETON 2 /* :145-152  curation_mode, caring_modality == 1           */
#define GML_MODE_XMODAL_OFF 3      /* :72-91    turnoff_cross_modal_flow with dataset means   */

/* flags for gml_mmtm_fwd / gml_mmtm_bwd */
#define GML_F_NO_RUNNING_UPDATE 1u /* skip the running-mean update (data-parallel callers do
                                      it themselves after all-reducing gate_sum)             */
#define GML_F_FORCE_STREAMING 2u   /* never pick the cluster/shared-memory-resident kernels  */
#define GML_F_FORCE_FUSED 4u       /* cluster kernels or GML_E_UNSUPPORTED, never a fallback  */
#define GML_F_FORCE_TILE 8u        /* tile-pipeline kernel or GML_E_UNSUPPORTED               */

/* learning-speed buckets (src/callbacks.py:207-223); a tensor may feed several */
#define GML_BUCKET_MAIN0 1
#define GML_BUCKET_MAIN1 2
#define GML_BUCKET_BYPASS0 4
#define GML_BUCKET_BYPASS1 8

typedef struct gml_mmtm_dims {
  int32_t n;     /* batch */
  int32_t c_v;   /* dim_visual                                   */
  int32_t c_s;   /* dim_skeleton                                 */
  int32_t hw_v;  /* H*W of the visual map                        */
  int32_t hw_s;  /* H*W of the skeleton map                      */
  int32_t d;     /* hidden dim, int(2*(c_v+c_s)/ratio) (:25-26)   */
} gml_mmtm_dims;

int gml_abi_version(void);
const char* gml_error_string(int code);
/* text of the last CUDA error seen by this thread inside the library (for GML_E_CUDA) */
const char* gml_last_cuda_error(void);
/* 1 when the running device is compute capability 10.x (B200/B300), else 0; <0 on error */
int gml_device_is_blackwell(void);

/* Launch accounting (measurement support; bench.py's `gpu_launches` and roofline legs).
 * gml_launch_count(tag): kernels this library has launched in the process so far, for one
 * kernel class (0 <= tag < gml_kernel_tag_count()) or all of them (tag < 0).
 * With profiling enabled every launch is bracketed by CUDA events on its stream;
 * gml_profile_read drains them (synchronising on the recorded events) and reports the
 * accumulated device time per class.  Do not enable while capturing a CUDA graph. */
int64_t gml_launch_count(int tag);
int gml_kernel_tag_count(void);
const char* gml_kernel_tag_name(int tag);
void gml_profile_enable(int on);
void gml_profile_reset(void);
int gml_profile_read(int tag, double* total_ms, int64_t* launches);
/* Process-wide tuning knobs for measurement sweeps (defaults are what ships):
 *   "l2_chunk_mb"    bytes of feature map the streaming path pushes through both passes at once
 *   "fused_kind"     0 auto | 1 shared-memory-resident cluster kernels | 2 L2-resident cluster kernels
 *   "fused_cluster"  0 auto | 4 | 8 | 16 CTAs per cluster
 *   "fused_threads"  0 auto | 256 | 512
 *   "fused_occ", "fused_stash_kb", "fused_group_kb", "fused_weight_ratio_x100"
 *                    occupancy / stash / group size / path-selection threshold ("fused_prefetch" is accepted and ignored)
 *   "fused_wsmem"    -1 auto | bit 0 / bit 1: first / second FC's weight slice of a CTA lives in shared memory
 *   "fused_stash_kb" -1 auto (24 KB; none for a backward whose weight slices are in shared memory) | KB per CTA
 *   "fused_hw_special" 1 plane-size-specialised cluster kernels for 28 x 28 planes | 0 generic instantiation
 *   "gemm_umma"      1 tcgen05 (TMEM) 3xTF32 kernel for FC problems above ~1e8 MACs | 0 never
 *   "gemm_tf32x3"    1 mma.sync 3xTF32 for large problems the tcgen05 kernel does not take | 0 CUDA cores
 *   "gemm_big_tiles" 1 opt-in 128x128 CUDA-core tiles
 *   "tile_kind"      0 auto | 1 tile pipeline whenever the shape allows | 2 never
 *   "tile_lag", "tile_m", "tile_gemm_ctas", "tile_chunk_kb", "tile_min_mb"
 *                    pipeline depth in tiles / samples per tile / CTAs on the FC role / chunk size /
 *                    smallest feature map (MB per modality) that takes the tile pipeline automatically
 *   "tile_wgrad"     1 weight-gradient GEMMs inside the backward pipeline launch | 0 separate launch
 *   "tile_light_fwd" 1 forward of blocks with < 1 MB of FC weights on the pipeline from "tile_min_mb_light" on | 0 never
 *   "overlap_wgrad"  1 weight-gradient GEMMs on the library's side stream (streaming backward)
 *   "fused_trace_ptr", "fused_occ_trace_ptr", "tile_stats_ptr", "gemm_trace_ptr"
 *                    device buffers for timing stamps; the first three only act in a tracing build
 *                    (make EXTRA="-DGML_L2_TRACE -DGML_TILE_TRACE")                                  */
int gml_set_tunable(const char* name, int64_t value);

/* ---------------------------------------------------------------------------------------
 * MMTM forward.  Replaces MMTM_mitigate.forward, src/balanced_mmtm.py:49-154.
 *
 *   s_a = mean_hw A, s_b = mean_hw B                                   (:94-97)
 *   H   = relu([s_a | s_b] Wsq^T + bsq)                                 (:99-101)
 *   g_a = sigmoid(H Wv^T + bv), g_b = sigmoid(H Ws^T + bs)              (:103-111)
 *   run_v, run_s <- (mean_n g_a + run * step) / (step + 1)   (both from g_a, sic; :113-114)
 *   A' = A * scale_a, B' = B * scale_b with scale = gate_scale * (g, or run_* on the
 *   substituted side in the curation modes :135-152).  gate_scale = 1 is the reference.
 *   Mode 3 computes g_a from [s_a | m_b] and g_b from [m_a | s_b] (:72-91); H then holds
 *   two [N, D] planes.
 *
 * Outputs kept for backward / export:
 *   z   [ZR, c_v+c_s]  the squeeze vectors as the FC consumes them, visual first.  ZR = N and
 *                      z[n] = [s_a[n] | s_b[n]] (what `return_squeezed_mps` exports, :123-124);
 *                      in mode 3 ZR = 2N, rows [0,N) = [s_a | m_b], rows [N,2N) = [m_a | s_b].
 *   h   [ZR, D]        post-ReLU hidden state (mode 3: rows [0,N) feed g_a, rows [N,2N) g_b)
 *   g_a [N,c_v], g_b [N,c_s]  live sigmoid gates (what `return_scale` exports, :118-121).
 * gate_sum [c_v] receives sum_n g_a (per-rank partial for the data-parallel running mean).
 * run_v/run_s [c_v] are updated in place unless GML_F_NO_RUNNING_UPDATE; `step` is the
 * value BEFORE this call (the caller increments its own counter, :116).
 * m_a [c_v], m_b [c_s] are only read in mode 3 (may be NULL otherwise).
 * a_out/b_out must not alias a/b (inputs are read through the read-only data path).
 * workspace: gml_mmtm_fwd_workspace_bytes(dims) bytes, 256-byte aligned.
 */
size_t gml_mmtm_fwd_workspace_bytes(const gml_mmtm_dims* dims);
int gml_mmtm_fwd(const float* a, const float* b, float* a_out, float* b_out,
                 const float* w_sq, const float* b_sq, const float* w_v, const float* b_v,
                 const float* w_s, const float* b_s,
                 float* z, float* h, float* g_a, float* g_b, float* gate_sum,
                 float* run_v, float* run_s, int64_t step,
                 const float* m_a, const float* m_b,
                 void* workspace, size_t workspace_bytes,
                 const gml_mmtm_dims* dims, int mode, float gate_scale, uint32_t flags, void* stream);

/* The same computation split at the point where a data-parallel caller must all-reduce
 * gate_sum before the substituted gate is known (curation modes):
 *   gml_mmtm_gates   : squeeze + excitation -> z, h, g_a, g_b, gate_sum
 *   gml_mmtm_running : run <- (gate_sum / n_total + run * step) / (step + 1)     (:113-114)
 *   gml_mmtm_apply   : A' = A * scale, B' = B * scale                            (:128-154)
 */
int gml_mmtm_gates(const float* a, const float* b,
                   const float* w_sq, const float* b_sq, const float* w_v, const float* b_v,
                   const float* w_s, const float* b_s,
                   float* z, float* h, float* g_a, float* g_b, float* gate_sum,
                   const float* m_a, const float* m_b,
                   void* workspace, size_t workspace_bytes,
                   const gml_mmtm_dims* dims, int mode, void* stream);
int gml_mmtm_running(float* run_v, float* run_s, const float* gate_sum, int32_t c_v,
                     int64_t n_total, int64_t step, void* stream);
int gml_mmtm_apply(const float* a, const float* b, float* a_out, float* b_out,
                   const float* g_a, const float* g_b, const float* run_v, const float* run_s,
                   const gml_mmtm_dims* dims, int mode, float gate_scale, void* stream);

/* ---------------------------------------------------------------------------------------
 * MMTM backward.  Replaces what autograd derives for src/balanced_mmtm.py:93-154
 * (MulBackward0, MeanBackward1, AddmmBackward0 x3, Sigmoid/ReluBackward).
 *
 *   dg   = gate_scale * sum_hw grad_out * input            (live sides only)
 *   dE   = dg * g * (1 - g)
 *   dH   = (dE_a Wv + dE_b Ws) * [H > 0]
 *   dZ   = dH Wsq ; ds_a = dZ[:, :c_v], ds_b = dZ[:, c_v:]
 *   dA   = grad_a_out * scale_a + ds_a / hw_v   (same for B)
 *   dWv  = dE_a^T H, dbv = sum_n dE_a, dWs = dE_b^T H, dbs = sum_n dE_b,
 *   dWsq = dH^T Z, dbsq = sum_rows dH
 * Weight gradients are OVERWRITTEN (the caller's autograd accumulates).  On a substituted
 * side (curation) the excitation FC receives no gradient: its dW/db, if requested, are
 * zero-filled (the reference leaves .grad = None there; the Python mirror passes NULL and
 * returns None for them).
 * run_v/run_s are the values the forward USED (pass a saved copy).  d_a/d_b must not alias
 * grad_a_out/grad_b_out.  Any of the six weight-gradient pointers may be NULL to skip it.
 * z, h, g_a, g_b are the buffers the forward of the SAME mode produced.
 */
size_t gml_mmtm_bwd_workspace_bytes(const gml_mmtm_dims* dims);
int gml_mmtm_bwd(const float* grad_a_out, const float* grad_b_out, const float* a, const float* b,
                 const float* w_sq, const float* w_v, const float* w_s,
                 const float* z, const float* h, const float* g_a, const float* g_b,
                 const float* run_v, const float* run_s, const float* m_a, const float* m_b,
                 float* d_a, float* d_b,
                 float* d_w_sq, float* d_b_sq, float* d_w_v, float* d_b_v, float* d_w_s, float* d_b_s,
                 void* workspace, size_t workspace_bytes,
                 const gml_mmtm_dims* dims, int mode, float gate_scale, uint32_t flags, void* stream);

/* ---------------------------------------------------------------------------------------
 * Conditional learning speed.  Replaces the 2 x n_tensors reductions + .item() syncs of
 * Bias_Mitigation_Strong.compute_BDR, src/callbacks.py:203-223, with ONE launch.
 *
 * tensors_host[i] / numel_host[i] / bucket_mask_host[i] describe n_tensors fp32 device
 * arrays (parameters and their gradients are simply listed as separate entries with
 * kind_host[i] = 0 for a weight, 1 for a gradient).  Host arrays are consumed before the
 * call returns (they travel as kernel parameters).  Two launches per 1024 tensors: a scan
 * (one warp per 4096-element chunk) and a one-cluster fold started with programmatic
 * dependent launch; no host synchronisation, no memset, graph-capturable.
 * out [8] doubles (device): {wn_main0, wn_main1, wn_bypass0, wn_bypass1,
 *                           gn_main0, gn_main1, gn_bypass0, gn_bypass1}.
 * per_tensor (device, n_tensors doubles) optionally receives each tensor's own sum.
 * Squares are taken in fp32, summed pairwise in fp32 inside a 4096-element chunk and in
 * fp64 across chunks, in a fixed order: results are bit-reproducible run to run.
 * workspace: gml_sqnorm_workspace_bytes(...) bytes.
 */
size_t gml_sqnorm_workspace_bytes(const int64_t* numel_host, int32_t n_tensors);
int gml_multi_tensor_sqnorm(const void* const* tensors_host, const int64_t* numel_host,
                            const int32_t* bucket_mask_host, const int32_t* kind_host,
                            int32_t n_tensors, double* out8, double* per_tensor,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Conditional utilization inputs.  On-device replacement for the pickle round trip of
 * get_mmtm_outputs + get_rescale_weights, src/balanced_mmtm.py:157-206: accumulate
 * sum over the selected samples of the recorded squeezes.
 *   sum [C] (double) += sum_{n : select[n] != 0} s[n, :];  count (int64) += #selected
 * select may be NULL (all rows).
 */
int gml_squeeze_accumulate(const float* s, const uint8_t* select, int32_t n, int32_t c,
                           double* sum, int64_t* count, void* stream);

/* ---------------------------------------------------------------------------------------
 * Accuracy counts.  Replaces the three `acc` calls + float() syncs per batch,
 * train.py:32-40 via src/framework.py:154-156,171-175.
 *   logits0, logits1 [N, K]; fused logits are (l0 + l1) / 2 (src/model.py:108).
 *   counts [3] int32 (device): #correct for {fused, view 0, view 1}.
 * argmax takes the first maximal index (torch .max(1)); with n == 2 every prediction is
 * compared with labels[0] (the reference's batch-size-2 quirk, train.py:36-37).
 */
int gml_accuracy_counts(const float* logits0, const float* logits1, const int64_t* labels,
                        int32_t n, int32_t k, int32_t* counts3, void* stream);

/* ---------------------------------------------------------------------------------------
 * The FC building block on its own (what torch.nn.Linear / AddmmBackward0 do inside
 * src/balanced_mmtm.py:101-108 and its autograd graph); exported so that the GEMM kernels
 * (CUDA-core, mma.sync 3xTF32, tcgen05 3xTF32 -- see gml_set_tunable) can be tested and timed
 * in isolation:
 *     C[i, j] = act(beta * C[i, j] + sum_k A(i, k) * B(j, k) + bias[j])
 * A(i, k) = a[i * lda + k] when a_kc != 0, else a[k * lda + i]; same for B with b_kc / ldb.
 * act: 0 none, 1 ReLU, 2 sigmoid.  workspace: gml_fc_gemm_workspace_bytes() bytes (split-K
 * partials and tickets); may be NULL (no split-K, generic kernel).
 */
size_t gml_fc_gemm_workspace_bytes(void);
int gml_fc_gemm(const float* a, const float* b, float* c, const float* bias,
                int32_t m, int32_t n, int32_t k, int32_t lda, int32_t ldb, int32_t ldc,
                int32_t a_kc, int32_t b_kc, int32_t act, int32_t beta,
                void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GML_B200_H_ */
