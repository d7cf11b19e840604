// extern "C" entry points of libgml_b200 (declared in include/gml_b200.h): argument checking,
// workspace carving, kernel-path selection.  No state, no allocation, nothing throws.
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace gml {

static thread_local char g_last_cuda_error[256] = "";

void set_last_cuda_error(cudaError_t e, const char* what) {
  snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s: %s (%s)", what, cudaGetErrorName(e),
           cudaGetErrorString(e));
}

// ---- launch accounting -------------------------------------------------------------------
struct ProfRecord {
  int tag;
  cudaEvent_t start, stop;
};
static std::mutex g_prof_mu;
static std::atomic<long long> g_launches[kTagCount];
static std::atomic<int> g_prof_on{0};
static std::vector<ProfRecord*> g_prof_pending;
static double g_prof_ms[kTagCount];
static long long g_prof_n[kTagCount];

LaunchScope::LaunchScope(int tag, cudaStream_t st) : tag_(tag), st_(st), rec_(nullptr) {
  g_launches[tag].fetch_add(1, std::memory_order_relaxed);
  if (g_prof_on.load(std::memory_order_relaxed)) {
    ProfRecord* r = new ProfRecord;
    r->tag = tag;
    if (cudaEventCreate(&r->start) != cudaSuccess || cudaEventCreate(&r->stop) != cudaSuccess) {
      delete r;
      return;
    }
    cudaEventRecord(r->start, st);
    rec_ = r;
  }
}
LaunchScope::~LaunchScope() {
  if (!rec_) return;
  ProfRecord* r = static_cast<ProfRecord*>(rec_);
  cudaEventRecord(r->stop, st_);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_pending.push_back(r);
}

static void prof_drain() {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (ProfRecord* r : g_prof_pending) {
    float ms = 0.f;
    if (cudaEventSynchronize(r->stop) == cudaSuccess && cudaEventElapsedTime(&ms, r->start, r->stop) == cudaSuccess) {
      g_prof_ms[r->tag] += ms;
      g_prof_n[r->tag] += 1;
    }
    cudaEventDestroy(r->start);
    cudaEventDestroy(r->stop);
    delete r;
  }
  g_prof_pending.clear();
}

// Lazily created per-thread, per-device helper stream: lets the weight-gradient GEMMs of the streaming
// backward run beside the dZ GEMM + gradient-apply pass (fork/join with events; capturable in a CUDA graph).
struct SideStream {
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
static SideStream* side_stream() {
  static thread_local SideStream ss[16];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  SideStream& s = ss[dev];
  if (s.device != dev) {
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    s.device = dev;
  }
  return &s;
}
std::atomic<int> g_overlap_wgrad{1};

namespace {

struct Dims {
  int n, c_v, c_s, hw_v, hw_s, d, ldz, zrows, hoff;
};

int check_dims(const gml_mmtm_dims* in, int mode, Dims* o) {
  if (!in) return GML_E_BADARG;
  if (mode < GML_MODE_NORMAL || mode > GML_MODE_XMODAL_OFF) return GML_E_BADARG;
  if (in->n < 0 || in->c_v <= 0 || in->c_s <= 0 || in->hw_v <= 0 || in->hw_s <= 0 || in->d <= 0) return GML_E_BADARG;
  if ((long long)in->n * in->c_v >= (1LL << 31) || (long long)in->n * in->c_s >= (1LL << 31)) return GML_E_UNSUPPORTED;
  // the reference sizes BOTH running means by dim_visual and feeds both from the visual gate
  // (balanced_mmtm.py:30-31,113-114); substituting the skeleton gate therefore needs c_s == c_v.
  if (mode == GML_MODE_CURATE_SKELETON && in->c_s != in->c_v) return GML_E_UNSUPPORTED;
  o->n = in->n; o->c_v = in->c_v; o->c_s = in->c_s; o->hw_v = in->hw_v; o->hw_s = in->hw_s; o->d = in->d;
  o->ldz = in->c_v + in->c_s;
  o->zrows = mode == GML_MODE_XMODAL_OFF ? 2 * in->n : in->n;
  o->hoff = mode == GML_MODE_XMODAL_OFF ? in->n : 0;  // row offset of the hidden state that feeds g_b
  return GML_OK;
}

// How many samples to push through both passes before moving on, so that the second pass
// (gating / gradient apply) re-reads its input from L2 instead of HBM.  B200 has ~126 MB of L2;
// the default budget leaves room for the output lines that are being written back.
std::atomic<long> g_l2_chunk_mb{-1};
size_t l2_budget_bytes() {
  long mb = g_l2_chunk_mb.load();
  if (mb < 0) {
    const char* e = getenv("GML_L2_CHUNK_MB");
    mb = e ? atol(e) : (1L << 20);  // default: no chunking (measured: per-chunk launches cost more than L2 reuse gains)
    if (mb < 1) mb = 1;
    g_l2_chunk_mb.store(mb);
  }
  return (size_t)mb << 20;
}

int chunk_samples(const Dims& d, size_t bytes_per_sample) {
  size_t n = l2_budget_bytes() / (bytes_per_sample ? bytes_per_sample : 1);
  if (n < 1) n = 1;
  return n > (size_t)d.n ? d.n : (int)n;
}

bool live_a(int mode) { return mode != GML_MODE_CURATE_VISUAL; }
bool live_b(int mode) { return mode != GML_MODE_CURATE_SKELETON; }

// squeeze + excitation for samples [n0, n0 + cn)
int gates_chunk(const float* a, const float* b, const float* w_sq, const float* b_sq, const float* w_v,
                const float* b_v, const float* w_s, const float* b_s, float* z, float* h, float* g_a, float* g_b,
                const float* m_a, const float* m_b, const Dims& d, int mode, int n0, int cn, bool keep, cudaStream_t st,
                void* gws, size_t gws_bytes) {
  const bool x3 = mode == GML_MODE_XMODAL_OFF;
  float* z_a = z + (size_t)n0 * d.ldz;                               // rows that receive s_a
  float* z_b = z + (size_t)(d.hoff + n0) * d.ldz;                    // rows that receive s_b
  if (x3) {
    GML_TRY(launch_fill_rows(z_a, cn, d.ldz, d.c_v, m_b, d.c_s, st));  // [s_a | m_b]
    GML_TRY(launch_fill_rows(z_b, cn, d.ldz, 0, m_a, d.c_v, st));      // [m_a | s_b]
  }
  ReduceSeg sa{a + (size_t)n0 * d.c_v * d.hw_v, nullptr, z_a, nullptr, cn * d.c_v, d.hw_v, d.c_v, d.ldz, 0, 1.f};
  ReduceSeg sb{b + (size_t)n0 * d.c_s * d.hw_s, nullptr, z_b, nullptr, cn * d.c_s, d.hw_s, d.c_s, d.ldz, d.c_v, 1.f};
  GML_TRY(launch_plane_mean(sa, sb, keep, st));
  // H = relu(Z Wsq^T + bsq)
  GemmDesc g1[2];
  int cnt1 = 1;
  g1[0] = GemmDesc{z_a, w_sq, h + (size_t)n0 * d.d, b_sq, nullptr, cn, d.d, d.ldz, d.ldz, d.ldz, d.d, 0, 1, 1, kActRelu, 0};
  if (x3) {
    g1[1] = g1[0];
    g1[1].a = z_b;
    g1[1].c = h + (size_t)(d.hoff + n0) * d.d;
    cnt1 = 2;
  }
  GML_TRY(launch_gemm(g1, cnt1, st, gws, gws_bytes));
  // gates
  GemmDesc g2[2];
  g2[0] = GemmDesc{h + (size_t)n0 * d.d, w_v, g_a + (size_t)n0 * d.c_v, b_v, nullptr, cn, d.c_v, d.d, d.d, d.d, d.c_v, 0, 1, 1, kActSigmoid, 0};
  g2[1] = GemmDesc{h + (size_t)(d.hoff + n0) * d.d, w_s, g_b + (size_t)n0 * d.c_s, b_s, nullptr, cn, d.c_s, d.d, d.d, d.d, d.c_s, 0, 1, 1, kActSigmoid, 0};
  GML_TRY(launch_gemm(g2, 2, st, gws, gws_bytes));
  return GML_OK;
}

int apply_chunk(const float* a, const float* b, float* a_out, float* b_out, const float* g_a, const float* g_b,
                const float* run_v, const float* run_s, const Dims& d, int mode, float gate_scale, int n0, int cn,
                bool do_a, bool do_b, cudaStream_t st) {
  ScaleSeg sa{a + (size_t)n0 * d.c_v * d.hw_v, a_out + (size_t)n0 * d.c_v * d.hw_v,
              live_a(mode) ? g_a + (size_t)n0 * d.c_v : run_v, nullptr, do_a ? cn * d.c_v : 0, d.hw_v, d.c_v,
              live_a(mode) ? 0 : 1, 0, 0, gate_scale};
  ScaleSeg sb{b + (size_t)n0 * d.c_s * d.hw_s, b_out + (size_t)n0 * d.c_s * d.hw_s,
              live_b(mode) ? g_b + (size_t)n0 * d.c_s : run_s, nullptr, do_b ? cn * d.c_s : 0, d.hw_s, d.c_s,
              live_b(mode) ? 0 : 1, 0, 0, gate_scale};
  return launch_plane_scale(sa, sb, false, st);
}

}  // namespace
}  // namespace gml

using namespace gml;

extern "C" int gml_abi_version(void) { return GML_ABI_VERSION; }

extern "C" const char* gml_error_string(int code) {
  switch (code) {
    case GML_OK: return "ok";
    case GML_E_BADARG: return "bad argument (null pointer, non-positive size or unknown mode)";
    case GML_E_ALIGN: return "misaligned pointer";
    case GML_E_WORKSPACE: return "workspace too small";
    case GML_E_CUDA: return "CUDA runtime error (see gml_last_cuda_error)";
    case GML_E_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown error code";
  }
}

extern "C" const char* gml_last_cuda_error(void) { return g_last_cuda_error; }

extern "C" int64_t gml_launch_count(int tag) {
  if (tag >= 0 && tag < kTagCount) return g_launches[tag].load();
  long long t = 0;
  for (int i = 0; i < kTagCount; ++i) t += g_launches[i].load();
  return t;
}

extern "C" const char* gml_kernel_tag_name(int tag) {
  static const char* names[kTagCount] = {"plane_mean", "plane_dgate", "plane_scale_fwd", "plane_scale_bwd",
                                         "fc_gemm", "small", "fused_fwd", "fused_bwd", "sqnorm", "stats"};
  return (tag >= 0 && tag < kTagCount) ? names[tag] : "";
}

extern "C" int gml_kernel_tag_count(void) { return kTagCount; }

extern "C" void gml_profile_enable(int on) {
  if (!on) prof_drain();
  g_prof_on.store(on ? 1 : 0);
}

extern "C" void gml_profile_reset(void) {
  prof_drain();
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < kTagCount; ++i) { g_prof_ms[i] = 0.0; g_prof_n[i] = 0; }
}

extern "C" int gml_profile_read(int tag, double* total_ms, int64_t* launches) {
  if (tag < 0 || tag >= kTagCount || !total_ms || !launches) return GML_E_BADARG;
  prof_drain();
  std::lock_guard<std::mutex> lk(g_prof_mu);
  *total_ms = g_prof_ms[tag];
  *launches = g_prof_n[tag];
  return GML_OK;
}

namespace gml {
extern int g_fused_cluster;
extern int g_fused_threads;
extern long long* g_fused_trace;
extern long long* g_fused_occ_trace;
extern int g_fused_kind;
extern int g_fused_prefetch;
extern int g_fused_occ;
extern int g_gemm_big_tiles;
extern int g_gemm_tf32x3;
extern int g_gemm_umma;
extern long long* g_gemm_trace;
extern int g_fused_weight_ratio_x100;
extern int g_fused_group_kb;
extern int g_fused_stash_kb;
extern int g_fused_wsmem;
extern int g_fused_hw_special;
extern int g_sq_variant;
extern int g_tile_kind, g_tile_lag, g_tile_gemm_ctas, g_tile_m, g_tile_chunk_kb, g_tile_min_mb;
extern long long* g_tile_stats;
extern int g_tile_nodeps, g_tile_ksplit_tiles, g_tile_switch, g_tile_trace_only, g_tile_rpol, g_tile_max_slots, g_tile_split_copies, g_tile_draw, g_tile_chunk_kb_fwd, g_tile_min_mb_light, g_tile_wgrad, g_tile_light_fwd;
}
extern "C" int gml_set_tunable(const char* name, int64_t value) {
  if (!name) return GML_E_BADARG;
  if (!strcmp(name, "l2_chunk_mb")) { g_l2_chunk_mb.store(value < 1 ? 1 : (long)value); return GML_OK; }
  if (!strcmp(name, "fused_cluster")) {
    if (value != 0 && value != 4 && value != 8 && value != 16) return GML_E_BADARG;
    g_fused_cluster = (int)value; return GML_OK;
  }
  if (!strcmp(name, "fused_kind")) {
    if (value < 0 || value > 2) return GML_E_BADARG;
    g_fused_kind = (int)value; return GML_OK;
  }
  if (!strcmp(name, "fused_weight_ratio_x100")) { g_fused_weight_ratio_x100 = (int)value; return GML_OK; }
  if (!strcmp(name, "overlap_wgrad")) { g_overlap_wgrad.store(value ? 1 : 0); return GML_OK; }
  if (!strcmp(name, "fused_hw_special")) { g_fused_hw_special = value != 0; return GML_OK; }
  if (!strcmp(name, "fused_wsmem")) { g_fused_wsmem = value < 0 ? -1 : (int)(value & 3); return GML_OK; }
  if (!strcmp(name, "fused_stash_kb")) { g_fused_stash_kb = value < 0 ? -1 : (value > 200 ? 200 : (int)value); return GML_OK; }
  if (!strcmp(name, "fused_group_kb")) { g_fused_group_kb = (int)value; return GML_OK; }
  if (!strcmp(name, "gemm_big_tiles")) { g_gemm_big_tiles = value ? 1 : 0; return GML_OK; }
  if (!strcmp(name, "gemm_tf32x3")) { g_gemm_tf32x3 = value ? 1 : 0; return GML_OK; }
  if (!strcmp(name, "gemm_umma")) { g_gemm_umma = value ? 1 : 0; return GML_OK; }
  if (!strcmp(name, "gemm_trace_ptr")) { g_gemm_trace = reinterpret_cast<long long*>(value); return GML_OK; }
  if (!strcmp(name, "fused_occ")) {
    if (value != 4 && value != 5) return GML_E_BADARG;
    g_fused_occ = (int)value; return GML_OK;
  }
  if (!strcmp(name, "fused_prefetch")) { g_fused_prefetch = value < 0 ? 0 : (int)value; return GML_OK; }
  if (!strcmp(name, "fused_occ_trace_ptr")) { g_fused_occ_trace = reinterpret_cast<long long*>(value); return GML_OK; }
  if (!strcmp(name, "fused_trace_ptr")) { g_fused_trace = reinterpret_cast<long long*>(value); return GML_OK; }
  if (!strcmp(name, "tile_kind")) {
    if (value < 0 || value > 2) return GML_E_BADARG;
    g_tile_kind = (int)value; return GML_OK;
  }
  if (!strcmp(name, "tile_lag")) { g_tile_lag = value < 0 ? 0 : (value > 16 ? 16 : (int)value); return GML_OK; }
  if (!strcmp(name, "tile_gemm_ctas")) { g_tile_gemm_ctas = value < 0 ? 0 : (int)value; return GML_OK; }
  if (!strcmp(name, "tile_m")) { g_tile_m = value < 0 ? 0 : (value > 128 ? 128 : (int)value); return GML_OK; }
  if (!strcmp(name, "tile_chunk_kb")) { g_tile_chunk_kb = value < 1 ? 1 : (value > 100 ? 100 : (int)value); return GML_OK; }
  if (!strcmp(name, "tile_ksplit_tiles")) { g_tile_ksplit_tiles = value < 0 ? 0 : (int)value; return GML_OK; }
  if (!strcmp(name, "tile_nodeps")) { g_tile_nodeps = (int)value; return GML_OK; }
  if (!strcmp(name, "tile_stats_ptr")) { g_tile_stats = reinterpret_cast<long long*>(value); return GML_OK; }
  if (!strcmp(name, "tile_trace_only")) { g_tile_trace_only = value != 0; return GML_OK; }
  if (!strcmp(name, "tile_rpol")) { g_tile_rpol = (int)value; return GML_OK; }
  if (!strcmp(name, "tile_max_slots")) { g_tile_max_slots = (int)value; return GML_OK; }
  if (!strcmp(name, "tile_split_copies")) { g_tile_split_copies = value < 1 ? 1 : (int)value; return GML_OK; }
  if (!strcmp(name, "tile_draw")) { g_tile_draw = value < 1 ? 1 : (value > 64 ? 64 : (int)value); return GML_OK; }
  if (!strcmp(name, "tile_chunk_kb_fwd")) { g_tile_chunk_kb_fwd = value < 1 ? 1 : (value > 100 ? 100 : (int)value); return GML_OK; }
  if (!strcmp(name, "tile_min_mb_light")) { g_tile_min_mb_light = value < 0 ? 0 : (int)value; return GML_OK; }
  if (!strcmp(name, "tile_light_fwd")) { g_tile_light_fwd = value != 0; return GML_OK; }
  if (!strcmp(name, "tile_wgrad")) { g_tile_wgrad = value != 0; return GML_OK; }
  if (!strcmp(name, "tile_switch")) { g_tile_switch = value != 0; return GML_OK; }
  if (!strcmp(name, "sq_variant")) { g_sq_variant = (int)value & 127; return GML_OK; }
  if (!strcmp(name, "tile_min_mb")) { g_tile_min_mb = value < 0 ? 0 : (int)value; return GML_OK; }
  if (!strcmp(name, "fused_threads")) {
    if (value != 0 && value != 256 && value != 512) return GML_E_BADARG;
    g_fused_threads = (int)value; return GML_OK;
  }
  return GML_E_BADARG;
}

extern "C" int gml_device_is_blackwell(void) {
  int dev = 0, major = 0;
  GML_CUDA_TRY(cudaGetDevice(&dev));
  GML_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  return major == 10 ? 1 : 0;
}

extern "C" size_t gml_mmtm_fwd_workspace_bytes(const gml_mmtm_dims* dims) {
  size_t need = gemm_workspace_bytes();  // split-K partials of the FC GEMMs (streaming path)
  if (dims && dims->c_v == dims->c_s && dims->hw_v == dims->hw_s) {  // counters + split-K partials of the tile pipeline
    const size_t t = tile_fwd_workspace_bytes(dims->n, dims->c_v, dims->hw_v, dims->d);
    if (t > need) need = t;
  }
  return need;
}

extern "C" int gml_mmtm_gates(const float* a, const float* b, const float* w_sq, const float* b_sq, const float* w_v,
                              const float* b_v, const float* w_s, const float* b_s, float* z, float* h, float* g_a,
                              float* g_b, float* gate_sum, const float* m_a, const float* m_b, void* workspace,
                              size_t workspace_bytes, const gml_mmtm_dims* dims, int mode, void* stream) {
  Dims d;
  GML_TRY(check_dims(dims, mode, &d));
  if (!a || !b || !w_sq || !b_sq || !w_v || !b_v || !w_s || !b_s || !z || !h || !g_a || !g_b) return GML_E_BADARG;
  if (mode == GML_MODE_XMODAL_OFF && (!m_a || !m_b)) return GML_E_BADARG;
  if (d.n == 0) return GML_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GML_TRY(prepare_gemm_workspace(workspace, workspace_bytes, st));
  GML_TRY(gates_chunk(a, b, w_sq, b_sq, w_v, b_v, w_s, b_s, z, h, g_a, g_b, m_a, m_b, d, mode, 0, d.n, false, st,
                      workspace, workspace_bytes));
  if (gate_sum) GML_TRY(launch_colsum(g_a, d.n, d.c_v, d.c_v, gate_sum, st));
  return GML_OK;
}

extern "C" int gml_mmtm_running(float* run_v, float* run_s, const float* gate_sum, int32_t c_v, int64_t n_total,
                                int64_t step, void* stream) {
  if (!run_v || !gate_sum || c_v <= 0 || n_total <= 0 || step < 0) return GML_E_BADARG;
  return launch_running_update(run_v, run_s, gate_sum, c_v, (double)n_total, (double)step,
                               static_cast<cudaStream_t>(stream));
}

extern "C" int gml_mmtm_apply(const float* a, const float* b, float* a_out, float* b_out, const float* g_a,
                              const float* g_b, const float* run_v, const float* run_s, const gml_mmtm_dims* dims,
                              int mode, float gate_scale, void* stream) {
  Dims d;
  GML_TRY(check_dims(dims, mode, &d));
  if (!a || !b || !a_out || !b_out || !g_a || !g_b) return GML_E_BADARG;
  if ((mode == GML_MODE_CURATE_VISUAL && !run_v) || (mode == GML_MODE_CURATE_SKELETON && !run_s)) return GML_E_BADARG;
  if (d.n == 0) return GML_OK;
  return apply_chunk(a, b, a_out, b_out, g_a, g_b, run_v, run_s, d, mode, gate_scale, 0, d.n, true, true,
                     static_cast<cudaStream_t>(stream));
}

extern "C" int gml_mmtm_fwd(const float* a, const float* b, float* a_out, float* b_out, const float* w_sq,
                            const float* b_sq, const float* w_v, const float* b_v, const float* w_s, const float* b_s,
                            float* z, float* h, float* g_a, float* g_b, float* gate_sum, float* run_v, float* run_s,
                            int64_t step, const float* m_a, const float* m_b, void* workspace, size_t workspace_bytes,
                            const gml_mmtm_dims* dims, int mode, float gate_scale, uint32_t flags, void* stream) {
  Dims d;
  GML_TRY(check_dims(dims, mode, &d));
  if (!a || !b || !a_out || !b_out || !w_sq || !b_sq || !w_v || !b_v || !w_s || !b_s || !z || !h || !g_a || !g_b ||
      !gate_sum)
    return GML_E_BADARG;
  if (mode == GML_MODE_XMODAL_OFF && (!m_a || !m_b)) return GML_E_BADARG;
  const bool update = !(flags & GML_F_NO_RUNNING_UPDATE);
  if (update && (!run_v || !run_s || step < 0)) return GML_E_BADARG;
  const bool curate = mode == GML_MODE_CURATE_VISUAL || mode == GML_MODE_CURATE_SKELETON;
  if (curate && (!run_v || !run_s)) return GML_E_BADARG;
  if (d.n == 0) return GML_OK;
  if (workspace && workspace_bytes < gemm_workspace_bytes()) return GML_E_WORKSPACE;  // NULL = no split-K, always fine
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  const bool can_tile = !(flags & (GML_F_FORCE_STREAMING | GML_F_FORCE_FUSED)) && !curate && workspace &&
                        tile_supported(d.n, d.c_v, d.c_s, d.hw_v, d.hw_s, d.d, mode) &&
                        workspace_bytes >= tile_fwd_workspace_bytes(d.n, d.c_v, d.hw_v, d.d);
  if ((flags & GML_F_FORCE_TILE) && !can_tile) return GML_E_UNSUPPORTED;
  if (can_tile && ((flags & GML_F_FORCE_TILE) || tile_preferred(d.n, d.c_v, d.hw_v, d.d, false))) {
    FusedFwdArgs fa{a, b, a_out, b_out, w_sq, b_sq, w_v, b_v, w_s, b_s, z, h, g_a, g_b, run_v, run_s,
                    d.n, d.c_v, d.hw_v, d.d, mode, gate_scale};
    const int rc = launch_tile_fwd(fa, gate_sum, update ? run_v : nullptr, update ? run_s : nullptr, (float)step,
                                   workspace, workspace_bytes, st);
    if (rc != GML_E_UNSUPPORTED || (flags & GML_F_FORCE_TILE)) return rc;  // unsupported = misaligned pointers
  }

  const bool can_fuse = !(flags & (GML_F_FORCE_STREAMING | GML_F_FORCE_TILE)) && !curate &&
                        fused_supported(d.n, d.c_v, d.c_s, d.hw_v, d.hw_s, d.d, mode);
  if ((flags & GML_F_FORCE_FUSED) && !can_fuse) return GML_E_UNSUPPORTED;
  if (can_fuse && ((flags & GML_F_FORCE_FUSED) || fused_fwd_preferred(d.n, d.c_v, d.hw_v, d.d))) {
    FusedFwdArgs fa{a, b, a_out, b_out, w_sq, b_sq, w_v, b_v, w_s, b_s, z, h, g_a, g_b, run_v, run_s,
                    d.n, d.c_v, d.hw_v, d.d, mode, gate_scale};
    const int rc = launch_fused_fwd(fa, st);
    if (rc == GML_OK) {
      ColsumSeg cs{g_a, gate_sum, d.n, d.c_v, d.c_v, update ? run_v : nullptr, update ? run_s : nullptr, (float)d.n,
                   (float)step};
      return launch_colsums(&cs, 1, st);
    }
    // misaligned pointers: the streaming path below takes any 4-byte alignment
    if (rc != GML_E_UNSUPPORTED || (flags & GML_F_FORCE_FUSED)) return rc;
  }

  GML_TRY(prepare_gemm_workspace(workspace, workspace_bytes, st));
  // Streaming path: batch chunks sized for L2 so the gating pass re-reads from L2.
  const size_t per_sample = ((size_t)d.c_v * d.hw_v + (size_t)d.c_s * d.hw_s) * sizeof(float);
  const int cs = chunk_samples(d, per_sample);
  const bool keep = true;
  for (int n0 = 0; n0 < d.n; n0 += cs) {
    const int cn = (d.n - n0) < cs ? (d.n - n0) : cs;
    GML_TRY(gates_chunk(a, b, w_sq, b_sq, w_v, b_v, w_s, b_s, z, h, g_a, g_b, m_a, m_b, d, mode, n0, cn, keep, st,
                        workspace, workspace_bytes));
    // live sides can be gated right away; a substituted side waits for the running mean
    GML_TRY(apply_chunk(a, b, a_out, b_out, g_a, g_b, run_v, run_s, d, mode, gate_scale, n0, cn, live_a(mode),
                        live_b(mode), st));
  }
  {
    ColsumSeg cs{g_a, gate_sum, d.n, d.c_v, d.c_v, update ? run_v : nullptr, update ? run_s : nullptr, (float)d.n,
                 (float)step};
    GML_TRY(launch_colsums(&cs, 1, st));
  }
  if (curate) {
    // balanced_mmtm.py:139-152: the substituted scale is the running mean INCLUDING this batch.
    // With GML_F_NO_RUNNING_UPDATE the caller has already folded the (all-reduced) batch in.
    GML_TRY(apply_chunk(a, b, a_out, b_out, g_a, g_b, run_v, run_s, d, mode, gate_scale, 0, d.n, !live_a(mode),
                        !live_b(mode), st));
  }
  return GML_OK;
}

extern "C" size_t gml_mmtm_bwd_workspace_bytes(const gml_mmtm_dims* dims) {
  if (!dims || dims->n < 0) return 0;
  const size_t n = (size_t)dims->n, ldz = (size_t)dims->c_v + dims->c_s, dd = (size_t)dims->d;
  // de_a [N,c_v] + de_b [N,c_s] + dh [2N,D] + dz [2N,ldz]   (2N rows cover mode 3)
  size_t need = 256 + sizeof(float) * (n * ldz + 2 * n * dd + 2 * n * ldz) + 4 * 256 + 2 * round_up(gemm_workspace_bytes(), 256);
  if (dims->c_v == dims->c_s && dims->hw_v == dims->hw_s)
    need += tile_bwd_workspace_bytes(dims->n, dims->c_v, dims->hw_v, dims->d);
  return need;
}

extern "C" int gml_mmtm_bwd(const float* go_a, const float* go_b, const float* a, const float* b, const float* w_sq,
                            const float* w_v, const float* w_s, const float* z, const float* h, const float* g_a,
                            const float* g_b, const float* run_v, const float* run_s, const float* m_a,
                            const float* m_b, float* d_a, float* d_b, float* d_w_sq, float* d_b_sq, float* d_w_v,
                            float* d_b_v, float* d_w_s, float* d_b_s, void* workspace, size_t workspace_bytes,
                            const gml_mmtm_dims* dims, int mode, float gate_scale, uint32_t flags, void* stream) {
  (void)m_a; (void)m_b;  // mode 3: the constant halves of Z are already stored in z
  Dims d;
  GML_TRY(check_dims(dims, mode, &d));
  if (!go_a || !go_b || !a || !b || !w_sq || !w_v || !w_s || !z || !h || !g_a || !g_b || !d_a || !d_b || !workspace)
    return GML_E_BADARG;
  if ((mode == GML_MODE_CURATE_VISUAL && !run_v) || (mode == GML_MODE_CURATE_SKELETON && !run_s)) return GML_E_BADARG;
  if (workspace_bytes < gml_mmtm_bwd_workspace_bytes(dims)) return GML_E_WORKSPACE;
  if (d.n == 0) return GML_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool la = live_a(mode), lb = live_b(mode), x3 = mode == GML_MODE_XMODAL_OFF;

  char* wp = static_cast<char*>(workspace);
  auto carve = [&](size_t floats) {
    float* p = reinterpret_cast<float*>(wp);
    wp += round_up(floats * sizeof(float), 256);
    return p;
  };
  float* de_a = carve((size_t)d.n * d.c_v);
  float* de_b = carve((size_t)d.n * d.c_s);
  float* dh = carve((size_t)2 * d.n * d.d);
  float* dz = carve((size_t)2 * d.n * d.ldz);
  void* gws = wp;
  const size_t gws_bytes = gemm_workspace_bytes();
  void* gws2 = wp + round_up(gws_bytes, 256);  // second split-K workspace for the side stream
  void* tile_ws = wp + 2 * round_up(gws_bytes, 256);  // counters, partials, transposed weights of the tile pipeline
  GML_TRY(prepare_gemm_workspace(gws, gws_bytes, st));
  GML_TRY(prepare_gemm_workspace(gws2, gws_bytes, st));

  // weight gradients over the whole batch (reduction over samples inside one CTA per tile or split-K
  // with an ordered fold: deterministic, no atomics)
  bool bias_done = false;   // the tile pipeline forms the bias gradients itself
  bool wgrad_tiled = false; // ... and, when all three are requested, the weight gradients too
  // part: 1 = the weight-gradient GEMMs, 2 = the bias gradients (column sums), 3 = both
  auto weight_grads = [&](cudaStream_t ws_stream, void* ws_mem, int part = 3) -> int {
    GemmDesc gw[3];
    int cw = 0;
    if (wgrad_tiled) return GML_OK;
    if (!(part & 1)) goto bias_part;
    if (d_w_v) {
      if (la) gw[cw++] = GemmDesc{de_a, h, d_w_v, nullptr, nullptr, d.c_v, d.d, d.n, d.c_v, d.d, d.d, 0, 0, 0, kActNone, 0};
      else GML_TRY(launch_fill_zero(d_w_v, (size_t)d.c_v * d.d, ws_stream));
    }
    if (d_w_s) {
      if (lb) gw[cw++] = GemmDesc{de_b, h + (size_t)d.hoff * d.d, d_w_s, nullptr, nullptr, d.c_s, d.d, d.n, d.c_s, d.d, d.d, 0, 0, 0, kActNone, 0};
      else GML_TRY(launch_fill_zero(d_w_s, (size_t)d.c_s * d.d, ws_stream));
    }
    if (d_w_sq) {
      // dW_sq reduces over the same samples as dW_v / dW_s unless the hidden state is doubled (mode 3): one launch
      GemmDesc gq{dh, z, d_w_sq, nullptr, nullptr, d.d, d.ldz, d.zrows, d.d, d.ldz, d.ldz, 0, 0, 0, kActNone, 0};
      if (cw && d.zrows == d.n) {
        gw[cw++] = gq;
      } else {
        GML_TRY(launch_gemm(&gq, 1, ws_stream, ws_mem, gws_bytes));
      }
    }
    if (cw) GML_TRY(launch_gemm(gw, cw, ws_stream, ws_mem, gws_bytes));
  bias_part:
    if (bias_done || !(part & 2)) return GML_OK;
    ColsumSeg segs[3];
    int nseg = 0;
    if (d_b_v) {
      if (la) segs[nseg++] = ColsumSeg{de_a, d_b_v, d.n, d.c_v, d.c_v, nullptr, nullptr, 1.f, 0.f};
      else GML_TRY(launch_fill_zero(d_b_v, d.c_v, ws_stream));
    }
    if (d_b_s) {
      if (lb) segs[nseg++] = ColsumSeg{de_b, d_b_s, d.n, d.c_s, d.c_s, nullptr, nullptr, 1.f, 0.f};
      else GML_TRY(launch_fill_zero(d_b_s, d.c_s, ws_stream));
    }
    if (d_b_sq) segs[nseg++] = ColsumSeg{dh, d_b_sq, d.zrows, d.d, d.d, nullptr, nullptr, 1.f, 0.f};
    if (nseg) GML_TRY(launch_colsums(segs, nseg, ws_stream));
    return GML_OK;
  };
  bool wgrad_done = false;

  const size_t tile_bytes = (d.c_v == d.c_s && d.hw_v == d.hw_s) ? tile_bwd_workspace_bytes(d.n, d.c_v, d.hw_v, d.d) : 0;
  const bool can_tile = !(flags & (GML_F_FORCE_STREAMING | GML_F_FORCE_FUSED)) && la && lb && tile_bytes &&
                        tile_supported(d.n, d.c_v, d.c_s, d.hw_v, d.hw_s, d.d, mode);
  if ((flags & GML_F_FORCE_TILE) && !can_tile) return GML_E_UNSUPPORTED;
  bool tiled = false, fused_done = false;
  if (can_tile && ((flags & GML_F_FORCE_TILE) || tile_preferred(d.n, d.c_v, d.hw_v, d.d, true))) {
    FusedBwdArgs fb{go_a, go_b, a, b, w_sq, w_v, w_s, z, h, g_a, g_b, run_v, run_s, d_a, d_b, de_a, de_b, dh,
                    d.n, d.c_v, d.hw_v, d.d, mode, gate_scale};
    const int rc = launch_tile_bwd(fb, dz, d_b_v, d_b_s, d_b_sq, d_w_v, d_w_s, d_w_sq, &wgrad_tiled, tile_ws, tile_bytes, st);
    if (rc == GML_OK) { tiled = true; bias_done = true; }
    else if (rc != GML_E_UNSUPPORTED || (flags & GML_F_FORCE_TILE)) return rc;
  }

  const bool can_fuse = !tiled && !(flags & (GML_F_FORCE_STREAMING | GML_F_FORCE_TILE)) && la && lb &&
                        fused_supported(d.n, d.c_v, d.c_s, d.hw_v, d.hw_s, d.d, mode);
  if ((flags & GML_F_FORCE_FUSED) && !can_fuse) return GML_E_UNSUPPORTED;
  if (tiled) {
    // d_input, dE, dH and the bias gradients are done; the weight-gradient GEMMs follow below
  } else if (can_fuse) {
    FusedBwdArgs fb{go_a, go_b, a, b, w_sq, w_v, w_s, z, h, g_a, g_b, run_v, run_s, d_a, d_b, de_a, de_b, dh,
                    d.n, d.c_v, d.hw_v, d.d, mode, gate_scale};
    const int rc = launch_fused_bwd(fb, st);
    if (rc != GML_OK && (rc != GML_E_UNSUPPORTED || (flags & GML_F_FORCE_FUSED))) return rc;
    fused_done = rc == GML_OK;
  }
  if (!tiled && !fused_done) {
    // grad_out is read twice (dot, then apply): keep it in L2 per chunk.  a/b stream through once.
    const size_t per_sample = ((size_t)d.c_v * d.hw_v + (size_t)d.c_s * d.hw_s) * sizeof(float);
    const int cs = chunk_samples(d, per_sample);
    for (int n0 = 0; n0 < d.n; n0 += cs) {
      const int cn = (d.n - n0) < cs ? (d.n - n0) : cs;
      const size_t oa = (size_t)n0 * d.c_v * d.hw_v, ob = (size_t)n0 * d.c_s * d.hw_s;
      // 1. dE = gate_scale * <grad_out, input>_hw * g (1 - g) on live sides
      ReduceSeg ra{go_a + oa, a + oa, de_a + (size_t)n0 * d.c_v, g_a + (size_t)n0 * d.c_v, la ? cn * d.c_v : 0,
                   d.hw_v, d.c_v, d.c_v, 0, gate_scale};
      ReduceSeg rb{go_b + ob, b + ob, de_b + (size_t)n0 * d.c_s, g_b + (size_t)n0 * d.c_s, lb ? cn * d.c_s : 0,
                   d.hw_s, d.c_s, d.c_s, 0, gate_scale};
      GML_TRY(launch_plane_dgate(ra, rb, true, st));
      // 2. dH = (dE_a Wv + dE_b Ws) * [H > 0]
      float* dh_a = dh + (size_t)n0 * d.d;
      float* dh_b = dh + (size_t)(d.hoff + n0) * d.d;
      const float* h_a = h + (size_t)n0 * d.d;
      const float* h_b = h + (size_t)(d.hoff + n0) * d.d;
      GemmDesc ga{de_a + (size_t)n0 * d.c_v, w_v, dh_a, nullptr, h_a, cn, d.d, d.c_v, d.c_v, d.d, d.d, d.d, 1, 0, kActNone, 0};
      GemmDesc gb{de_b + (size_t)n0 * d.c_s, w_s, dh_b, nullptr, h_b, cn, d.d, d.c_s, d.c_s, d.d, d.d, d.d, 1, 0, kActReluMask, 0};
      if (x3) {  // two independent hidden states
        ga.act = kActReluMask;
        GemmDesc both[2] = {ga, gb};
        GML_TRY(launch_gemm(both, 2, st, gws, gws_bytes));
      } else if (la && lb) {
        // dH = ([dE_a | dE_b] [W_v ; W_s]) * [H > 0]: one launch over the concatenated K when the pipelined
        // kernels can take it, else two accumulating launches
        GemmDesc gm = ga;
        gm.k = d.c_v + d.c_s;
        gm.k_split = d.c_v;
        gm.a2 = gb.a; gm.b2 = gb.b; gm.lda2 = gb.lda; gm.ldb2 = gb.ldb;
        gm.act = kActReluMask;
        int rc = GML_E_UNSUPPORTED;
        if (gemm_ksplit_ok(gm) && gws) rc = launch_gemm(&gm, 1, st, gws, gws_bytes);
        if (rc == GML_E_UNSUPPORTED) {
          GML_TRY(launch_gemm(&ga, 1, st, gws, gws_bytes));
          gb.beta = 1;
          GML_TRY(launch_gemm(&gb, 1, st, gws, gws_bytes));
        } else {
          GML_TRY(rc);
        }
      } else if (la) {
        ga.act = kActReluMask;
        GML_TRY(launch_gemm(&ga, 1, st, gws, gws_bytes));
      } else {
        GML_TRY(launch_gemm(&gb, 1, st, gws, gws_bytes));
      }
      // the weight gradients only need dE / dH: with the whole batch in one chunk they run on the side
      // stream, next to the dZ GEMM and the (HBM-bound) gradient-apply pass
      SideStream* side = (cn == d.n && g_overlap_wgrad.load()) ? side_stream() : nullptr;
      if (side) {
        GML_CUDA_TRY(cudaEventRecord(side->fork, st));
        GML_CUDA_TRY(cudaStreamWaitEvent(side->stream, side->fork, 0));
        GML_TRY(weight_grads(side->stream, gws2));
        GML_CUDA_TRY(cudaEventRecord(side->join, side->stream));
        wgrad_done = true;
      }
      // 3. dZ = dH Wsq
      GemmDesc gz[2];
      gz[0] = GemmDesc{dh_a, w_sq, dz + (size_t)n0 * d.ldz, nullptr, nullptr, cn, d.ldz, d.d, d.d, d.ldz, d.ldz, 0, 1, 0, kActNone, 0};
      int cz = 1;
      if (x3) {
        gz[1] = gz[0];
        gz[1].a = dh_b;
        gz[1].c = dz + (size_t)(d.hoff + n0) * d.ldz;
        cz = 2;
      }
      GML_TRY(launch_gemm(gz, cz, st, gws, gws_bytes));
      // 4. d_input = grad_out * scale + ds / HW
      ScaleSeg sa{go_a + oa, d_a + oa, la ? g_a + (size_t)n0 * d.c_v : run_v, dz + (size_t)n0 * d.ldz, cn * d.c_v,
                  d.hw_v, d.c_v, la ? 0 : 1, d.ldz, 0, gate_scale};
      ScaleSeg sb{go_b + ob, d_b + ob, lb ? g_b + (size_t)n0 * d.c_s : run_s, dz + (size_t)(d.hoff + n0) * d.ldz,
                  cn * d.c_s, d.hw_s, d.c_s, lb ? 0 : 1, d.ldz, d.c_v, gate_scale};
      GML_TRY(launch_plane_scale(sa, sb, true, st));
      if (side) GML_CUDA_TRY(cudaStreamWaitEvent(st, side->join, 0));
    }
  }

  if (!wgrad_done) {
    // after a cluster kernel: the bias-gradient column sums run on the side stream next to the weight-gradient GEMMs
    // (two ~10 us launches that only share read-only inputs)
    SideStream* side = (fused_done && !bias_done && g_overlap_wgrad.load()) ? side_stream() : nullptr;
    if (side) {
      GML_CUDA_TRY(cudaEventRecord(side->fork, st));
      GML_CUDA_TRY(cudaStreamWaitEvent(side->stream, side->fork, 0));
      GML_TRY(weight_grads(side->stream, gws2, 2));
      GML_CUDA_TRY(cudaEventRecord(side->join, side->stream));
      GML_TRY(weight_grads(st, gws, 1));
      GML_CUDA_TRY(cudaStreamWaitEvent(st, side->join, 0));
    } else {
      GML_TRY(weight_grads(st, gws));
    }
  }
  return GML_OK;
}

// ---- the FC GEMM on its own (diagnostics, kernel tests, timing) -------------------------------
extern "C" size_t gml_fc_gemm_workspace_bytes(void) { return gml::gemm_workspace_bytes(); }

extern "C" int gml_fc_gemm(const float* a, const float* b, float* c, const float* bias, int32_t m, int32_t n, int32_t k,
                           int32_t lda, int32_t ldb, int32_t ldc, int32_t a_kc, int32_t b_kc, int32_t act, int32_t beta,
                           void* workspace, size_t workspace_bytes, void* stream) {
  using namespace gml;
  if (!a || !b || !c || m <= 0 || n <= 0 || k <= 0 || ldc < n) return GML_E_BADARG;
  if (lda < (a_kc ? k : m) || ldb < (b_kc ? k : n)) return GML_E_BADARG;
  if (act < kActNone || act > kActSigmoid || (beta != 0 && beta != 1)) return GML_E_BADARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GemmDesc d{};
  d.a = a; d.b = b; d.c = c; d.bias = bias; d.mask = nullptr;
  d.m = m; d.n = n; d.k = k;
  d.lda = lda; d.ldb = ldb; d.ldc = ldc; d.ldmask = 0;
  d.a_kc = a_kc ? 1 : 0; d.b_kc = b_kc ? 1 : 0;
  d.act = act; d.beta = beta;
  if (workspace && workspace_bytes < gemm_workspace_bytes()) return GML_E_WORKSPACE;
  GML_TRY(prepare_gemm_workspace(workspace, workspace_bytes, st));
  return launch_gemm(&d, 1, st, workspace, workspace_bytes);
}
