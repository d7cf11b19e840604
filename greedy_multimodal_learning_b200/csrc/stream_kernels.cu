// Streaming (two-pass) MMTM kernels: per-plane reductions and per-plane scaling.
//
// These are the general-shape kernels: any N, C, HW, any 4-byte aligned pointers.  A feature
// map [N, C, HW] is seen as rows = N*C planes of HW contiguous floats.  Each plane is owned by
// a group of L consecutive lanes (L = 1..32, power of two) so that
//   - a warp's 128-bit loads cover (32/L) adjacent planes = one contiguous span of memory,
//   - the plane sum is a fixed-order shuffle tree  -> bit-reproducible results,
//   - the per-plane scalar (gate, squeeze gradient) is fetched once per plane, no index division
//     in the inner loop.
// Both modalities travel in one launch (two "segments", block ranges back to back).
//
// HBM traffic (u = N*C*HW*4 bytes per modality): reduce = 1u read (2u for the backward dot),
// scale = 1u read + 1u write.  The second pass re-reads what the first pass just streamed; the
// first pass therefore loads with L2 evict_last and the host side chunks the batch so that one
// chunk fits L2 (see capi.cu), the second pass reads with evict_first.
#include "common.cuh"
#include "kernels.h"

namespace gml {

namespace {

constexpr int kThreads = 256;

template <int L, int HW_T, bool VEC, bool DUAL, int EPI>
__global__ void __launch_bounds__(kThreads) plane_reduce_kernel(ReduceLaunch p) {
  const int seg_id = (blockIdx.x >= (unsigned)p.seg_blocks0) ? 1 : 0;
  const ReduceSeg s = seg_id ? p.seg[1] : p.seg[0];  // field-wise select from the constant bank (no local copy)
  const int block_in_seg = blockIdx.x - (seg_id ? p.seg_blocks0 : 0);
  const int hw = HW_T > 0 ? HW_T : s.hw;
  constexpr int kRowsPerBlock = kThreads / L;
  const int lane = threadIdx.x % L;
  const int row = block_in_seg * kRowsPerBlock + threadIdx.x / L;
  float acc = 0.f;
  if (row < s.rows) {
    const size_t base = (size_t)row * hw;
    if constexpr (VEC) {
      const int hw4 = hw >> 2;
      const float4* x4 = reinterpret_cast<const float4*>(s.x + base);
      const float4* y4 = DUAL ? reinterpret_cast<const float4*>(s.y + base) : nullptr;
      constexpr int kUnroll = 4;
      // first pass of two: ask L2 to hold on to what we stream (second pass re-reads it)
      const uint64_t pol = p.keep_in_l2 ? policy_evict_last() : policy_evict_first();
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int j = lane;
      // main loop: kUnroll independent 128-bit loads in flight per stream
      for (; j + (kUnroll - 1) * L < hw4; j += kUnroll * L) {
        float4 xv[kUnroll], yv[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          xv[u] = ldg_hint(x4 + j + u * L, pol);
          if constexpr (DUAL) yv[u] = ldg_stream(y4 + j + u * L);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          if constexpr (DUAL) {
            a0 = fmaf(xv[u].x, yv[u].x, a0); a1 = fmaf(xv[u].y, yv[u].y, a1);
            a2 = fmaf(xv[u].z, yv[u].z, a2); a3 = fmaf(xv[u].w, yv[u].w, a3);
          } else {
            a0 += xv[u].x; a1 += xv[u].y; a2 += xv[u].z; a3 += xv[u].w;
          }
        }
      }
      for (; j < hw4; j += L) {
        float4 xv = ldg_hint(x4 + j, pol);
        if constexpr (DUAL) {
          float4 yv = ldg_stream(y4 + j);
          a0 = fmaf(xv.x, yv.x, a0); a1 = fmaf(xv.y, yv.y, a1);
          a2 = fmaf(xv.z, yv.z, a2); a3 = fmaf(xv.w, yv.w, a3);
        } else {
          a0 += xv.x; a1 += xv.y; a2 += xv.z; a3 += xv.w;
        }
      }
      acc = (a0 + a1) + (a2 + a3);
    } else {
      // planes that are not 16-byte tileable (HW = 49): scalar loads, 8 independent ones in flight per lane
      const float* x = s.x + base;
      const float* y = DUAL ? s.y + base : nullptr;
      const uint64_t pol = p.keep_in_l2 ? policy_evict_last() : policy_evict_first();
      float a0 = 0.f, a1 = 0.f;
      for (int j0 = lane; j0 < hw; j0 += 8 * L) {
        float xv[8], yv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = j0 + u * L;
          xv[u] = j < hw ? ldg_hint_f32(x + j, pol) : 0.f;
          if constexpr (DUAL) yv[u] = j < hw ? __ldg(y + j) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; u += 2) {
          if constexpr (DUAL) {
            a0 = fmaf(xv[u], yv[u], a0); a1 = fmaf(xv[u + 1], yv[u + 1], a1);
          } else {
            a0 += xv[u]; a1 += xv[u + 1];
          }
        }
      }
      acc = a0 + a1;
    }
  }
  acc = group_sum<L>(acc);
  if (row < s.rows && lane == 0) {
    const int n = row / s.c, c = row - n * s.c;
    const size_t o = (size_t)n * s.out_ld + s.out_off + c;
    if constexpr (EPI == kEpiMean) {
      s.out[o] = acc / (float)hw;  // torch.mean = sum / count (balanced_mmtm.py:97)
    } else {                      // kEpiDGate: dE = gate_scale * dot * g * (1 - g)
      const float g = s.gate[row];
      s.out[o] = acc * s.mul * g * (1.f - g);
    }
  }
}

template <int L, int HW_T, bool VEC, bool HAS_ADD>
__global__ void __launch_bounds__(kThreads) plane_scale_kernel(ScaleLaunch p) {
  const int seg_id = (blockIdx.x >= (unsigned)p.seg_blocks0) ? 1 : 0;
  const ScaleSeg s = seg_id ? p.seg[1] : p.seg[0];
  const int block_in_seg = blockIdx.x - (seg_id ? p.seg_blocks0 : 0);
  const int hw = HW_T > 0 ? HW_T : s.hw;
  constexpr int kRowsPerBlock = kThreads / L;
  const int lane = threadIdx.x % L;
  const int row = block_in_seg * kRowsPerBlock + threadIdx.x / L;
  if (row >= s.rows) return;
  const int n = row / s.c, c = row - n * s.c;
  const float sc = s.scale[s.scale_bcast ? c : row] * s.mul;
  float ad = 0.f;
  if constexpr (HAS_ADD) ad = s.add[(size_t)n * s.add_ld + s.add_off + c] / (float)hw;  // MeanBackward: grad / HW
  const size_t base = (size_t)row * hw;
  if constexpr (VEC) {
    const int hw4 = hw >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(s.x + base);
    float4* o4 = reinterpret_cast<float4*>(s.out + base);
    constexpr int kUnroll = 4;
    const uint64_t pol = policy_evict_first();  // last use of this data
    int j = lane;
    for (; j + (kUnroll - 1) * L < hw4; j += kUnroll * L) {
      float4 v[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) v[u] = ldg_hint(x4 + j + u * L, pol);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        float4 r;
        r.x = fmaf(v[u].x, sc, ad); r.y = fmaf(v[u].y, sc, ad);
        r.z = fmaf(v[u].z, sc, ad); r.w = fmaf(v[u].w, sc, ad);
        stg_stream(o4 + j + u * L, r);
      }
    }
    for (; j < hw4; j += L) {
      float4 v = ldg_hint(x4 + j, pol), r;
      r.x = fmaf(v.x, sc, ad); r.y = fmaf(v.y, sc, ad); r.z = fmaf(v.z, sc, ad); r.w = fmaf(v.w, sc, ad);
      stg_stream(o4 + j, r);
    }
  } else {
    const float* x = s.x + base;
    float* o = s.out + base;
    const uint64_t pol = policy_evict_first();
    for (int j0 = lane; j0 < hw; j0 += 8 * L) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = j0 + u * L;
        v[u] = j < hw ? ldg_hint_f32(x + j, pol) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = j0 + u * L;
        if (j < hw) o[j] = fmaf(v[u], sc, ad);
      }
    }
  }
}


// ---- planes that are not a multiple of 4 floats (HW = 49): four adjacent planes form one span of
// HW float4 that IS 16-byte tileable.  A lane group owns such a quad; every element picks its plane
// with three compares.  Same traffic and coalescing as the vector path above, ~4x the ALU work per
// element (irrelevant: the kernels are HBM-bound).
__device__ __forceinline__ int plane_in_quad(int e, int hw) { return (e >= hw) + (e >= 2 * hw) + (e >= 3 * hw); }

template <int L, bool DUAL, int EPI>
__global__ void __launch_bounds__(kThreads) quad_reduce_kernel(ReduceLaunch p) {
  const int seg_id = (blockIdx.x >= (unsigned)p.seg_blocks0) ? 1 : 0;
  const ReduceSeg s = seg_id ? p.seg[1] : p.seg[0];
  const int block_in_seg = blockIdx.x - (seg_id ? p.seg_blocks0 : 0);
  const int hw = s.hw;
  constexpr int kQuadsPerBlock = kThreads / L;
  const int lane = threadIdx.x % L;
  const int quad = block_in_seg * kQuadsPerBlock + threadIdx.x / L;
  const int nquads = s.rows >> 2;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (quad < nquads) {
    const size_t base = (size_t)quad * 4 * hw;
    const float4* x4 = reinterpret_cast<const float4*>(s.x + base);
    const float4* y4 = DUAL ? reinterpret_cast<const float4*>(s.y + base) : nullptr;
    const uint64_t pol = p.keep_in_l2 ? policy_evict_last() : policy_evict_first();
    for (int j0 = lane; j0 < hw; j0 += 8 * L) {
      float4 xv[8], yv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = j0 + u * L;
        xv[u] = j < hw ? ldg_hint(x4 + j, pol) : make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (DUAL) yv[u] = j < hw ? ldg_stream(y4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int e0 = 4 * (j0 + u * L);
        float v[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
        if constexpr (DUAL) { v[0] *= yv[u].x; v[1] *= yv[u].y; v[2] *= yv[u].z; v[3] *= yv[u].w; }
        const int p0 = plane_in_quad(e0, hw), p3 = plane_in_quad(e0 + 3, hw);  // out-of-range lanes carry zeros
        if (p0 == p3) {  // the common case: the whole float4 lies inside one plane
          const float sum4 = (v[0] + v[1]) + (v[2] + v[3]);
#pragma unroll
          for (int t = 0; t < 4; ++t) acc[t] += (p0 == t) ? sum4 : 0.f;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int pq = plane_in_quad(e0 + q, hw);
#pragma unroll
            for (int t = 0; t < 4; ++t) acc[t] += (pq == t) ? v[q] : 0.f;
          }
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) acc[t] = group_sum<L>(acc[t]);
  if (quad < nquads && lane < 4) {
    const int row = quad * 4 + lane;
    const float a = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
    const int n = row / s.c, c = row - n * s.c;
    const size_t o = (size_t)n * s.out_ld + s.out_off + c;
    if constexpr (EPI == kEpiMean) {
      s.out[o] = a / (float)hw;
    } else {
      const float g = s.gate[row];
      s.out[o] = a * s.mul * g * (1.f - g);
    }
  }
}

template <int L, bool HAS_ADD>
__global__ void __launch_bounds__(kThreads) quad_scale_kernel(ScaleLaunch p) {
  const int seg_id = (blockIdx.x >= (unsigned)p.seg_blocks0) ? 1 : 0;
  const ScaleSeg s = seg_id ? p.seg[1] : p.seg[0];
  const int block_in_seg = blockIdx.x - (seg_id ? p.seg_blocks0 : 0);
  const int hw = s.hw;
  constexpr int kQuadsPerBlock = kThreads / L;
  const int lane = threadIdx.x % L;
  const int quad = block_in_seg * kQuadsPerBlock + threadIdx.x / L;
  if (quad >= (s.rows >> 2)) return;
  float sc[4], ad[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int row = quad * 4 + t;
    const int n = row / s.c, c = row - n * s.c;
    sc[t] = s.scale[s.scale_bcast ? c : row] * s.mul;
    ad[t] = 0.f;
    if constexpr (HAS_ADD) ad[t] = s.add[(size_t)n * s.add_ld + s.add_off + c] / (float)hw;
  }
  const size_t base = (size_t)quad * 4 * hw;
  const float4* x4 = reinterpret_cast<const float4*>(s.x + base);
  float4* o4 = reinterpret_cast<float4*>(s.out + base);
  const uint64_t pol = policy_evict_first();
  for (int j0 = lane; j0 < hw; j0 += 8 * L) {
    float4 xv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = j0 + u * L;
      if (j < hw) xv[u] = ldg_hint(x4 + j, pol);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = j0 + u * L;
      if (j < hw) {
        float v[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
        const int p0 = plane_in_quad(4 * j, hw), p3 = plane_in_quad(4 * j + 3, hw);
        if (p0 == p3) {
          const float m = p0 == 0 ? sc[0] : (p0 == 1 ? sc[1] : (p0 == 2 ? sc[2] : sc[3]));
          const float a = p0 == 0 ? ad[0] : (p0 == 1 ? ad[1] : (p0 == 2 ? ad[2] : ad[3]));
#pragma unroll
          for (int q = 0; q < 4; ++q) v[q] = fmaf(v[q], m, a);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int pq = plane_in_quad(4 * j + q, hw);
            const float m = pq == 0 ? sc[0] : (pq == 1 ? sc[1] : (pq == 2 ? sc[2] : sc[3]));
            const float a = pq == 0 ? ad[0] : (pq == 1 ? ad[1] : (pq == 2 ? ad[2] : ad[3]));
            v[q] = fmaf(v[q], m, a);
          }
        }
        stg_stream(o4 + j, make_float4(v[0], v[1], v[2], v[3]));
      }
    }
  }
}

inline int pick_quad_lanes(int hw) {  // hw float4 per quad
  int l = 4;
  while (l < 32 && hw > l * 8) l <<= 1;
  return l;
}

// lanes per plane: keep ~4-8 128-bit loads per lane
inline int pick_lanes(int hw, bool vec) {
  const int items = vec ? hw / 4 : hw;
  int l = 1;
  while (l < 32 && items > l * 8) l <<= 1;
  return l;
}

template <int HW_T, bool VEC, bool DUAL, int EPI>
int launch_reduce_l(int lanes, const ReduceLaunch& p, int blocks, cudaStream_t st) {
  LaunchScope ls(DUAL ? kTagDGate : kTagMean, st);
  switch (lanes) {
#define GML_CASE(LL) \
  case LL: plane_reduce_kernel<LL, HW_T, VEC, DUAL, EPI><<<blocks, kThreads, 0, st>>>(p); break;
    GML_CASE(1) GML_CASE(2) GML_CASE(4) GML_CASE(8) GML_CASE(16) GML_CASE(32)
#undef GML_CASE
    default: return GML_E_BADARG;
  }
  GML_LAUNCH_CHECK();
  return GML_OK;
}

template <int HW_T, bool VEC, bool HAS_ADD>
int launch_scale_l(int lanes, const ScaleLaunch& p, int blocks, cudaStream_t st) {
  LaunchScope ls(HAS_ADD ? kTagScaleBwd : kTagScaleFwd, st);
  switch (lanes) {
#define GML_CASE(LL) \
  case LL: plane_scale_kernel<LL, HW_T, VEC, HAS_ADD><<<blocks, kThreads, 0, st>>>(p); break;
    GML_CASE(1) GML_CASE(2) GML_CASE(4) GML_CASE(8) GML_CASE(16) GML_CASE(32)
#undef GML_CASE
    default: return GML_E_BADARG;
  }
  GML_LAUNCH_CHECK();
  return GML_OK;
}

}  // namespace

// Launch the reduction for up to two segments.  Segments with rows == 0 are skipped.  Segments of
// different HW / alignment are launched separately (each launch is homogeneous).
template <bool DUAL, int EPI>
static int reduce_dispatch(ReduceSeg a, ReduceSeg b, bool keep_in_l2, cudaStream_t st) {
  auto vec_ok = [](const ReduceSeg& s) {
    return s.hw % 4 == 0 && aligned16(s.x) && (!DUAL || aligned16(s.y));
  };
  auto quad_ok = [](const ReduceSeg& s) {
    return s.hw % 4 != 0 && s.rows % 4 == 0 && aligned16(s.x) && (!DUAL || aligned16(s.y));
  };
  auto launch = [&](const ReduceSeg& s0, const ReduceSeg* s1) -> int {
    if (quad_ok(s0) && (!s1 || quad_ok(*s1))) {
      const int lanes = pick_quad_lanes(s0.hw);
      const int qpb = kThreads / lanes;
      ReduceLaunch p;
      p.keep_in_l2 = keep_in_l2 ? 1 : 0;
      p.seg[0] = s0;
      p.seg_blocks0 = ceil_div(s0.rows / 4, qpb);
      int blocks = p.seg_blocks0;
      if (s1) { p.seg[1] = *s1; blocks += ceil_div(s1->rows / 4, qpb); } else { p.seg[1] = s0; p.seg[1].rows = 0; }
      LaunchScope ls(DUAL ? kTagDGate : kTagMean, st);
      switch (lanes) {
        case 4: quad_reduce_kernel<4, DUAL, EPI><<<blocks, kThreads, 0, st>>>(p); break;
        case 8: quad_reduce_kernel<8, DUAL, EPI><<<blocks, kThreads, 0, st>>>(p); break;
        case 16: quad_reduce_kernel<16, DUAL, EPI><<<blocks, kThreads, 0, st>>>(p); break;
        default: quad_reduce_kernel<32, DUAL, EPI><<<blocks, kThreads, 0, st>>>(p); break;
      }
      GML_LAUNCH_CHECK();
      return GML_OK;
    }
    const bool vec = vec_ok(s0);
    const int lanes = pick_lanes(s0.hw, vec);
    const int rpb = kThreads / lanes;
    ReduceLaunch p;
    p.keep_in_l2 = keep_in_l2 ? 1 : 0;
    p.seg[0] = s0;
    p.seg_blocks0 = ceil_div(s0.rows, rpb);
    int blocks = p.seg_blocks0;
    if (s1) {
      p.seg[1] = *s1;
      blocks += ceil_div(s1->rows, rpb);
    } else {
      p.seg[1] = s0;
      p.seg[1].rows = 0;
    }
    if (vec) return launch_reduce_l<0, true, DUAL, EPI>(lanes, p, blocks, st);
    return launch_reduce_l<0, false, DUAL, EPI>(lanes, p, blocks, st);
  };
  const bool has_a = a.rows > 0, has_b = b.rows > 0;
  if (has_a && has_b && a.hw == b.hw && vec_ok(a) == vec_ok(b) && quad_ok(a) == quad_ok(b)) return launch(a, &b);
  if (has_a) GML_TRY(launch(a, nullptr));
  if (has_b) GML_TRY(launch(b, nullptr));
  return GML_OK;
}

int launch_plane_mean(const ReduceSeg& a, const ReduceSeg& b, bool keep_in_l2, cudaStream_t st) {
  return reduce_dispatch<false, kEpiMean>(a, b, keep_in_l2, st);
}

int launch_plane_dgate(const ReduceSeg& a, const ReduceSeg& b, bool keep_in_l2, cudaStream_t st) {
  return reduce_dispatch<true, kEpiDGate>(a, b, keep_in_l2, st);
}

int launch_plane_scale(const ScaleSeg& a, const ScaleSeg& b, bool has_add, cudaStream_t st) {
  auto vec_ok = [](const ScaleSeg& s) { return s.hw % 4 == 0 && aligned16(s.x) && aligned16(s.out); };
  auto quad_ok = [](const ScaleSeg& s) {
    return s.hw % 4 != 0 && s.rows % 4 == 0 && aligned16(s.x) && aligned16(s.out);
  };
  auto launch = [&](const ScaleSeg& s0, const ScaleSeg* s1) -> int {
    if (quad_ok(s0) && (!s1 || quad_ok(*s1))) {
      const int lanes = pick_quad_lanes(s0.hw);
      const int qpb = kThreads / lanes;
      ScaleLaunch p;
      p.seg[0] = s0;
      p.seg_blocks0 = ceil_div(s0.rows / 4, qpb);
      int blocks = p.seg_blocks0;
      if (s1) { p.seg[1] = *s1; blocks += ceil_div(s1->rows / 4, qpb); } else { p.seg[1] = s0; p.seg[1].rows = 0; }
      LaunchScope ls(has_add ? kTagScaleBwd : kTagScaleFwd, st);
#define GML_QS(LL)                                                              \
  if (has_add) quad_scale_kernel<LL, true><<<blocks, kThreads, 0, st>>>(p);     \
  else quad_scale_kernel<LL, false><<<blocks, kThreads, 0, st>>>(p)
      switch (lanes) {
        case 4: GML_QS(4); break;
        case 8: GML_QS(8); break;
        case 16: GML_QS(16); break;
        default: GML_QS(32); break;
      }
#undef GML_QS
      GML_LAUNCH_CHECK();
      return GML_OK;
    }
    const bool vec = vec_ok(s0);
    const int lanes = pick_lanes(s0.hw, vec);
    const int rpb = kThreads / lanes;
    ScaleLaunch p;
    p.seg[0] = s0;
    p.seg_blocks0 = ceil_div(s0.rows, rpb);
    int blocks = p.seg_blocks0;
    if (s1) {
      p.seg[1] = *s1;
      blocks += ceil_div(s1->rows, rpb);
    } else {
      p.seg[1] = s0;
      p.seg[1].rows = 0;
    }
    if (vec) {
      return has_add ? launch_scale_l<0, true, true>(lanes, p, blocks, st)
                     : launch_scale_l<0, true, false>(lanes, p, blocks, st);
    }
    return has_add ? launch_scale_l<0, false, true>(lanes, p, blocks, st)
                   : launch_scale_l<0, false, false>(lanes, p, blocks, st);
  };
  const bool has_a = a.rows > 0, has_b = b.rows > 0;
  if (has_a && has_b && a.hw == b.hw && vec_ok(a) == vec_ok(b) && quad_ok(a) == quad_ok(b)) return launch(a, &b);
  if (has_a) GML_TRY(launch(a, nullptr));
  if (has_b) GML_TRY(launch(b, nullptr));
  return GML_OK;
}

}  // namespace gml
