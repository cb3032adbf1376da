// Tiled rasterize-and-composite over the cell-sorted records produced by bin.cu.
//
// Replaces pytorch3d's RasterizePointsNaiveCudaKernel (O(N*H*W*P), reached from
// pgdvs_renderer_dyn.py:690-717 with bin_size=0), PointsRenderer's weight computation,
// the NormWeighted/Alpha compositors, the background fill and PGDVS's second all-ones
// render for the mask (pgdvs_renderer_dyn.py:719-722) with ONE pass:
//   * each thread owns one pixel and keeps its K nearest hits as a z-sorted list of
//     (z, record slot) pairs in registers;
//   * candidates are only the records filed under cells within `halo` of the pixel — for
//     every window row that is one contiguous run of the cell-sorted arrays;
//   * the hit test reproduces the reference arithmetic bit for bit (dist2_rn, strict <);
//   * the common-case insertion is a branch-free compare-exchange chain on z alone; exact
//     fp32 z ties (which the CPU rasterizer's (z, idx, dist2) priority queue orders by the
//     smaller packed index) are DETECTED in the fast path and the few affected pixels are
//     redone by a tie-aware slow path, so idx/zbuf/dists stay deterministic and bit-exact.
#include "common.cuh"

namespace pgdvs {

struct RasterParams {
  const int* cell_end;  // per-cell END offsets; start(c) = cell_end[c - 1] (cell_end[-1] == 0)
  const float4* recA;
  const float4* recB;
  int N, H, W, K, C, halo, GW, GH;
  NdcAxis ax, ay;
  float r2;          // scalar radius^2 (fp32 r*r) or < 0: per-point radius in recB.w
  float rr_weight;   // divisor of PointsRenderer's  1 - dists/(r*r)
  int compositor;
  float bg[4];
  const float* static_rgb;
  int32_t* idx;
  float* zbuf;
  float* dists;
  float* image;
  float* mask;
};

constexpr float kInf = __builtin_huge_valf();

// candidate (z, idx) strictly before list element (ze, slot se)?  Total order (z, idx).
__device__ __forceinline__ bool cand_less(float z, int idx, float ze, int se,
                                          const float4* __restrict__ recA) {
  bool lt = z < ze;
  if (z == ze) {  // exact fp32 z tie -> smaller packed index first
    lt = (se < 0) ? true : (idx < __float_as_int(recA[kRecStride * se].w));
  }
  return lt;
}

template <int KP>
struct KList {
  float z[KP];
  int s[KP];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < KP; ++i) {
      z[i] = kInf;
      s[i] = -1;
    }
  }
  // Fast path: strict-< compare-exchange chain (5 ALU ops per slot, no branches, no loads).
  // Returns true if an exact z tie was involved in a way that could change the result:
  //   - the candidate ties with the current last element (it may have to displace it), or
  //   - the evicted element ties with the new last element (the wrong twin may have left).
  // Ties that stay inside the list are caught by has_adjacent_tie() at the end.
  __device__ __forceinline__ bool insert_fast(float cz, int cs) {
    bool tie = (cz == z[KP - 1]);
    if (cz < z[KP - 1]) {
#pragma unroll
      for (int i = 0; i < KP; ++i) {
        const bool p = cz < z[i];
        const float tz = z[i];
        const int ts = s[i];
        z[i] = p ? cz : tz;
        s[i] = p ? cs : ts;
        cz = p ? tz : cz;
        cs = p ? ts : cs;
      }
      tie = tie || (cs >= 0 && cz == z[KP - 1]);
    }
    return tie;
  }
  __device__ __forceinline__ bool has_adjacent_tie() const {
    bool t = false;
#pragma unroll
    for (int i = 1; i < KP; ++i) t = t || (s[i] >= 0 && z[i] == z[i - 1]);
    return t;
  }
  // Exact path: full (z, idx) order.
  __device__ __forceinline__ void insert_exact(float cz, int cidx, int cslot,
                                               const float4* __restrict__ recA) {
    if (!cand_less(cz, cidx, z[KP - 1], s[KP - 1], recA)) return;
    bool p_hi = true;
#pragma unroll
    for (int i = KP - 1; i > 0; --i) {
      const bool p_lo = cand_less(cz, cidx, z[i - 1], s[i - 1], recA);
      z[i] = p_lo ? z[i - 1] : (p_hi ? cz : z[i]);
      s[i] = p_lo ? s[i - 1] : (p_hi ? cslot : s[i]);
      p_hi = p_lo;
    }
    if (p_hi) {
      z[0] = cz;
      s[0] = cslot;
    }
  }
};

struct PixelCtx {
  float xf, yf, r2;
};

// PPR = per-point radius (recB.w); the scalar-radius instantiation carries no extra load/branch
template <bool PPR>
__device__ __forceinline__ bool hit_test(const PixelCtx& c, const float4 a,
                                         const float4* __restrict__ recB, int j) {
  const float d2 = dist2_rn(a.x, a.y, c.xf, c.yf);
  float r2 = c.r2;
  if (PPR) {
    const float r = __ldg(&recB[kRecStride * j].w);
    r2 = __fmul_rn(r, r);
  }
  return d2 < r2;
}

// Tie-aware rescan of one pixel (rare).  Kept out of line so the fast path stays small.
template <int KP, bool PPR>
__device__ __noinline__ void rescan_exact(const RasterParams& p, const PixelCtx& c, int n, int x,
                                          int y, float* zout, int* sout) {
  KList<KP> q;
  q.init();
  const int span = 2 * p.halo + 1;
  for (int ry = 0; ry < span; ++ry) {
    const int64_t cell0 = ((int64_t)n * p.GH + (y + ry)) * p.GW + x;
    const int s = __ldg(p.cell_end + cell0 - 1);
    const int e = __ldg(p.cell_end + cell0 + span - 1);
    for (int j = s; j < e; ++j) {
      const float4 a = __ldg(p.recA + kRecStride * j);
      if (hit_test<PPR>(c, a, p.recB, j)) q.insert_exact(a.z, __float_as_int(a.w), j, p.recA);
    }
  }
#pragma unroll
  for (int i = 0; i < KP; ++i) {
    zout[i] = q.z[i];
    sout[i] = q.s[i];
  }
}

#ifndef PGDVS_RASTER_MINBLOCKS
#define PGDVS_RASTER_MINBLOCKS 1
#endif
#ifndef PGDVS_RASTER_QUEUE
#define PGDVS_RASTER_QUEUE 16
#endif
constexpr int kQueueCap = PGDVS_RASTER_QUEUE;  // hits buffered per pixel between two drains

// Two-phase pixel loop.  Testing a candidate is cheap (~15 instructions) but inserting a hit
// into the z-sorted list is a KP-slot compare-exchange chain, and in SIMT a chain costs the same
// whether 3 or 32 lanes need it.  So hits are first appended to a lane-private column of a
// shared-memory queue (phase 1, no chain in the loop); then the warp drains the queues
// (phase 2): the chain now runs max_lanes(#hits) times with most lanes active instead of once
// per candidate iteration with ~40 % of the lanes.  Columns are private to a thread, so no
// synchronisation is needed; a full column triggers a drain for the whole warp.
template <int KP, bool PPR>
__global__ void __launch_bounds__(256, PGDVS_RASTER_MINBLOCKS) k_raster_cells(const __grid_constant__ RasterParams p) {
  // Measured on B200 (C2 workload): the queue's own overhead (vote + smem round trip per
  // candidate) outweighs the better chain utilisation — 3.34 ms queued vs 2.80 ms direct — so the
  // direct path is the default; the queued path is kept behind a build flag for dense / large-K
  // experiments.
#ifdef PGDVS_RASTER_USE_QUEUE
  constexpr bool QUEUED = (KP >= 4) && (KP <= 64);
#else
  constexpr bool QUEUED = false;
#endif
  __shared__ float s_qz[QUEUED ? kQueueCap : 1][256];
  __shared__ int s_qs[QUEUED ? kQueueCap : 1][256];
  const unsigned full = 0xffffffffu;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  const int n = blockIdx.z;
  // out-of-image lanes stay alive (warp votes below) but scan nothing and store nothing
  const bool inside = (x < p.W) && (y < p.H);
  PixelCtx c;
  c.xf = pixel_center_ndc(p.ax, x);
  c.yf = pixel_center_ndc(p.ay, y);
  c.r2 = p.r2;
  const float4* __restrict__ recA = p.recA;
  const float4* __restrict__ recB = p.recB;

  KList<KP> q;
  q.init();
  bool tie = false;
  int cnt = 0;

  auto drain = [&]() {
    const int nmax = __reduce_max_sync(full, cnt);
    for (int i = 0; i < nmax; ++i) {
      if (i < cnt) tie |= q.insert_fast(s_qz[i][tid], s_qs[i][tid]);
    }
    cnt = 0;
  };
  // phase-1 body for one candidate (record a at slot j)
  auto consider = [&](bool live, const float4& a, int j) {
    bool hit = live && hit_test<PPR>(c, a, recB, j);
    if (QUEUED) {
      if (hit) {
        // early reject against the current K-th nearest; an exact tie is left to the slow path
        tie |= (a.z == q.z[KP - 1]);
        hit = a.z < q.z[KP - 1];
      }
      if (hit) {
        s_qz[cnt][tid] = a.z;
        s_qs[cnt][tid] = j;
        ++cnt;
      }
      if (__any_sync(full, cnt == kQueueCap)) drain();
    } else {
      if (hit) tie |= q.insert_fast(a.z, j);
    }
  };

  const int span = 2 * p.halo + 1;
  // cs[c] = start of cell (x + c) of the first window row = cell_end[... - 1]
  const int* __restrict__ cs = p.cell_end + ((int64_t)n * p.GH + (inside ? y : 0)) * p.GW + (inside ? x : 0) - 1;
  if (p.halo == 1) {
    // 3x3 cell window (every PGDVS configuration with r_px < 1.5): the three row runs are
    // walked by ONE flattened loop so that lanes with uneven rows do not wait for each other
    // three times.
    const int s0 = __ldg(cs), e0 = __ldg(cs + 3);
    const int s1 = __ldg(cs + p.GW), e1 = __ldg(cs + p.GW + 3);
    const int s2 = __ldg(cs + 2 * p.GW), e2 = __ldg(cs + 2 * p.GW + 3);
    const int c0 = e0 - s0, c01 = c0 + (e1 - s1);
    const int total = inside ? c01 + (e2 - s2) : 0;
    const int o1 = s1 - c0, o2 = s2 - c01;
    const int tmax = __reduce_max_sync(full, total);  // warp-uniform trip count
    // software-pipelined: the record of iteration t+1 is in flight while t is processed
    int j = (0 < c0 ? s0 : (0 < c01 ? o1 : o2));
    float4 a = (total > 0) ? __ldg(recA + kRecStride * j) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < tmax; ++t) {
      const int tn = t + 1;
      const int jn = tn + (tn < c0 ? s0 : (tn < c01 ? o1 : o2));
      float4 an = a;
      if (tn < total) an = __ldg(recA + kRecStride * jn);
      consider(t < total, a, j);
      a = an;
      j = jn;
    }
  } else {
    for (int ry = 0; ry < span; ++ry) {
      const int s = __ldg(cs + (int64_t)ry * p.GW);
      const int len = inside ? __ldg(cs + (int64_t)ry * p.GW + span) - s : 0;
      const int lmax = __reduce_max_sync(full, len);
      for (int i = 0; i < lmax; ++i) {
        const bool live = i < len;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) a = __ldg(recA + kRecStride * (s + i));
        consider(live, a, s + i);
      }
    }
  }
  if (QUEUED) drain();
  if (!inside) return;
  if (tie || q.has_adjacent_tie()) {
    float zt[KP];
    int st[KP];
    rescan_exact<KP, PPR>(p, c, n, x, y, zt, st);
#pragma unroll
    for (int i = 0; i < KP; ++i) {
      q.z[i] = zt[i];
      q.s[i] = st[i];
    }
  }

  // ---------------------------------------------------------------- epilogue
  const int K = p.K;
  const int64_t pix = ((int64_t)n * p.H + y) * p.W + x;
  const int mode = p.compositor;
  // pass 1: fragments (idx / zbuf / dists) and the PointsRenderer weights
  float w[KP];
  float t_alpha = 0.f;
  const bool vec4 = (KP % 4 == 0) && (K == KP);  // K*4 B per pixel is a multiple of 16 B
#pragma unroll
  for (int k0 = 0; k0 < KP; k0 += 4) {
    int o_idx[4];
    float o_z[4], o_d[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int k = k0 + kk;
      o_idx[kk] = -1;
      o_z[kk] = -1.0f;
      o_d[kk] = -1.0f;
      if (k < KP) {
        w[k] = 0.f;
        const int sl = (k < K) ? q.s[k] : -1;
        if (sl >= 0) {
          const float4 a = __ldg(recA + kRecStride * sl);
          o_d[kk] = dist2_rn(a.x, a.y, c.xf, c.yf);
          o_idx[kk] = __float_as_int(a.w);
          o_z[kk] = a.z;
          if (mode != PGDVS_COMPOSITE_NONE) {
            w[k] = __fsub_rn(1.0f, __fdiv_rn(o_d[kk], p.rr_weight));  // 1 - dists/(r*r)
            t_alpha = __fadd_rn(t_alpha, w[k]);
          }
        }
      }
    }
    if (vec4) {
      const int64_t o = pix * K + k0;
      if (p.idx) *reinterpret_cast<int4*>(p.idx + o) = make_int4(o_idx[0], o_idx[1], o_idx[2], o_idx[3]);
      if (p.zbuf) *reinterpret_cast<float4*>(p.zbuf + o) = make_float4(o_z[0], o_z[1], o_z[2], o_z[3]);
      if (p.dists) *reinterpret_cast<float4*>(p.dists + o) = make_float4(o_d[0], o_d[1], o_d[2], o_d[3]);
    } else {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = k0 + kk;
        if (k < K) {
          if (p.idx) p.idx[pix * K + k] = o_idx[kk];
          if (p.zbuf) p.zbuf[pix * K + k] = o_z[kk];
          if (p.dists) p.dists[pix * K + k] = o_d[kk];
        }
      }
    }
  }
  if (mode == PGDVS_COMPOSITE_NONE) return;

  // pass 2: compositor.  The K-ordered accumulation follows the pytorch3d CPU loops; the
  // norm-weighted division by max(sum w, 1e-4) is applied as one reciprocal multiply
  // (images are tolerance-matched, |delta| <= 1e-5; fragments above are bit-exact).
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float ones_acc = 0.f;  // the same compositor applied to all-ones features (mask render)
  float cum_alpha = 1.0f;
  const float inv_t = __frcp_rn(fmaxf(t_alpha, 1e-4f));
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    if (k < K && q.s[k] >= 0) {
      const float4 f4 = __ldg(recB + kRecStride * q.s[k]);
      float wk = w[k];
      if (mode == PGDVS_COMPOSITE_NORM_WEIGHTED) {
        wk = __fmul_rn(wk, inv_t);
      } else if (mode == PGDVS_COMPOSITE_ALPHA) {
        const float a = wk;
        wk = __fmul_rn(cum_alpha, a);
        cum_alpha = __fmul_rn(cum_alpha, __fsub_rn(1.0f, a));
      }
      acc[0] = __fadd_rn(acc[0], __fmul_rn(wk, f4.x));
      acc[1] = __fadd_rn(acc[1], __fmul_rn(wk, f4.y));
      acc[2] = __fadd_rn(acc[2], __fmul_rn(wk, f4.z));
      acc[3] = __fadd_rn(acc[3], __fmul_rn(wk, f4.w));
      ones_acc = __fadd_rn(ones_acc, wk);
    }
  }
  const bool is_bg = q.s[0] < 0;  // _add_background_color_to_images: idx[:, 0] < 0
  const float m = (ones_acc > 0.0f) ? 1.0f : 0.0f;
  if (p.mask) p.mask[pix] = m;
  if (p.image) {
    for (int ch = 0; ch < p.C; ++ch) {
      float v = is_bg ? p.bg[ch] : acc[ch];
      if (p.static_rgb) {
        // combined = (1 - mask) * static + mask * dyn   (pgdvs_renderer.py:169-172)
        const float st = __ldg(p.static_rgb + pix * p.C + ch);
        v = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, m), st), __fmul_rn(m, v));
      }
      p.image[pix * p.C + ch] = v;
    }
  }
}

template <int KP>
static int launch_raster(const RasterParams& p, cudaStream_t stream) {
  dim3 block(32, 8);
  dim3 grid((p.W + 31) / 32, (p.H + 7) / 8, p.N);
  if (p.r2 < 0.0f)
    k_raster_cells<KP, true><<<grid, block, 0, stream>>>(p);
  else
    k_raster_cells<KP, false><<<grid, block, 0, stream>>>(p);
  return check_launch();
}

}  // namespace pgdvs

using namespace pgdvs;

extern "C" int pgdvs_rasterize_composite(const void* workspace, size_t workspace_bytes, int N,
                                         int64_t P, int H, int W, int K, float radius_max,
                                         int per_point_radius, int C, int compositor,
                                         float rr_weight, const float* background,
                                         const float* static_rgb, int32_t* idx, float* zbuf,
                                         float* dists, float* image, float* mask, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (workspace == nullptr || N < 0 || H <= 0 || W <= 0 || P < 0 || K < 1 ||
      !(radius_max >= 0.0f))
    return PGDVS_E_BADARG;
  if (K > PGDVS_MAX_POINTS_PER_PIXEL) return PGDVS_E_K_TOO_LARGE;
  if (compositor < PGDVS_COMPOSITE_NONE || compositor > PGDVS_COMPOSITE_WEIGHTED_SUM)
    return PGDVS_E_BADARG;
  if (compositor != PGDVS_COMPOSITE_NONE) {
    if (C < 1 || C > PGDVS_MAX_FUSED_CHANNELS) return PGDVS_E_CHANNELS;
    if (!(rr_weight > 0.0f)) return PGDVS_E_BADARG;
    if (static_rgb != nullptr && (image == nullptr)) return PGDVS_E_BADARG;
  }
  BinLayout L = make_bin_layout(N, H, W, P, radius_max);
  if (workspace_bytes < L.total) return PGDVS_E_WORKSPACE;
  if (N == 0) return PGDVS_OK;

  const char* ws = static_cast<const char*>(workspace);
  RasterParams p;
  p.cell_end = reinterpret_cast<const int*>(ws + L.off_cells);
  p.recA = reinterpret_cast<const float4*>(ws + L.off_recA);
  p.recB = reinterpret_cast<const float4*>(ws + L.off_recB);
  p.N = N;
  p.H = H;
  p.W = W;
  p.K = K;
  p.C = (compositor == PGDVS_COMPOSITE_NONE) ? 0 : C;
  p.halo = L.halo;
  p.GW = L.GW;
  p.GH = L.GH;
  p.ax = make_ndc_axis(W, H);
  p.ay = make_ndc_axis(H, W);
  // fp32 r*r, as `radius2 = radius * radius` upstream; negative selects the per-point path
  p.r2 = per_point_radius ? -1.0f : radius_max * radius_max;
  p.rr_weight = rr_weight;
  p.compositor = compositor;
  for (int c = 0; c < 4; ++c) p.bg[c] = (background != nullptr && c < C) ? background[c] : 0.0f;
  p.static_rgb = static_rgb;
  p.idx = idx;
  p.zbuf = zbuf;
  p.dists = dists;
  p.image = image;
  p.mask = mask;

  if (K <= 1) return launch_raster<1>(p, stream);
  if (K <= 2) return launch_raster<2>(p, stream);
  if (K <= 3) return launch_raster<3>(p, stream);
  if (K <= 4) return launch_raster<4>(p, stream);
  if (K <= 8) return launch_raster<8>(p, stream);
  if (K <= 16) return launch_raster<16>(p, stream);
  if (K <= 32) return launch_raster<32>(p, stream);
  if (K <= 64) return launch_raster<64>(p, stream);
  return launch_raster<PGDVS_MAX_POINTS_PER_PIXEL>(p, stream);
}
