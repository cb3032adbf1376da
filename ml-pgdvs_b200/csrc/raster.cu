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
//   * the hit test reproduces the reference arithmetic bit for bit (dist2_rn, strict <),
//     ties in z are broken by the smaller packed index exactly like the CPU rasterizer's
//     (z, idx, dist2) priority queue, so idx/zbuf/dists are deterministic and bit-exact.
#include "common.cuh"

namespace pgdvs {

struct RasterParams {
  const int* cell_start;
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

// candidate (z, idx) strictly before list element (ze, slot se)?  Total order (z, idx).
__device__ __forceinline__ bool cand_less(float z, int idx, float ze, int se,
                                          const float4* __restrict__ recA) {
  bool lt = z < ze;
  if (z == ze) {  // rare: exact fp32 z tie -> smaller packed index first
    lt = (se < 0) ? true : (idx < __float_as_int(recA[se].w));
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
      z[i] = __int_as_float(0x7f800000);  // +inf
      s[i] = -1;
    }
  }
  __device__ __forceinline__ void insert(float cz, int cidx, int cslot,
                                         const float4* __restrict__ recA) {
    if (!cand_less(cz, cidx, z[KP - 1], s[KP - 1], recA)) return;
    bool p_hi = true;  // cand < element i (known true for i = KP-1)
#pragma unroll
    for (int i = KP - 1; i > 0; --i) {
      const bool p_lo = cand_less(cz, cidx, z[i - 1], s[i - 1], recA);
      // new[i] = p_lo ? old[i-1] : (p_hi ? cand : old[i])
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

template <int KP>
__global__ void __launch_bounds__(256) k_raster_cells(const __grid_constant__ RasterParams p) {
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  const int n = blockIdx.z;
  if (x >= p.W || y >= p.H) return;
  const float xf = pixel_center_ndc(p.ax, x);
  const float yf = pixel_center_ndc(p.ay, y);
  const float4* __restrict__ recA = p.recA;
  const float4* __restrict__ recB = p.recB;
  const bool per_point_r = p.r2 < 0.0f;

  KList<KP> q;
  q.init();

  const int span = 2 * p.halo + 1;
  for (int ry = 0; ry < span; ++ry) {
    // extended-grid row (y + ry), cells [x, x + 2*halo] <-> image cells [x-halo, x+halo]
    const int64_t cell0 = ((int64_t)n * p.GH + (y + ry)) * p.GW + x;
    const int s = __ldg(p.cell_start + cell0);
    const int e = __ldg(p.cell_start + cell0 + span);
    for (int j = s; j < e; ++j) {
      const float4 a = __ldg(recA + j);
      const float d2 = dist2_rn(a.x, a.y, xf, yf);
      float r2 = p.r2;
      if (per_point_r) {
        const float r = __ldg(&recB[j].w);
        r2 = __fmul_rn(r, r);
      }
      if (d2 < r2) q.insert(a.z, __float_as_int(a.w), j, recA);
    }
  }

  // ---------------------------------------------------------------- epilogue
  const int K = p.K;
  const int64_t pix = ((int64_t)n * p.H + y) * p.W + x;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float ones_acc = 0.f;  // the same compositor applied to all-ones features (mask render)
  const int mode = p.compositor;
  float t_alpha = 0.f;
  if (mode == PGDVS_COMPOSITE_NORM_WEIGHTED) {
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      if (k < K && q.s[k] >= 0) {
        const float4 a = __ldg(recA + q.s[k]);
        const float d2 = dist2_rn(a.x, a.y, xf, yf);
        const float w = __fsub_rn(1.0f, __fdiv_rn(d2, p.rr_weight));
        t_alpha = __fadd_rn(t_alpha, w);
      }
    }
    t_alpha = fmaxf(t_alpha, 1e-4f);  // kEps of norm_weighted_sum
  }
  float cum_alpha = 1.0f;
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    if (k < K) {
      const int sl = q.s[k];
      int o_idx = -1;
      float o_z = -1.0f, o_d = -1.0f;
      if (sl >= 0) {
        const float4 a = __ldg(recA + sl);
        const float d2 = dist2_rn(a.x, a.y, xf, yf);
        o_idx = __float_as_int(a.w);
        o_z = a.z;
        o_d = d2;
        if (mode != PGDVS_COMPOSITE_NONE) {
          const float w = __fsub_rn(1.0f, __fdiv_rn(d2, p.rr_weight));
          const float4 f4 = __ldg(recB + sl);
          const float f[4] = {f4.x, f4.y, f4.z, f4.w};
          if (mode == PGDVS_COMPOSITE_NORM_WEIGHTED) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              acc[c] = __fadd_rn(acc[c], __fdiv_rn(__fmul_rn(w, f[c]), t_alpha));
            ones_acc = __fadd_rn(ones_acc, __fdiv_rn(w, t_alpha));
          } else if (mode == PGDVS_COMPOSITE_ALPHA) {
            const float cw = __fmul_rn(cum_alpha, w);
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[c] = __fadd_rn(acc[c], __fmul_rn(cw, f[c]));
            ones_acc = __fadd_rn(ones_acc, cw);
            cum_alpha = __fmul_rn(cum_alpha, __fsub_rn(1.0f, w));
          } else {  // weighted sum
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[c] = __fadd_rn(acc[c], __fmul_rn(w, f[c]));
            ones_acc = __fadd_rn(ones_acc, w);
          }
        }
      }
      if (p.idx) p.idx[pix * K + k] = o_idx;
      if (p.zbuf) p.zbuf[pix * K + k] = o_z;
      if (p.dists) p.dists[pix * K + k] = o_d;
    }
  }
  if (mode != PGDVS_COMPOSITE_NONE) {
    const bool is_bg = q.s[0] < 0;  // _add_background_color_to_images: idx[:, 0] < 0
    const float m = (ones_acc > 0.0f) ? 1.0f : 0.0f;
    if (p.mask) p.mask[pix] = m;
    if (p.image) {
      for (int c = 0; c < p.C; ++c) {
        float v = is_bg ? p.bg[c] : acc[c];
        if (p.static_rgb) {
          // combined = (1 - mask) * static + mask * dyn   (pgdvs_renderer.py:169-172)
          const float st = __ldg(p.static_rgb + pix * p.C + c);
          v = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, m), st), __fmul_rn(m, v));
        }
        p.image[pix * p.C + c] = v;
      }
    }
  }
}

template <int KP>
static int launch_raster(const RasterParams& p, cudaStream_t stream) {
  dim3 block(32, 8);
  dim3 grid((p.W + 31) / 32, (p.H + 7) / 8, p.N);
  k_raster_cells<KP><<<grid, block, 0, stream>>>(p);
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
  p.cell_start = reinterpret_cast<const int*>(ws + L.off_start);
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
