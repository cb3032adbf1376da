// Shared device/host helpers for the PGDVS point-splat kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/pgdvs_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "pgdvs_b200 kernels are written for sm_100a (B200) only"
#endif

namespace pgdvs {

// ---------------------------------------------------------------------------------------
// Pixel-centre NDC coordinates, bit-identical to pytorch3d's PixToNonSquareNdc
// (rasterization_utils.cuh; restated in oracle/raster_cpu.cpp).  Every operation is an
// explicitly rounded fp32 op so that nvcc cannot contract a*b+c into an FMA.
// ---------------------------------------------------------------------------------------
struct NdcAxis {
  float range;   // NonSquareNdcRange(S1, S2)
  float offset;  // range / 2
  int S1;
};

inline NdcAxis make_ndc_axis(int S1, int S2) {
  NdcAxis a;
  float range = 2.0f;
  if (S1 > S2) range = ((float)S1 * range) / (float)S2;
  a.range = range;
  a.offset = range / 2.0f;
  a.S1 = S1;
  return a;
}

// centre of pixel `pix` (0 = left/top of the OUTPUT image; pytorch3d flips both axes)
__device__ __forceinline__ float pixel_center_ndc(const NdcAxis& a, int pix) {
  const int i = a.S1 - 1 - pix;
  const float t = __fadd_rn(__fmul_rn(a.range, (float)i), a.offset);
  return __fadd_rn(-a.offset, __fdiv_rn(t, (float)a.S1));
}

// squared NDC distance exactly as the x86 build of pytorch3d evaluates it:
// dx*dx + dy*dy with two roundings (no FMA)
__device__ __forceinline__ float dist2_rn(float px, float py, float xf, float yf) {
  const float dx = __fsub_rn(px, xf);
  const float dy = __fsub_rn(py, yf);
  return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

// ---------------------------------------------------------------------------------------
// Workspace layout of the binning pass (all offsets 256-byte aligned).
//
// The image is covered by 1-pixel cells; the grid is extended by `halo` cells on every
// side so that a pixel's candidate window [x-halo, x+halo] never needs clamping.  A point
// is filed under the cell of its NEAREST pixel centre, so every pixel it can hit lies
// within halo = floor(r_px + 0.5 + 1/64) cells of that cell (r_px = radius in pixels; the
// 1/64 absorbs the fp32 error of the cell computation for images up to 16k pixels wide).
// Cells are numbered view-major, row-major: one image row of a tile (plus halo) is ONE
// contiguous run of the cell-sorted record arrays — that is what lets the rasterizer
// fetch tile rows with 1-D bulk copies.
// ---------------------------------------------------------------------------------------
struct BinLayout {
  int halo, GW, GH;
  int64_t cells;      // N * GH * GW
  int64_t n_tiles;    // scan tiles
  // int32 [cells + 1], preceded by a zeroed 256-byte pad so that element [-1] reads 0.
  // Life cycle: per-cell counts -> (k_scan) exclusive starts -> (fill: atomicAdd cursor)
  // per-cell ENDS.  The rasterizer therefore reads start(c) = cell_end[c - 1].
  size_t off_cells;
  size_t off_state;   // uint64 [n_tiles]    decoupled look-back state
  size_t off_ticket;  // int32 [4]           scan ticket counter (+pad)
  size_t off_zrange;  // uint32 [2 N]        per view: max(~bits(z)), max(bits(z)) over the filed points
                      //                     (zero-initialised == empty); sizes the rasterizer's z keys
  size_t off_zero_end;  // everything in [0, off_zero_end) is zeroed before each binning
  size_t off_cell_of; // int32  [P]          cell of every packed point (-1 = never rasterized);
                      //                     staged path only (the fused path carries it in its records)
  size_t off_zmin;    // float [cells + pad] smallest z of every cell (NaN = empty), written by the
                      //                     rasterizer's k_sort_cells pass when it z-sorts the cells
  size_t off_recA;    // 32-byte records [P], cell-sorted: part A (x_ndc, y_ndc, z, packed idx as
  size_t off_recB;    // int bits) and part B (features C<=4, or f0,f1,f2,radius) at rec_a(j) / rec_b(j)
  size_t total;
};

constexpr int kScanTile = 4096;  // ints per scan tile (1024 threads x int4)
constexpr int kUwpTilePixels = 1024;  // source pixels per tile of the uwp pass (uwp.cu); k_fill_pre_tiles moves one per CTA
#ifndef PGDVS_REC_STRIDE
#define PGDVS_REC_STRIDE 2
#endif
// 2: recA/recB interleaved as one 32-byte record per point; 1: two separate float4 arrays
constexpr int kRecStride = PGDVS_REC_STRIDE;

// Record j is 32 bytes: part A (x_ndc, y_ndc, z, packed idx) and part B (features / radius).
// The two 16-byte halves swap places every 4 records: rec_a / rec_b give the float4 index of a
// part relative to the record array.  A warp that reads the A parts of arbitrary records with
// 128-bit shared-memory loads then spreads over all 8 four-bank groups instead of 4 (a plain
// 32-byte stride would leave the odd groups to the B parts only).  The staging copy keeps
// (shared slot - global slot) a multiple of 8 so both sides agree on which half is which.
__host__ __device__ __forceinline__ int rec_a(int j) {
#ifdef __CUDA_ARCH__
  // spelled in PTX (3 instructions): the C++ form below is "simplified" into twice as many
  unsigned r;
  asm("{\n\t.reg .b32 t;\n\t"
      "shr.u32 t, %1, 2;\n\t"
      "and.b32 t, t, 1;\n\t"
      "mad.lo.u32 %0, %1, 2, t;\n\t}"
      : "=r"(r)
      : "r"(j));
  return (int)r;
#else
  const unsigned u = (unsigned)j;
  return (int)((u << 1) | ((u >> 2) & 1u));
#endif
}
__host__ __device__ __forceinline__ int rec_b(int j) { return rec_a(j) ^ 1; }
static_assert(PGDVS_REC_STRIDE == 2, "records are interleaved 32-byte pairs");
constexpr int64_t kMaxRecords = (int64_t)1 << 30;  // rec_a / rec_b return int32 float4 indices

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

inline int halo_cells(float radius_max, int H, int W) {
  const float s = 0.5f * (float)(H < W ? H : W);
  const float r_px = fabsf(radius_max) * s;
  return (int)floorf(r_px + 0.5f + 1.0f / 64.0f);
}

inline BinLayout make_bin_layout(int N, int H, int W, int64_t P, float radius_max) {
  BinLayout L;
  L.halo = halo_cells(radius_max, H, W);
  L.GW = W + 2 * L.halo;
  L.GH = H + 2 * L.halo;
  L.cells = (int64_t)N * L.GH * L.GW;
  L.n_tiles = (L.cells + 1 + kScanTile - 1) / kScanTile;
  size_t o = 256;  // zero pad in front of the cell array (cell_end[-1] == 0)
  L.off_cells = o;
  o = align256(o + sizeof(int32_t) * (size_t)(L.n_tiles * kScanTile));
  L.off_state = o;
  o = align256(o + sizeof(uint64_t) * (size_t)L.n_tiles);
  L.off_ticket = o;
  o = align256(o + 256);
  L.off_zrange = o;
  o = align256(o + sizeof(uint32_t) * 2 * (size_t)(N > 0 ? N : 1));
  L.off_zero_end = o;
  L.off_cell_of = o;
  o = align256(o + sizeof(int32_t) * (size_t)(P > 0 ? P : 1));
  L.off_zmin = o;
  o = align256(o + sizeof(float) * (size_t)(L.n_tiles * kScanTile));
  // records of one point sit next to each other (one full 32 B sector per point): the fill
  // pass's scatter then never leaves half-written sectors behind
  L.off_recA = o;
  L.off_recB = o + (kRecStride == 2 ? sizeof(float4) : align256(sizeof(float4) * (size_t)(P > 0 ? P : 1)));
  o = align256(o + 2 * align256(sizeof(float4) * (size_t)(P > 0 ? P : 1)));
  L.total = o;
  return L;
}

// The fused path (uwp kernel counts cells itself) parks its not-yet-sorted records behind
// the regular layout, so the rasterizer sees the same front part in both modes.
struct FusedTail {
  size_t off_preA;  // float4 [P]  (x_ndc, y_ndc, z, cell id) in packed (reference) order
  size_t off_preB;  // 3 floats per point (r, g, b), unpadded (the area is sized for float4)
  size_t total;
};

inline FusedTail make_fused_tail(const BinLayout& L, int64_t P) {
  FusedTail T;
  size_t o = L.total;
  T.off_preA = o;
  o = align256(o + sizeof(float4) * (size_t)(P > 0 ? P : 1));
  T.off_preB = o;
  o = align256(o + sizeof(float4) * (size_t)(P > 0 ? P : 1));
  T.total = o;
  return T;
}

// Geometry of the extended 1-pixel cell grid (see BinLayout).
struct CellGrid {
  int H, W, halo, GW, GH;
  float xf0, yf0;  // NDC centre of output column 0 / row 0
  float inv_pix;   // pixels per NDC unit = min(H,W)/2
};

inline CellGrid make_cell_grid(int H, int W, int halo) {
  CellGrid g;
  g.H = H;
  g.W = W;
  g.halo = halo;
  g.GW = W + 2 * halo;
  g.GH = H + 2 * halo;
  const NdcAxis ax = make_ndc_axis(W, H), ay = make_ndc_axis(H, W);
  // centre of output column 0 is PixToNonSquareNdc(W-1, W, H) (host evaluation, fp32)
  g.xf0 = -ax.offset + (ax.range * (float)(W - 1) + ax.offset) / (float)W;
  g.yf0 = -ay.offset + (ay.range * (float)(H - 1) + ay.offset) / (float)H;
  g.inv_pix = 0.5f * (float)(H < W ? H : W);
  return g;
}

// cell of the nearest pixel centre, in extended-grid coordinates of view n; -1 if the point
// can never be rasterized (behind the camera, or further than `halo` cells outside the image)
__device__ __forceinline__ int point_cell(const CellGrid& g, int n, float x, float y, float z) {
  if (z < 0.0f) return -1;  // pytorch3d: `if (pz < 0) continue;`
  // pixel centres: xf(col) = xf0 - col / inv_pix  ->  col = (xf0 - x) * inv_pix
  const float colf = (g.xf0 - x) * g.inv_pix;
  const float rowf = (g.yf0 - y) * g.inv_pix;
  // NaN / huge coordinates fail these comparisons and are dropped (they can never satisfy
  // dist2 < r2 either)
  if (!(colf > -(float)g.halo - 1.0f && colf < (float)(g.W + g.halo))) return -1;
  if (!(rowf > -(float)g.halo - 1.0f && rowf < (float)(g.H + g.halo))) return -1;
  const int gx = __float2int_rn(colf) + g.halo;
  const int gy = __float2int_rn(rowf) + g.halo;
  if (gx < 0 || gx >= g.GW || gy < 0 || gy >= g.GH) return -1;
  return (n * g.GH + gy) * g.GW + gx;
}

// Range of the z bit patterns of the points filed under view n (sizes the rasterizer's sorted
// 32-bit keys, raster.cu KeyCode).  z >= 0 for every filed point, so the pattern orders like the
// value; -0.0 counts as +0.0.  Stored as (max ~bits, max bits) so that zero means "empty".
// Every lane named in `mask` calls this with its own partial range (nlo = ~min pattern, or 0;
// hi = max pattern, or 0): one pair of atomics per warp.
__device__ __forceinline__ uint32_t z_pattern(float z) { return __float_as_uint(__fadd_rn(z, 0.0f)); }
__device__ __forceinline__ void zrange_accumulate(uint32_t* zrange, int n, unsigned mask, uint32_t nlo,
                                                  uint32_t hi) {
  nlo = __reduce_max_sync(mask, nlo);
  hi = __reduce_max_sync(mask, hi);
  if ((threadIdx.x & 31) == (__ffs(mask) - 1) && (nlo | hi) != 0u) {
    // thousands of warps update the same two words of a view: look first (an L2 read), and only
    // the few warps that really widen the range pay for an atomic
    const uint2 cur = __ldcg(reinterpret_cast<const uint2*>(zrange) + n);
    if (nlo > cur.x) atomicMax(zrange + 2 * n, nlo);
    if (hi > cur.y) atomicMax(zrange + 2 * n + 1, hi);
  }
}

// PointsRasterizer.transform for PerspectiveCameras(in_ndc=True):
//   view = p @ R + T ; x = (fx*X + px*Z)/Z ; y = (fy*Y + py*Z)/Z ; z = Z (view depth)
__device__ __forceinline__ float3 world_to_ndc(const PgdvsCamera& c, float wx, float wy, float wz) {
  const float X = wx * c.R[0] + wy * c.R[3] + wz * c.R[6] + c.T[0];
  const float Y = wx * c.R[1] + wy * c.R[4] + wz * c.R[7] + c.T[1];
  const float Z = wx * c.R[2] + wy * c.R[5] + wz * c.R[8] + c.T[2];
  float3 o;
  o.x = (c.focal[0] * X + c.p0[0] * Z) / Z;
  o.y = (c.focal[1] * Y + c.p0[1] * Z) / Z;
  o.z = Z;
  return o;
}

inline int check_launch() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace pgdvs
