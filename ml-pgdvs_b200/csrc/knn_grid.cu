// Uniform-grid K nearest neighbours for the statistical outlier filter (exact).
//
// The reference calls `pytorch3d.ops.knn_points(pcl, pcl, K=51)` — brute force, O(P^2) — on every
// source pair with the filter enabled (pgdvs_renderer_dyn.py:405-427; ON in every published
// benchmark).  At P = 156 672 (one 288x544 frame) the brute-force kernel of
// knn.cu needs 58 ms on a B200, 2 000x the rest of the per-view path.  This file answers the
// same question with a counting-sorted uniform grid:
//
//   k_knn_bbox   bounding box of the reference cloud (warp shuffles + ordered-int atomics)
//   k_knn_setup  one thread: cell size from the cloud's footprint (depth-map clouds are
//                surfaces: ~4 points per occupied cell), grid dims capped at 4 M cells
//   k_knn_count  cell of every reference point + per-cell counts (RED)
//   k_scan       (bin.cu) exclusive scan of the counters
//   k_knn_fill   points scattered into cell order as float4 (x, y, z, original index)
//   k_knn_query  one thread per query walks cubic shells of cells around its own cell — each
//                (dz, dy) row of a shell face is ONE contiguous run of the sorted array — keeps
//                the K smallest squared distances sorted in registers, and stops as soon as the
//                K-th is no farther than the nearest unvisited cell can be.  Exact, not
//                approximate: the stopping bound is the true distance to the shell boundary
//                (sides that coincide with the grid boundary are unbounded).
//
// Everything stays on the device (the grid geometry is a device struct), no host sync.
#include "common.cuh"

#include <stdlib.h>

namespace pgdvs {

int scan_exclusive_inplace(int* data, int64_t n_tiles, unsigned long long* state, int* ticket,
                           cudaStream_t stream);

__host__ __device__ __forceinline__ float kInfF() {
#ifdef __CUDA_ARCH__
  return __int_as_float(0x7f800000);
#else
  return __builtin_huge_valf();
#endif
}

constexpr int kGridMaxCells = 1 << 22;
constexpr int kGridMaxK = 64;
constexpr float kGridTargetPerCell = 4.0f;
constexpr int kGridSamples = 128;  // sample queries that calibrate the cell size
constexpr float kGridCellScale = 1.0f;

struct KnnGrid {
  float lo[3];
  float h, inv_h;
  int n[3];
};

struct KnnGridLayout {
  size_t off_grid, off_bbox, off_est, off_cells, off_state, off_ticket, off_zero_end, off_cell_of, off_sorted, total;
  int64_t scan_tiles;
};

static inline KnnGridLayout make_knn_grid_layout(int64_t R) {
  KnnGridLayout L;
  size_t o = 0;
  L.off_grid = o;
  o += 256;
  L.off_bbox = o;
  o += 256;
  L.off_est = o;
  o += sizeof(float) * kGridSamples;
  o = align256(o);
  o += 256;  // zero pad in front of the cells: cells[-1] == 0
  L.off_cells = o;
  L.scan_tiles = ((int64_t)kGridMaxCells + 1 + kScanTile - 1) / kScanTile;
  o = align256(o + sizeof(int) * (size_t)(L.scan_tiles * kScanTile));
  L.off_state = o;
  o = align256(o + sizeof(unsigned long long) * (size_t)L.scan_tiles);
  L.off_ticket = o;
  o = align256(o + 256);
  L.off_zero_end = o;
  L.off_cell_of = o;
  o = align256(o + sizeof(int) * (size_t)(R > 0 ? R : 1));
  L.off_sorted = o;
  o = align256(o + sizeof(float4) * (size_t)(R > 0 ? R : 1));
  L.total = o;
  return L;
}

// order-preserving float <-> int encoding for atomicMin/Max
__device__ __forceinline__ int float_to_ordered(float f) {
  const int i = __float_as_int(f);
  return (i >= 0) ? i : (i ^ 0x7fffffff);
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float((i >= 0) ? i : (i ^ 0x7fffffff)); }

__global__ void __launch_bounds__(256) k_fill_f32(float* out, int64_t n, float v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = v;
}

__global__ void k_knn_bbox_init(int* bbox) {
  if (threadIdx.x < 3) bbox[threadIdx.x] = 0x7fffffff;        // mins
  else if (threadIdx.x < 6) bbox[threadIdx.x] = (int)0x80000000;  // maxs
}

__global__ void __launch_bounds__(256) k_knn_bbox(const float* __restrict__ ref, int64_t R, int* bbox) {
  float lo[3] = {kInfF(), kInfF(), kInfF()}, hi[3] = {-kInfF(), -kInfF(), -kInfF()};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < R; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float v = __ldg(ref + i * 3 + a);
      if (v == v && fabsf(v) < 1e30f) {  // NaN / inf points do not stretch the grid
        lo[a] = fminf(lo[a], v);
        hi[a] = fmaxf(hi[a], v);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], d));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], d));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(bbox + a, float_to_ordered(lo[a]));
      atomicMax(bbox + 3 + a, float_to_ordered(hi[a]));
    }
  }
}

// Calibration: for 128 sample points, a histogram of log2(squared distance) to ALL reference
// points gives the radius that holds K + 1 of them (within a factor sqrt(2)).  The median of
// those radii becomes the cell size, so a typical query is done after the first or second shell
// whatever the local shape of the cloud (thin surface, rough surface, volume).
// Large clouds are read with a stride (every `stride`-th reference point, the count threshold scaled with
// it): the estimate only has to land in the right power-of-two bin.
__global__ void __launch_bounds__(256) k_knn_sample(const float* __restrict__ ref, int64_t R, int K, int stride,
                                                    float* __restrict__ est) {
  __shared__ int s_hist[256];  // bin = biased exponent of d2 (0..255)
  s_hist[threadIdx.x] = 0;
  __syncthreads();
  const int64_t si = (int64_t)blockIdx.x * R / kGridSamples;
  const float qx = __ldg(ref + si * 3), qy = __ldg(ref + si * 3 + 1), qz = __ldg(ref + si * 3 + 2);
  for (int64_t i = (int64_t)threadIdx.x * stride; i < R; i += (int64_t)blockDim.x * stride) {
    const float dx = qx - __ldg(ref + i * 3), dy = qy - __ldg(ref + i * 3 + 1), dz = qz - __ldg(ref + i * 3 + 2);
    const float d = dx * dx + dy * dy + dz * dz;
    if (d == d) atomicAdd(&s_hist[(__float_as_uint(d) >> 23) & 255], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int cum = 0, bin = 255;
    for (int b = 0; b < 256; ++b) {
      cum += s_hist[b];
      if (cum * stride >= K + 1) {
        bin = b;
        break;
      }
    }
    // upper edge of the bin: d2 < 2^(bin - 126)
    est[blockIdx.x] = (bin >= 254 || !(qx == qx && qy == qy && qz == qz)) ? 0.0f
                                                                           : sqrtf(__uint_as_float((unsigned)(bin + 1) << 23));
  }
}

// one CTA of kGridSamples threads: the median of the calibrated radii by rank counting (every thread
// ranks its own sample), then thread 0 sizes the grid
__global__ void __launch_bounds__(kGridSamples) k_knn_setup(const int* bbox, int64_t R, const float* est, float cell_scale,
                                                           KnnGrid* g) {
  __shared__ float s_v[kGridSamples];
  __shared__ float s_median;
  {
    const int t = threadIdx.x;
    const float v = est[t];
    s_v[t] = v;
    if (t == 0) s_median = 0.0f;
    const int m = __syncthreads_count(v > 0.0f);  // zeros = unusable samples
    if (v > 0.0f) {
      int rank = 0;
      for (int j = 0; j < kGridSamples; ++j) {
        const float w = s_v[j];
        rank += (w > 0.0f && (w < v || (w == v && j < t))) ? 1 : 0;
      }
      if (rank == m / 2) s_median = v;  // element m / 2 of the ascending order
    }
    __syncthreads();
  }
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float lo[3], e[3];
  for (int a = 0; a < 3; ++a) {
    lo[a] = ordered_to_float(bbox[a]);
    const float hi = ordered_to_float(bbox[3 + a]);
    e[a] = (hi >= lo[a]) ? (hi - lo[a]) : 0.0f;
    if (!(lo[a] == lo[a]) || fabsf(lo[a]) > 1e30f) {  // empty / degenerate cloud
      lo[a] = 0.0f;
      e[a] = 0.0f;
    }
  }
  // footprint = the largest face of the box (depth-map clouds are height fields over it)
  const float emax = fmaxf(e[0], fmaxf(e[1], e[2]));
  const float emin = fminf(e[0], fminf(e[1], e[2]));
  const float emid = e[0] + e[1] + e[2] - emax - emin;
  float area = emax * emid;
  if (!(area > 0.0f)) area = emax * emax;
  float h = sqrtf(kGridTargetPerCell * area / (float)(R > 0 ? R : 1));  // fallback: thin-surface model
  if (!(h > 0.0f)) h = 1.0f;
  if (s_median > 0.0f) h = s_median * cell_scale;  // median of the calibrated K-neighbourhood radii, scaled
  int n[3];
  for (int it = 0; it < 64; ++it) {
    double cells = 1.0;
    for (int a = 0; a < 3; ++a) {
      const float f = floorf(e[a] / h) + 1.0f;
      n[a] = (int)fminf(f, 4096.0f);
      cells *= (double)n[a];
    }
    if (cells <= (double)kGridMaxCells && n[0] < 4096 && n[1] < 4096 && n[2] < 4096) break;
    h *= 1.2599f;
  }
  for (int a = 0; a < 3; ++a) {
    g->lo[a] = lo[a];
    g->n[a] = n[a];
  }
  g->h = h;
  g->inv_h = 1.0f / h;
}

__device__ __forceinline__ int3 grid_coord(const KnnGrid& g, float x, float y, float z) {
  int3 c;
  c.x = min(max((int)floorf((x - g.lo[0]) * g.inv_h), 0), g.n[0] - 1);
  c.y = min(max((int)floorf((y - g.lo[1]) * g.inv_h), 0), g.n[1] - 1);
  c.z = min(max((int)floorf((z - g.lo[2]) * g.inv_h), 0), g.n[2] - 1);
  return c;
}
__device__ __forceinline__ int grid_cell(const KnnGrid& g, int3 c) { return (c.z * g.n[1] + c.y) * g.n[0] + c.x; }

__global__ void __launch_bounds__(256) k_knn_count(const float* __restrict__ ref, int64_t R,
                                                   const KnnGrid* __restrict__ gp, int* cells, int* cell_of) {
  const KnnGrid g = *gp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < R; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = __ldg(ref + i * 3), y = __ldg(ref + i * 3 + 1), z = __ldg(ref + i * 3 + 2);
    int cell = -1;
    if (x == x && y == y && z == z) {  // NaN points are never anybody's neighbour (d2 = NaN fails `<`)
      cell = grid_cell(g, grid_coord(g, x, y, z));
      atomicAdd(cells + cell, 1);
    }
    cell_of[i] = cell;
  }
}

__global__ void __launch_bounds__(256) k_knn_fill(const float* __restrict__ ref, int64_t R, int* cells,
                                                  const int* __restrict__ cell_of, float4* sorted) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < R; i += (int64_t)gridDim.x * blockDim.x) {
    const int cell = cell_of[i];
    if (cell < 0) continue;
    const int pos = atomicAdd(cells + cell, 1);  // starts -> ends
    sorted[pos] = make_float4(__ldg(ref + i * 3), __ldg(ref + i * 3 + 1), __ldg(ref + i * 3 + 2),
                              __int_as_float((int)i));
  }
}

// SELF: queries are the reference points themselves, processed in cell order (coherent warps).
// KT > 0: K is the compile-time constant KT (the list is exactly K registers, kept sorted with a
// min/max chain, its last element is the K-th distance) — instantiated for the reference's default
// dyn_pcl_outlier_knn = 50 (K = 51); KT = 0: any K <= 64 at run time.
template <bool SELF, int KT>
__global__ void __launch_bounds__(128, (KT > 0 && KT <= 52) ? 5 : 1) k_knn_query(const float* __restrict__ query, int64_t Q,
                                                   const KnnGrid* __restrict__ gp, const int* __restrict__ cell_end,
                                                   const float4* __restrict__ sorted, int64_t R_sorted_hint,
                                                   int K, int skip, float* __restrict__ mean_out) {
  const KnnGrid g = *gp;
  const int64_t qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= Q) return;
  float qx, qy, qz;
  int64_t out_i = qi;
  if (SELF) {
    // sorted holds only the finite points: the tail of an array with NaN points is unused
    const int total_sorted = __ldg(cell_end + (int64_t)g.n[0] * g.n[1] * g.n[2] - 1);
    if (qi >= total_sorted) return;
    const float4 p = __ldg(sorted + qi);
    qx = p.x;
    qy = p.y;
    qz = p.z;
    out_i = __float_as_int(p.w);
  } else {
    qx = __ldg(query + qi * 3);
    qy = __ldg(query + qi * 3 + 1);
    qz = __ldg(query + qi * 3 + 2);
  }
  (void)R_sorted_hint;
  constexpr int NB = (KT > 0) ? KT : kGridMaxK;
  float best[NB];
#pragma unroll
  for (int i = 0; i < NB; ++i) best[i] = kInfF();
  float kth = kInfF();
  const int3 c = grid_coord(g, qx, qy, qz);
  const int rmax = max(g.n[0], max(g.n[1], g.n[2]));
  for (int r = 0; r <= rmax; ++r) {
    const int z0 = max(c.z - r, 0), z1 = min(c.z + r, g.n[2] - 1);
    const int y0 = max(c.y - r, 0), y1 = min(c.y + r, g.n[1] - 1);
    const int x0 = max(c.x - r, 0), x1 = min(c.x + r, g.n[0] - 1);
    for (int z = z0; z <= z1; ++z) {
      const bool zface = (z == c.z - r) || (z == c.z + r);
      for (int y = y0; y <= y1; ++y) {
        const bool face = zface || (y == c.y - r) || (y == c.y + r);
        const int row = (z * g.n[1] + y) * g.n[0];
        // a face row of the shell: the whole x-run; an interior row: only its two end cells
        for (int part = 0; part < (face ? 1 : 2); ++part) {
          int xa, xb;
          if (face) {
            xa = x0;
            xb = x1;
          } else {
            xa = xb = (part == 0) ? c.x - r : c.x + r;
            if (xa < 0 || xa >= g.n[0] || (part == 1 && r == 0)) continue;
          }
          const int s = __ldg(cell_end + row + xa - 1);
          const int e = __ldg(cell_end + row + xb);
          for (int j = s; j < e; ++j) {
            const float4 p = __ldg(sorted + j);
            const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
            const float d = dx * dx + dy * dy + dz * dz;
            if (d < kth) {
              float cd = d;
              if (KT > 0) {
#pragma unroll
                for (int t = 0; t < NB; ++t) {  // sorted insertion: two min/max per slot
                  const float b = best[t];
                  best[t] = fminf(b, cd);
                  cd = fmaxf(b, cd);
                }
                kth = best[NB - 1];
              } else {
#pragma unroll
                for (int t = 0; t < NB; ++t) {
                  if (t < K) {
                    const float b = best[t];
                    const bool lt = cd < b;
                    best[t] = lt ? cd : b;
                    cd = lt ? b : cd;
                  }
                }
                float k2 = best[0];
#pragma unroll
                for (int t = 1; t < NB; ++t)
                  if (t == K - 1) k2 = best[t];
                kth = k2;
              }
            }
          }
        }
      }
    }
    // every unvisited point lies outside the cube of cells [c - r, c + r]; sides of the cube that
    // reach the grid boundary have nothing behind them
    float bound = kInfF();
    const float q[3] = {qx, qy, qz};
    const int cc[3] = {c.x, c.y, c.z};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (cc[a] - r > 0) bound = fminf(bound, q[a] - (g.lo[a] + (float)(cc[a] - r) * g.h));
      if (cc[a] + r < g.n[a] - 1) bound = fminf(bound, (g.lo[a] + (float)(cc[a] + r + 1) * g.h) - q[a]);
    }
    if (bound == kInfF()) break;                 // the cube covers the whole grid
    bound = fmaxf(bound - 1e-4f * g.h, 0.0f);    // slack for the rounding of the cell assignment
    if (kth <= bound * bound) break;
  }
  const int kk = K;  // (the caller guarantees R >= K on this path)
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < NB; ++t)
    if (t >= skip && t < kk && best[t] < kInfF()) sum += best[t];
  int cnt = 0;
#pragma unroll
  for (int t = 0; t < NB; ++t)
    if (t >= skip && t < kk && best[t] < kInfF()) ++cnt;
  mean_out[out_i] = cnt > 0 ? sum / (float)cnt : 0.f;
}

// ---------------------------------------------------------------------------------------
// Few queries against a large cloud (the track branch: a few hundred track points against the
// closest-pair cloud, pgdvs_renderer_dyn_track.py "track2base" filter): one WARP per query.
// A query that lies off the surface walks r shells of mostly empty cells before it meets K points —
// O(r^3) cells, 9.3 ms per call with one thread per query.  Here the lanes share that walk:
//   * the points inside the cube of cells [c - r, c + r] are COUNTED from the cell table alone (every
//     (z, y) row of the cube is one run: two table reads), rows spread over the lanes, and r grows
//     until the cube holds K points;
//   * their squared distances go to a per-warp shared-memory array (or are re-read from the sorted
//     array when the cube holds more than kWarpCap of them), and the K-th smallest is found by
//     bisection on the bit pattern (d2 >= 0: patterns order like values): 31 counting passes;
//   * if the K-th lies within the distance to the cube's boundary the K nearest are all inside and
//     mean = (sum of d2 below the K-th + (K - their number) * K-th - the `skip` smallest) / (K - skip);
//     otherwise the cube grows by exactly the missing distance and the step repeats.
// Exact like k_knn_query (same stopping bound); the sum is formed in lane order + shuffle tree
// (deterministic), not in ascending order, so it can differ from the thread kernel in the last ulps.
// ---------------------------------------------------------------------------------------
constexpr int kWarpCap = 2048;       // cached squared distances per warp
constexpr int kWarpQueryWarps = 4;   // warps (queries) per CTA
constexpr int kCollectUnroll = 4;    // points per lane and trip of the candidate collection
constexpr int64_t kWarpQueryMaxQ = (int64_t)1 << 20;  // cross queries (query != reference)
#ifndef PGDVS_KNN_WARP_SELF_MAXQ
#define PGDVS_KNN_WARP_SELF_MAXQ (1 << 20)
#endif
constexpr int64_t kWarpSelfMaxQ = PGDVS_KNN_WARP_SELF_MAXQ;  // self queries

#ifdef PGDVS_KNN_STATS
__device__ unsigned long long g_knn_wstats[8];
#endif

struct KnnCube {
  int x0, x1, y0, y1, z0, z1, ny, nrows;
  float bound;      // every point outside the cube is at least this far from the query
  bool covers_all;  // nothing lies outside
};

__device__ __forceinline__ KnnCube knn_cube(const KnnGrid& g, const int3 c, const int r, const float qx, const float qy,
                                            const float qz) {
  KnnCube C;
  C.z0 = max(c.z - r, 0), C.z1 = min(c.z + r, g.n[2] - 1);
  C.y0 = max(c.y - r, 0), C.y1 = min(c.y + r, g.n[1] - 1);
  C.x0 = max(c.x - r, 0), C.x1 = min(c.x + r, g.n[0] - 1);
  C.ny = C.y1 - C.y0 + 1;
  C.nrows = (C.z1 - C.z0 + 1) * C.ny;
  float bound = kInfF();
  const float q[3] = {qx, qy, qz};
  const int cc[3] = {c.x, c.y, c.z};
#pragma unroll
  for (int a = 0; a < 3; ++a) {  // (sides on the grid boundary are open)
    if (cc[a] - r > 0) bound = fminf(bound, q[a] - (g.lo[a] + (float)(cc[a] - r) * g.h));
    if (cc[a] + r < g.n[a] - 1) bound = fminf(bound, (g.lo[a] + (float)(cc[a] + r + 1) * g.h) - q[a]);
  }
  C.covers_all = (bound == kInfF());
  C.bound = fmaxf(bound - 1e-4f * g.h, 0.0f);  // slack for the rounding of the cell assignment
  return C;
}

// SELF: the queries are the reference points themselves, taken in cell order from the sorted array
template <bool SELF>
__global__ void __launch_bounds__(32 * kWarpQueryWarps) k_knn_query_warp(
    const float* __restrict__ query, int64_t Q, const KnnGrid* __restrict__ gp, const int* __restrict__ cell_end,
    const float4* __restrict__ sorted, int K, int skip, float* __restrict__ mean_out) {
  __shared__ float s_d2[kWarpQueryWarps][kWarpCap];
  __shared__ int s_hist[kWarpQueryWarps][256];
  __shared__ float s_sel[kWarpQueryWarps][32];
  const KnnGrid g = *gp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t qi = (int64_t)blockIdx.x * kWarpQueryWarps + warp;
  if (qi >= Q) return;
  float* cache = s_d2[warp];
  int* hist = s_hist[warp];
  float* selbuf = s_sel[warp];
  float qx, qy, qz;
  if (SELF) {
    // sorted holds only the finite points: the tail of an array with NaN points is unused
    const int total_sorted = __ldg(cell_end + (int64_t)g.n[0] * g.n[1] * g.n[2] - 1);
    if (qi >= total_sorted) return;
    const float4 p = __ldg(sorted + qi);
    qx = p.x;
    qy = p.y;
    qz = p.z;
    qi = __float_as_int(p.w);
  } else {
    qx = __ldg(query + qi * 3);
    qy = __ldg(query + qi * 3 + 1);
    qz = __ldg(query + qi * 3 + 2);
  }
  if (!(qx == qx && qy == qy && qz == qz)) {  // NaN query: no distance compares below anything (k_knn_query: 0)
    if (lane == 0) mean_out[qi] = 0.f;
    return;
  }
  const int3 c = grid_coord(g, qx, qy, qz);
  const int rmax = max(g.n[0], max(g.n[1], g.n[2]));

  // f(d2) over every point of cube C, each point visited by exactly one lane (rows spread over lanes)
  auto for_each_point = [&](const KnnCube& C, auto&& f) {
    for (int i = lane; i < C.nrows; i += 32) {
      const int z = C.z0 + i / C.ny, y = C.y0 + i % C.ny;
      const int row = (z * g.n[1] + y) * g.n[0];
      const int s = __ldg(cell_end + row + C.x0 - 1), e = __ldg(cell_end + row + C.x1);
      for (int j = s; j < e; ++j) {
        const float4 p = __ldg(sorted + j);
        const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
        f(dx * dx + dy * dy + dz * dz);
      }
    }
  };
  auto warp_sum = [](int v) { return __reduce_add_sync(0xffffffffu, v); };

  // ---- A: the first cube (geometric growth) that holds K points, counted from the table alone
  int r = 0;
  KnnCube C;
  int cnt;
  while (true) {
    C = knn_cube(g, c, r, qx, qy, qz);
    int n = 0;
    for (int i = lane; i < C.nrows; i += 32) {
      const int z = C.z0 + i / C.ny, y = C.y0 + i % C.ny;
      const int row = (z * g.n[1] + y) * g.n[0];
      n += __ldg(cell_end + row + C.x1) - __ldg(cell_end + row + C.x0 - 1);
    }
    cnt = warp_sum(n);
    if (cnt >= K || C.covers_all || r >= rmax) break;
    r += max(1, r >> 1);
  }
  const int kk = min(K, cnt);
  if (kk <= skip) {
    if (lane == 0) mean_out[qi] = 0.f;
    return;
  }

  // the candidates of cube C whose d2 <= limit, appended to the cache by warp-prefix offsets; returns
  // how many there are (more than kWarpCap: the cache is not valid)
  // The candidates of cube C whose d2 <= limit, appended to the cache; returns how many there are (more
  // than kWarpCap: the cache is not valid).  Rows are taken 32 at a time (one per lane); their POINTS are
  // then spread over the lanes — point t of the chunk belongs to the first row whose inclusive count
  // exceeds t, found by a 5-step shuffle search — so that every lane has one independent load in flight
  // per trip (a lane walking its own row alone waits for each of its ~15 loads in turn: 25 - 50 k cycles
  // per query), and survivors are compacted with a ballot.
  auto collect = [&](const KnnCube& Cc, const float limit) -> int {
    __syncwarp();  // (the reads of an earlier selection pass are done before the cache is rewritten)
    int base = 0;
    for (int i0 = 0; i0 < Cc.nrows; i0 += 32) {
      const int i = i0 + lane;
      int s = 0, e = 0;
      if (i < Cc.nrows) {
        const int z = Cc.z0 + i / Cc.ny, y = Cc.y0 + i % Cc.ny;
        const int row = (z * g.n[1] + y) * g.n[0];
        s = __ldg(cell_end + row + Cc.x0 - 1);
        e = __ldg(cell_end + row + Cc.x1);
      }
      const int len = e - s;
      if (!__any_sync(0xffffffffu, len > 0)) continue;  // 32 empty rows
      int inc = len;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
      }
      const int T = __shfl_sync(0xffffffffu, inc, 31);
      // kCollectUnroll points per lane and trip: their loads are issued back to back
      for (int t0 = 0; t0 < T; t0 += 32 * kCollectUnroll) {
        float d2[kCollectUnroll];
        bool keep[kCollectUnroll];
#pragma unroll
        for (int u = 0; u < kCollectUnroll; ++u) {
          const int t = t0 + 32 * u + lane, tt = min(t, T - 1);
          int lo = 0, hi = 31;  // the smallest lane whose inclusive count exceeds tt
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            const int mid = (lo + hi) >> 1;
            const int v = __shfl_sync(0xffffffffu, inc, mid);
            if (v > tt)
              hi = mid;
            else
              lo = mid + 1;
          }
          const int inc_o = __shfl_sync(0xffffffffu, inc, lo), len_o = __shfl_sync(0xffffffffu, len, lo);
          const int s_o = __shfl_sync(0xffffffffu, s, lo);
          keep[u] = t < T;
          d2[u] = 0.f;
          if (keep[u]) {
            const float4 p = __ldg(sorted + s_o + (tt - (inc_o - len_o)));
            const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
            d2[u] = dx * dx + dy * dy + dz * dz;
          }
        }
#pragma unroll
        for (int u = 0; u < kCollectUnroll; ++u) {
          const bool k = keep[u] && d2[u] <= limit;
          const unsigned m = __ballot_sync(0xffffffffu, k);
          const int dst = base + __popc(m & ((1u << lane) - 1u));
          if (k && dst < kWarpCap) cache[dst] = d2[u];
          base += __popc(m);
        }
      }
    }
    __syncwarp();
    return base;
  };
  // n-th smallest (1-based) of the candidate set: the least pattern T with #{d2 <= T} >= n.
  // d2 >= 0, so the bit patterns order like the values.
  auto nth_smallest = [&](const KnnCube& Cc, const int n_cand, const bool cached, const float limit, const int n) -> float {
    uint32_t lo = 0u, hi = 0x7f800000u;
    if (cached) {
      // four-way search: three thresholds per pass over the cache, 16 dependent passes instead of 31
      while (lo < hi) {
        const uint32_t span = hi - lo;
        const uint32_t q2 = lo + (span >> 1), q1 = lo + (span >> 2), q3 = q2 + (span >> 2);  // lo <= q1 <= q2 <= q3 < hi
        int c1 = 0, c2 = 0, c3 = 0;
        for (int i = lane; i < n_cand; i += 32) {
          const uint32_t v = __float_as_uint(cache[i]);
          c1 += (v <= q1) ? 1 : 0;
          c2 += (v <= q2) ? 1 : 0;
          c3 += (v <= q3) ? 1 : 0;
        }
        c1 = warp_sum(c1);
        c2 = warp_sum(c2);
        c3 = warp_sum(c3);
        if (c1 >= n) {
          hi = q1;
        } else if (c2 >= n) {
          lo = q1 + 1u;
          hi = q2;
        } else if (c3 >= n) {
          lo = q2 + 1u;
          hi = q3;
        } else {
          lo = q3 + 1u;
        }
      }
      return __uint_as_float(lo);
    }
    while (lo < hi) {  // (the cache overflowed: bisection over the cube's points themselves)
      const uint32_t mid = lo + ((hi - lo) >> 1);
      int n_le = 0;
      for_each_point(Cc, [&](float d) { n_le += (d <= limit && __float_as_uint(d) <= mid) ? 1 : 0; });
      n_le = warp_sum(n_le);
      if (n_le >= n)
        hi = mid;
      else
        lo = mid + 1u;
    }
    return __uint_as_float(lo);
  };

  // The n smallest of the cached candidates by COUNTING: a 256-bin histogram over [0, max d2] (shared-memory
  // atomics), the bin b* of the n-th by a scan of the bins (8 per lane), the sum of everything below b*, and
  // the at most 32 candidates inside b* ranked against each other with shuffles.  ~350 warp instructions
  // where the four-way search over all candidates takes ~2 000.  ok = false (b* holds more than 32
  // candidates: clouds with many duplicates): the caller falls back to the search.
  struct Sel {
    bool ok;
    float kth, sum, dmin;
  };
  auto select_hist = [&](const int n_cand, const int n) -> Sel {
    float dmax = 0.f, dmin = kInfF();
    for (int i = lane; i < n_cand; i += 32) {
      const float v = cache[i];
      dmax = fmaxf(dmax, v);
      dmin = fminf(dmin, v);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, d));
      dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, d));
    }
    if (!(dmax > 0.f)) return Sel{true, 0.f, 0.f, dmin};  // every candidate at distance 0
    const float scale = 255.0f / dmax;
    auto bin_of = [&](float v) { return min((int)(v * scale), 255); };  // monotone in v
#pragma unroll
    for (int k = 0; k < 8; ++k) hist[lane + 32 * k] = 0;
    __syncwarp();
    for (int i = lane; i < n_cand; i += 32) atomicAdd(&hist[bin_of(cache[i])], 1);
    __syncwarp();
    int h[8], tot = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      h[k] = hist[lane * 8 + k];
      tot += h[k];
    }
    int inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    const int excl = inc - tot;
    const bool mine = (excl < n) && (n <= inc);  // exactly one lane: n <= n_cand
    int bstar = 0, below = 0, m = 0;
    if (mine) {
      int cum = excl;
      bool found = false;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (!found && cum + h[k] >= n) {
          bstar = lane * 8 + k;
          below = cum;
          m = h[k];
          found = true;
        }
        cum += h[k];
      }
    }
    const unsigned owner = __ballot_sync(0xffffffffu, mine);
    if (owner == 0u) return Sel{false, 0.f, 0.f, dmin};  // (cannot happen for n <= n_cand)
    const int src = __ffs(owner) - 1;
    bstar = __shfl_sync(0xffffffffu, bstar, src);
    below = __shfl_sync(0xffffffffu, below, src);
    m = __shfl_sync(0xffffffffu, m, src);
    if (m > 32) return Sel{false, 0.f, 0.f, dmin};
    // sum below b*, and the candidates of b* one per lane
    float sum = 0.f;
    int filled = 0;
    for (int i0 = 0; i0 < n_cand; i0 += 32) {
      const int i = i0 + lane;
      const bool valid = i < n_cand;
      const float v = valid ? cache[i] : 0.f;
      const int b = bin_of(v);
      if (valid && b < bstar) sum += v;
      const bool in = valid && b == bstar;
      const unsigned msk = __ballot_sync(0xffffffffu, in);
      if (in) selbuf[filled + __popc(msk & ((1u << lane) - 1u))] = v;
      filled += __popc(msk);
    }
    __syncwarp();
    const float mv = (lane < m) ? selbuf[lane] : kInfF();
    int rank = 0;
    for (int j = 0; j < m; ++j) {
      const float vj = __shfl_sync(0xffffffffu, mv, j);
      rank += (vj < mv || (vj == mv && j < lane)) ? 1 : 0;
    }
    const int need = n - below;  // 1 .. m
    const unsigned kmask = __ballot_sync(0xffffffffu, lane < m && rank == need - 1);
    const float kth = __shfl_sync(0xffffffffu, mv, __ffs(kmask) - 1);
    if (lane < m && rank < need) sum += mv;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    __syncwarp();  // (selbuf / hist are reused by the next selection)
    return Sel{true, kth, sum, dmin};
  };

  // ---- B: the kk-th smallest of cube A; it bounds the true kk-th from above
  int n_cand = collect(C, kInfF());
  bool cached = n_cand <= kWarpCap;
  float limit = kInfF();
  Sel sel = cached ? select_hist(n_cand, kk) : Sel{false, 0.f, 0.f, 0.f};
  float kth = sel.ok ? sel.kth : nth_smallest(C, n_cand, cached, limit, kk);
  if (!C.covers_all && !(kth <= C.bound * C.bound)) {
    // ---- C: the cube whose boundary is at least sqrt(kth) away holds every point with d2 <= kth,
    //      the true kk nearest among them: collect only those
    const float miss = sqrtf(kth) - C.bound;
    r = min(r + max(1, (int)ceilf(miss * g.inv_h)), rmax);
    C = knn_cube(g, c, r, qx, qy, qz);
    if (!C.covers_all && !(kth <= C.bound * C.bound)) {  // (rounding of the step: one more shell)
      r = min(r + 1, rmax);
      C = knn_cube(g, c, r, qx, qy, qz);
    }
    limit = kth;
    n_cand = collect(C, limit);
    cached = n_cand <= kWarpCap;
    sel = cached ? select_hist(n_cand, kk) : Sel{false, 0.f, 0.f, 0.f};
    kth = sel.ok ? sel.kth : nth_smallest(C, n_cand, cached, limit, kk);
  }
  // ---- mean of the kk smallest without the `skip` smallest
  float sum = 0.f, dmin = kInfF();
  if (sel.ok) {
    sum = sel.sum;
    dmin = sel.dmin;
  } else {
    int n_lt = 0;
    auto acc = [&](float d) {
      if (d < kth) {
        sum += d;
        ++n_lt;
      }
      dmin = fminf(dmin, d);
    };
    if (cached) {
      for (int i = lane; i < n_cand; i += 32) acc(cache[i]);
    } else {
      for_each_point(C, [&](float d) { if (d <= limit) acc(d); });
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      sum += __shfl_xor_sync(0xffffffffu, sum, d);
      dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, d));
    }
    n_lt = warp_sum(n_lt);
    sum += (float)(kk - n_lt) * kth;
  }
  if (skip == 1) sum -= dmin;
  if (lane == 0) mean_out[qi] = sum / (float)(kk - skip);
#ifdef PGDVS_KNN_STATS
  if (lane == 0) {
    atomicAdd(&g_knn_wstats[0], 1ull);
    atomicAdd(&g_knn_wstats[1], (unsigned long long)r);
    atomicAdd(&g_knn_wstats[2], (unsigned long long)C.nrows);
    atomicAdd(&g_knn_wstats[3], (unsigned long long)n_cand);
    if (r > 8) atomicAdd(&g_knn_wstats[4], 1ull);
    if (r > 32) atomicAdd(&g_knn_wstats[5], 1ull);
    if (!cached) atomicAdd(&g_knn_wstats[6], 1ull);
    atomicMax(&g_knn_wstats[7], (unsigned long long)C.nrows);
  }
#endif
}

// Cell size relative to the calibrated K-neighbourhood radius.  Developer override, read once:
// PGDVS_KNN_CELL_SCALE=<float>.
static float knn_cell_scale() {
  static const float v = [] {
    const char* e = getenv("PGDVS_KNN_CELL_SCALE");
    const float f = (e && *e) ? (float)atof(e) : kGridCellScale;
    return (f > 0.01f && f < 100.0f) ? f : kGridCellScale;
  }();
  return v;
}

// Host driver used by pgdvs_knn_mean_dist (knn.cu) when the caller provides the workspace.
size_t knn_grid_workspace_bytes(int64_t R) { return make_knn_grid_layout(R).total; }

int knn_grid_mean_dist(const float* query, int64_t Q, const float* ref, int64_t R, int K, int skip,
                       float* mean_out, void* workspace, cudaStream_t stream) {
  const KnnGridLayout L = make_knn_grid_layout(R);
  char* ws = static_cast<char*>(workspace);
  KnnGrid* grid = reinterpret_cast<KnnGrid*>(ws + L.off_grid);
  int* bbox = reinterpret_cast<int*>(ws + L.off_bbox);
  int* cells = reinterpret_cast<int*>(ws + L.off_cells);
  int* cell_of = reinterpret_cast<int*>(ws + L.off_cell_of);
  float4* sorted = reinterpret_cast<float4*>(ws + L.off_sorted);
  cudaError_t e = cudaMemsetAsync(ws, 0, L.off_zero_end, stream);
  if (e != cudaSuccess) return (int)e;
  const int blocks = (int)((R + 255) / 256 < 148 * 8 ? (R + 255) / 256 : 148 * 8);
  k_knn_bbox_init<<<1, 32, 0, stream>>>(bbox);
  k_knn_bbox<<<blocks, 256, 0, stream>>>(ref, R, bbox);
  float* est = reinterpret_cast<float*>(ws + L.off_est);
  k_knn_sample<<<kGridSamples, 256, 0, stream>>>(ref, R, K, R > 32768 ? 4 : 1, est);
  k_knn_setup<<<1, kGridSamples, 0, stream>>>(bbox, R, est, knn_cell_scale(), grid);
  k_knn_count<<<blocks, 256, 0, stream>>>(ref, R, grid, cells, cell_of);
  if (int rc = check_launch()) return rc;
  if (int rc = scan_exclusive_inplace(cells, L.scan_tiles, reinterpret_cast<unsigned long long*>(ws + L.off_state),
                                      reinterpret_cast<int*>(ws + L.off_ticket), stream))
    return rc;
  k_knn_fill<<<blocks, 256, 0, stream>>>(ref, R, cells, cell_of, sorted);
  const unsigned qb = (unsigned)((Q + 127) / 128);
  if (query == ref && Q == R) {
    // points with NaN coordinates are not in the sorted array: like the brute-force kernel they
    // get +inf (no finite neighbour distance)
    k_fill_f32<<<blocks, 256, 0, stream>>>(mean_out, Q, kInfF());
    // the warp kernel (count-then-select) against the sorted-list thread kernel, grid build included:
    // 20 k points 0.20 vs 0.47 ms, 60 k 0.47 vs 0.79 ms, 157 k 1.13 vs 1.19 ms (before the lane-spread
    // candidate collection and the four-way search it lost above 60 k: 1.77 ms at 157 k)
    if (Q <= kWarpSelfMaxQ && skip <= 1) {
      const unsigned wb = (unsigned)((Q + kWarpQueryWarps - 1) / kWarpQueryWarps);
      k_knn_query_warp<true><<<wb, 32 * kWarpQueryWarps, 0, stream>>>(query, Q, grid, cells, sorted, K, skip, mean_out);
    } else if (K == 51)
      k_knn_query<true, 51><<<qb, 128, 0, stream>>>(query, Q, grid, cells, sorted, R, K, skip, mean_out);
    else
      k_knn_query<true, 0><<<qb, 128, 0, stream>>>(query, Q, grid, cells, sorted, R, K, skip, mean_out);
  } else if (Q <= kWarpQueryMaxQ && skip <= 1) {
    const unsigned wb = (unsigned)((Q + kWarpQueryWarps - 1) / kWarpQueryWarps);
    k_knn_query_warp<false><<<wb, 32 * kWarpQueryWarps, 0, stream>>>(query, Q, grid, cells, sorted, K, skip, mean_out);
  } else {
    if (K == 51)
      k_knn_query<false, 51><<<qb, 128, 0, stream>>>(query, Q, grid, cells, sorted, R, K, skip, mean_out);
    else
      k_knn_query<false, 0><<<qb, 128, 0, stream>>>(query, Q, grid, cells, sorted, R, K, skip, mean_out);
  }
  return check_launch();
}

}  // namespace pgdvs

#ifdef PGDVS_KNN_STATS
extern "C" int pgdvs_debug_knn_wstats(unsigned long long* out8) {
  cudaError_t e = cudaMemcpyFromSymbol(out8, pgdvs::g_knn_wstats, sizeof(unsigned long long) * 8);
  if (e != cudaSuccess) return (int)e;
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  return (int)cudaMemcpyToSymbol(pgdvs::g_knn_wstats, z, sizeof(z));
}
#endif
