// Binning pass: exact count -> scan -> fill cell sort of a packed batch of NDC point clouds.
//
// The reference forces bin_size=0 (pgdvs_renderer_dyn.py:689-695) because pytorch3d's
// coarse binning has a fixed max_points_per_bin that overflows on real clouds; here the
// per-cell lists are sized exactly by a counting sort, so nothing can overflow.
//
//   k_count : one fire-and-forget atomic (RED) per point on its cell counter; the cell id is
//             kept so the fill pass does not redo the projection-to-cell arithmetic.
//   k_scan  : single-pass chained scan (decoupled look-back) over the cell counters,
//             warp-shuffle prefix sums inside a tile, int4-vectorised loads/stores.
//   k_fill  : pos = atomicAdd(cursor[cell], 1) on the scanned array itself, then scatter
//             (x,y,z,idx) and the feature record.  The cursor array ends up holding per-cell
//             END offsets; the rasterizer reads start(c) = end(c-1).  Four points per thread
//             are kept in flight because the cell -> cursor -> scatter chain is latency-bound.
#include "common.cuh"

#include <stdlib.h>

namespace pgdvs {

struct BinParams {
  const float* points;
  const float* features;
  const float* radius;
  const int64_t* first_idx;
  const int64_t* num_points;
  int C;
  CellGrid g;
  int* cells;     // counts -> starts -> ends
  uint32_t* zrange;  // [2 N] per-view range of the filed z patterns (common.cuh)
  int* cell_of;   // [P]
  float4* recA;
  float4* recB;
  // fused path only: records in packed order written by the uwp kernel, and their count
  const float4* preA;
  const float4* preB;
  const int64_t* total;
};

__global__ void __launch_bounds__(256) k_count(BinParams p) {
  const int n = blockIdx.y;
  const int64_t first = p.first_idx[n];
  const int64_t num = p.num_points[n];
  uint32_t nlo = 0u, hi = 0u;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = first + i;
    const float x = __ldg(p.points + q * 3 + 0);
    const float y = __ldg(p.points + q * 3 + 1);
    const float z = __ldg(p.points + q * 3 + 2);
    const int cell = point_cell(p.g, n, x, y, z);
    if (cell >= 0) {
      atomicAdd(p.cells + cell, 1);  // result unused -> RED
      const uint32_t zb = z_pattern(z);
      nlo = max(nlo, ~zb);
      hi = max(hi, zb);
    }
    p.cell_of[q] = cell;
  }
  zrange_accumulate(p.zrange, n, 0xffffffffu, nlo, hi);  // (the loop has re-converged: all 32 lanes)
}

__global__ void __launch_bounds__(256) k_fill(BinParams p) {
  const int n = blockIdx.y;
  const int64_t first = p.first_idx[n];
  const int64_t num = p.num_points[n];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = first + i;
    const int cell = p.cell_of[q];
    if (cell < 0) continue;
    const int pos = atomicAdd(p.cells + cell, 1);
    float4 a;
    a.x = __ldg(p.points + q * 3 + 0);
    a.y = __ldg(p.points + q * 3 + 1);
    a.z = __ldg(p.points + q * 3 + 2);
    a.w = __int_as_float((int)q);
    p.recA[rec_a(pos)] = a;
    if (p.features != nullptr || p.radius != nullptr) {
      float f[4] = {0.f, 0.f, 0.f, 0.f};
      if (p.features != nullptr) {
        for (int c = 0; c < p.C; ++c) f[c] = __ldg(p.features + q * p.C + c);
      }
      if (p.radius != nullptr) f[3] = __ldg(p.radius + q);
      p.recA[rec_b(pos)] = make_float4(f[0], f[1], f[2], f[3]);
    }
  }
}

// Fused path: the uwp kernel already produced the packed-order records, cell id included.
// Each thread moves kFillUnroll points with all loads, then all atomics, then all stores issued
// back to back.
#ifndef PGDVS_FILL_UNROLL
#define PGDVS_FILL_UNROLL 4
#endif
constexpr int kFillUnroll = PGDVS_FILL_UNROLL;
__global__ void __launch_bounds__(256) k_fill_pre(BinParams p) {
  const int64_t total = *p.total;
  const int64_t base = ((int64_t)blockIdx.x * blockDim.x) * kFillUnroll + threadIdx.x;
  int cell[kFillUnroll];
  float4 a[kFillUnroll], b[kFillUnroll];
#pragma unroll
  for (int u = 0; u < kFillUnroll; ++u) {
    const int64_t q = base + (int64_t)u * blockDim.x;
    cell[u] = -1;
    if (q < total) {
      // packed-order record of the uwp kernel: (x, y, z, cell) + (r, g, b)
      a[u] = __ldg(p.preA + q);
      const float* pb = reinterpret_cast<const float*>(p.preB) + q * 3;
      b[u] = make_float4(__ldg(pb), __ldg(pb + 1), __ldg(pb + 2), 0.0f);
      cell[u] = __float_as_int(a[u].w);
      a[u].w = __int_as_float((int)q);  // the packed index is the position itself
    }
  }
  int pos[kFillUnroll];
#pragma unroll
  for (int u = 0; u < kFillUnroll; ++u) pos[u] = (cell[u] >= 0) ? atomicAdd(p.cells + cell[u], 1) : -1;
#pragma unroll
  for (int u = 0; u < kFillUnroll; ++u) {
    if (pos[u] >= 0) {
      p.recA[rec_a(pos[u])] = a[u];
      p.recA[rec_b(pos[u])] = b[u];
    }
  }
}

// The same move in TILE-MAJOR order over the jobs of a view: CTA o takes the packed range of one
// (job, 1024-pixel source tile) — the scanned tile offsets of the uwp pass give it — and consecutive
// CTAs take the same source tile of all the jobs of a view before moving on to the next tile.  Source
// pairs of a view cover the same image region, so the cell lines a tile touches are completed while they
// are still in L2.  In packed (job-major) order every job sweeps the whole record array of its view;
// at C5 (8 source pairs, 530 MB of records per view) the lines leave L2 one-eighth full and DRAM sees
// eight partial writes per line: k_fill_pre ran at 40 % of the copy peak there, at 72 % on C2.
__global__ void __launch_bounds__(256) k_fill_pre_tiles(BinParams p, const PgdvsUwpJob* __restrict__ jobs, int n_jobs,
                                                        int tiles_per_job, const int* __restrict__ tile_off) {
  static_assert(kFillUnroll * 256 >= kUwpTilePixels, "one CTA moves one tile of the uwp pass");
  __shared__ int s_range[2];
  if (threadIdx.x == 0) {
    const int o = blockIdx.x, jb = o / tiles_per_job;
    const int v = jobs[jb].view;
    int f = jb, l = jb;
    while (f > 0 && jobs[f - 1].view == v) --f;
    while (l + 1 < n_jobs && jobs[l + 1].view == v) ++l;
    const int c = l - f + 1, local = o - f * tiles_per_job;
    const int jt = local / c, j = f + local % c;
    const int64_t t = (int64_t)j * tiles_per_job + jt;
    s_range[0] = __ldg(tile_off + t);
    s_range[1] = __ldg(tile_off + t + 1);
  }
  __syncthreads();
  const int q0 = s_range[0], q1 = s_range[1];
  int cell[kFillUnroll];
  float4 a[kFillUnroll], b[kFillUnroll];
#pragma unroll
  for (int u = 0; u < kFillUnroll; ++u) {
    const int q = q0 + u * 256 + (int)threadIdx.x;
    cell[u] = -1;
    if (q < q1) {
      a[u] = __ldg(p.preA + q);
      const float* pb = reinterpret_cast<const float*>(p.preB) + (int64_t)q * 3;
      b[u] = make_float4(__ldg(pb), __ldg(pb + 1), __ldg(pb + 2), 0.0f);
      cell[u] = __float_as_int(a[u].w);
      a[u].w = __int_as_float(q);  // the packed index is the position itself
    }
  }
  int pos[kFillUnroll];
#pragma unroll
  for (int u = 0; u < kFillUnroll; ++u) pos[u] = (cell[u] >= 0) ? atomicAdd(p.cells + cell[u], 1) : -1;
#pragma unroll
  for (int u = 0; u < kFillUnroll; ++u) {
    if (pos[u] >= 0) {
      p.recA[rec_a(pos[u])] = a[u];
      p.recA[rec_b(pos[u])] = b[u];
    }
  }
}

// ---------------------------------------------------------------------------------------
// In-place exclusive scan, single pass with decoupled look-back.
// state[t] = (flag << 32) | value; flag 1 = tile aggregate, 2 = inclusive prefix.
// Tiles take their index from an atomic ticket so a tile's predecessors are always
// already running (forward progress without relying on block scheduling order).
// ---------------------------------------------------------------------------------------
constexpr unsigned long long kFlagAgg = 1ull << 32;
constexpr unsigned long long kFlagPrefix = 2ull << 32;

__global__ void __launch_bounds__(1024) k_scan(int* data, unsigned long long* state, int* ticket) {
  __shared__ int s_tile;
  __shared__ int s_warp[32];
  __shared__ int s_prefix;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1);
  __syncthreads();
  const int tile = s_tile;
  int4* base = reinterpret_cast<int4*>(data + (int64_t)tile * kScanTile);
  int4 v = base[threadIdx.x];
  const int t_sum = v.x + v.y + v.z + v.w;
  // warp inclusive scan of per-thread sums
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = t_sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = s_warp[lane];
    int winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int o = __shfl_up_sync(0xffffffffu, winc, d);
      if (lane >= d) winc += o;
    }
    s_warp[lane] = winc - w;  // exclusive offset of each warp
    const int aggregate = __shfl_sync(0xffffffffu, winc, 31);
    // publish the aggregate, then look back for the exclusive prefix of this tile
    int prefix = 0;
    if (tile == 0) {
      if (lane == 0) atomicExch(state + tile, kFlagPrefix | (unsigned int)aggregate);
    } else {
      if (lane == 0) atomicExch(state + tile, kFlagAgg | (unsigned int)aggregate);
      int look = tile - 1;
      while (true) {
        const int idx = look - lane;
        unsigned long long st = kFlagPrefix;  // lanes past the start behave like a zero prefix
        if (idx >= 0) {
          do {
            st = *reinterpret_cast<volatile unsigned long long*>(state + idx);
          } while ((st >> 32) == 0);
        }
        const unsigned has_prefix = __ballot_sync(0xffffffffu, (st >> 32) == 2);
        int val = (int)(unsigned int)(st & 0xffffffffull);
        if (has_prefix) {
          const int firstp = __ffs(has_prefix) - 1;  // nearest tile that has a full prefix
          if (lane > firstp) val = 0;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
        prefix += val;
        if (has_prefix) break;
        look -= 32;
      }
      if (lane == 0) atomicExch(state + tile, kFlagPrefix | (unsigned int)(prefix + aggregate));
    }
    if (lane == 0) s_prefix = prefix;
  }
  __syncthreads();
  const int excl = s_prefix + s_warp[warp] + (inc - t_sum);
  int4 o;
  o.x = excl;
  o.y = excl + v.x;
  o.z = o.y + v.y;
  o.w = o.z + v.z;
  base[threadIdx.x] = o;
}

// In-place exclusive scan of n_tiles * kScanTile ints (state / ticket must be zeroed).
int scan_exclusive_inplace(int* data, int64_t n_tiles, unsigned long long* state, int* ticket,
                           cudaStream_t stream) {
  k_scan<<<(unsigned)n_tiles, 1024, 0, stream>>>(data, state, ticket);
  return check_launch();
}

// developer switch, read once: PGDVS_FILL_TILES=0 never / 1 always the tile-major fill; unset: automatic
static int debug_fill_tiles() {
  static const int v = [] {
    const char* e = getenv("PGDVS_FILL_TILES");
    return (e && *e) ? (atoi(e) != 0 ? 1 : 0) : -1;
  }();
  return v;
}

// Second half of the fused path (called by pgdvs_uwp_bin in uwp.cu): scan the cell counters
// the uwp kernel accumulated, then scatter its packed-order records into cell order.
int bin_scan_fill_fused(char* ws, const BinLayout& L, const FusedTail& T, int64_t capacity,
                        const int64_t* total_dev, const PgdvsUwpJob* jobs, int n_jobs, int n_views,
                        int tiles_per_job, const int* tile_off, cudaStream_t stream) {
  BinParams p = {};
  p.cells = reinterpret_cast<int*>(ws + L.off_cells);
  p.cell_of = reinterpret_cast<int*>(ws + L.off_cell_of);
  p.recA = reinterpret_cast<float4*>(ws + L.off_recA);
  p.recB = reinterpret_cast<float4*>(ws + L.off_recB);
  p.preA = reinterpret_cast<const float4*>(ws + T.off_preA);
  p.preB = reinterpret_cast<const float4*>(ws + T.off_preB);
  p.total = total_dev;

  k_scan<<<(unsigned)L.n_tiles, 1024, 0, stream>>>(
      p.cells, reinterpret_cast<unsigned long long*>(ws + L.off_state),
      reinterpret_cast<int*>(ws + L.off_ticket));
  if (int rc = check_launch()) return rc;
  if (capacity > 0) {
    // tile-major when the records of one view are too many to stay in L2 while its jobs take turns
    const bool tile_major = debug_fill_tiles() != 0 && jobs != nullptr && n_views > 0 &&
                            (debug_fill_tiles() > 0 || (capacity / n_views) * 32 > ((int64_t)64 << 20));
    if (tile_major) {
      k_fill_pre_tiles<<<(unsigned)((int64_t)n_jobs * tiles_per_job), 256, 0, stream>>>(p, jobs, n_jobs, tiles_per_job, tile_off);
    } else {
      const int64_t g = (capacity + 256 * kFillUnroll - 1) / (256 * kFillUnroll);
      k_fill_pre<<<(unsigned)g, 256, 0, stream>>>(p);
    }
    if (int rc = check_launch()) return rc;
  }
  return 0;
}

}  // namespace pgdvs

using namespace pgdvs;

extern "C" int pgdvs_bin_workspace_bytes(int N, int H, int W, int64_t P, float radius_max,
                                         size_t* bytes) {
  if (bytes == nullptr || N < 0 || H <= 0 || W <= 0 || P < 0 || !(radius_max >= 0.0f))
    return PGDVS_E_BADARG;
  if (P >= kMaxRecords) return PGDVS_E_BADARG;
  BinLayout L = make_bin_layout(N, H, W, P, radius_max);
  if (L.cells + kScanTile >= (int64_t)INT32_MAX) return PGDVS_E_BADARG;
  *bytes = L.total;
  return PGDVS_OK;
}

extern "C" int pgdvs_bin_points(const float* points, const float* features, int C,
                                const int64_t* first_idx, const int64_t* num_points, int N,
                                int64_t P, const float* radius, float radius_max, int H, int W,
                                void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (N < 0 || H <= 0 || W <= 0 || P < 0 || !(radius_max >= 0.0f) || workspace == nullptr)
    return PGDVS_E_BADARG;
  if (P >= kMaxRecords) return PGDVS_E_BADARG;  // record float4 indices (2 per record) are int32
  if (features != nullptr && (C < 1 || C > PGDVS_MAX_FUSED_CHANNELS)) return PGDVS_E_CHANNELS;
  if (features != nullptr && radius != nullptr && C > 3) return PGDVS_E_CHANNELS;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return PGDVS_E_ALIGN;
  BinLayout L = make_bin_layout(N, H, W, P, radius_max);
  if (L.cells + kScanTile >= (int64_t)INT32_MAX) return PGDVS_E_BADARG;
  if (workspace_bytes < L.total) return PGDVS_E_WORKSPACE;
  if (N == 0) return PGDVS_OK;
  if (P > 0 && (points == nullptr || first_idx == nullptr || num_points == nullptr))
    return PGDVS_E_BADARG;

  char* ws = static_cast<char*>(workspace);
  BinParams p = {};
  p.points = points;
  p.features = features;
  p.radius = radius;
  p.first_idx = first_idx;
  p.num_points = num_points;
  p.C = features ? C : 0;
  p.g = make_cell_grid(H, W, L.halo);
  p.cells = reinterpret_cast<int*>(ws + L.off_cells);
  p.zrange = reinterpret_cast<uint32_t*>(ws + L.off_zrange);
  p.cell_of = reinterpret_cast<int*>(ws + L.off_cell_of);
  p.recA = reinterpret_cast<float4*>(ws + L.off_recA);
  p.recB = reinterpret_cast<float4*>(ws + L.off_recB);

  // front pad, counters, scan state and ticket are contiguous at the front of the workspace
  cudaError_t e = cudaMemsetAsync(ws, 0, L.off_zero_end, stream);
  if (e != cudaSuccess) return (int)e;
  int gx = 1;
  if (P > 0) {
    // enough blocks to fill the machine even for one cloud; grid-stride inside
    const int64_t per_cloud = (P + N - 1) / N;
    gx = (int)((per_cloud + 255) / 256);
    if (gx < 1) gx = 1;
    if (gx > 148 * 16) gx = 148 * 16;
    k_count<<<dim3(gx, N), 256, 0, stream>>>(p);
    if (int rc = check_launch()) return rc;
  }
  k_scan<<<(unsigned)L.n_tiles, 1024, 0, stream>>>(
      p.cells, reinterpret_cast<unsigned long long*>(ws + L.off_state),
      reinterpret_cast<int*>(ws + L.off_ticket));
  if (int rc = check_launch()) return rc;
  if (P > 0) {
    k_fill<<<dim3(gx, N), 256, 0, stream>>>(p);
    if (int rc = check_launch()) return rc;
  }
  return PGDVS_OK;
}
