// Fused unproject -> flow-warp -> time-lerp -> project kernel.
//
// One launch turns a batch of (target view, source-frame pair) jobs into a packed NDC point
// cloud ready for the rasterizer.  It replaces ~25 separate torch kernels and six
// boolean-index compactions (each a host sync) of the reference:
//   get_batched_rays                 pgdvs_renderer_base.py:17-57
//   compute_dyn_pcl (geometry part)  pgdvs_renderer_dyn.py:304-388
//   w2c / camera / transform         pgdvs_renderer_dyn.py:676-687 + PointsRasterizer.transform
//
// Each thread owns 4 consecutive source pixels (float4-vectorised, fully coalesced reads of
// depth / mask / occlusion / flow / rgb).  Frame-2 colour (bilinear) and depth (nearest) are
// gathered from a packed (r,g,b,depth) float4 plane when the caller provides one: the nearest
// pixel is always one of the four bilinear taps, so 4 x 128-bit loads replace 13 scalar ones.
// Surviving points are written in the reference's order (job-major, row-major pixels): a
// block-level prefix sum orders points inside a 1024-pixel tile and a single-pass chained
// scan (decoupled look-back, ticketed tiles) orders the tiles, so the packed indices — and
// therefore the rasterizer's idx output — are identical to the reference's boolean-mask
// compaction.  In the fused mode (pgdvs_uwp_bin) the kernel also files every point under its
// raster cell (one atomicAdd), which removes the separate counting pass over the cloud.
#include "common.cuh"

namespace pgdvs {

int bin_scan_fill_fused(char* ws, const BinLayout& L, const FusedTail& T, int64_t capacity,
                        const int64_t* total_dev, cudaStream_t stream);

constexpr int kUwpThreads = 256;
constexpr int kUwpPix = 4;                               // pixels per thread
constexpr int kUwpTile = kUwpThreads * kUwpPix;          // pixels per tile
constexpr unsigned long long kUFlagAgg = 1ull << 32;
constexpr unsigned long long kUFlagPrefix = 2ull << 32;

struct UwpParams {
  const PgdvsUwpJob* jobs;
  const PgdvsCamera* cams;
  int n_jobs, H, W;
  int tiles_per_job;
  int64_t n_tiles;
  float* xyz_ndc;    // [cap,3] or null
  float* rgb;        // [cap,3] or null
  float* xyz_world;  // [cap,3] or null
  int32_t* src_pix;  // [cap] or null
  unsigned long long* state;  // [n_tiles]
  int* ticket;
  int64_t* job_start;         // [n_jobs + 1]
  // fused binning (all null/0 in the plain mode)
  CellGrid g;
  int* cell_count;
  int* cell_of;
  float4* preA;
  float4* preB;
};

// torch grid_sample(align_corners=False) source index of pixel coordinate c on an axis of
// `size` pixels, as compute_dyn_pcl builds it (pgdvs_renderer_dyn.py:341):
//   g = 2*c/size - 1 ;  ix = ((g + 1) * size - 1) / 2          (every op rounded to fp32;
// identical to ATen's CPU `(g + 1) * (size/2) - 0.5`, checked in tests)
__device__ __forceinline__ float grid_unnormalize(float c, float size) {
  const float g = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, c), size), 1.0f);
  return __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.0f), size), 1.0f), 0.5f);
}

__device__ __forceinline__ void load4(const float* p, int64_t i, bool vec, int64_t limit, float o[4]) {
  if (vec) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p + i));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = (i + k < limit) ? __ldg(p + i + k) : 0.0f;
  }
}

template <bool FUSED>
__global__ void __launch_bounds__(kUwpThreads, 2) k_uwp(const __grid_constant__ UwpParams p) {
  __shared__ int s_tile;
  __shared__ int s_warp[kUwpThreads / 32];
  __shared__ long long s_prefix;
  if (threadIdx.x == 0) s_tile = atomicAdd(p.ticket, 1);
  __syncthreads();
  const int tile = s_tile;
  const int job_i = tile / p.tiles_per_job;
  const int jt = tile - job_i * p.tiles_per_job;
  const PgdvsUwpJob& J = p.jobs[job_i];
  const int64_t HW = (int64_t)p.H * p.W;
  const int64_t pix0 = (int64_t)jt * kUwpTile + (int64_t)threadIdx.x * kUwpPix;
  const bool in_range = pix0 < HW;
  const bool vec = ((HW & 3) == 0) && in_range;  // host guarantees 16-byte aligned planes
  // (u, v) of the first pixel; the other three follow by incrementing (one div per thread)
  const int v0 = (int)(pix0 / p.W), u0 = (int)(pix0 - (int64_t)v0 * p.W);

  // ------------------------------------------------------------ validity (cheap loads only)
  float m[4] = {0, 0, 0, 0}, fl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  unsigned valid = 0;
  if (in_range) {
    load4(J.mask1, pix0, vec, HW, m);
    load4(J.flow12, pix0 * 2, vec, HW * 2, fl);
    load4(J.flow12, pix0 * 2 + 4, vec, HW * 2, fl + 4);
    float oc[4] = {0, 0, 0, 0};
    if (J.occ12 != nullptr) load4(J.occ12, pix0, vec, HW, oc);
    int uu = u0, vv = v0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t pix = pix0 + k;
      if (pix < HW) {
        bool ok = (m[k] != 0.0f);                            // dyn_mask.bool()
        if (J.occ12 != nullptr) ok = ok && !(oc[k] > 0.0f);  // ~(occ > 0) & mask
        const float u2 = __fadd_rn((float)uu, fl[2 * k]), v2 = __fadd_rn((float)vv, fl[2 * k + 1]);
        ok = ok && (u2 >= 0.0f) && (u2 <= (float)(p.W - 1)) && (v2 >= 0.0f) && (v2 <= (float)(p.H - 1));
        if (ok && J.keep != nullptr) ok = (J.keep[pix] != 0);
        if (ok) valid |= 1u << k;
      }
      if (++uu == p.W) { uu = 0; ++vv; }
    }
  }
  const int cnt = __popc(valid);

  // ------------------------------------------------------------ order inside the tile
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  int aggregate = 0;
  if (warp == 0) {
    const int w = (lane < kUwpThreads / 32) ? s_warp[lane] : 0;
    int winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, winc, d);
      if (lane >= d) winc += o;
    }
    if (lane < kUwpThreads / 32) s_warp[lane] = winc - w;  // read after the next barrier
    aggregate = __shfl_sync(0xffffffffu, winc, 31);
    // publish this tile's count NOW, so that successors never wait on our geometry
    if (lane == 0)
      atomicExch(p.state + tile, (tile == 0 ? kUFlagPrefix : kUFlagAgg) | (unsigned int)aggregate);
  }

  // ------------------------------------------------------------ geometry (branch-free over the
  // 4 pixels so that all 16 frame-2 taps are in flight together; stores are predicated later)
  float d1[4], c1[12];
  float ox[4], oy[4], oz[4], cr[4], cg[4], cb[4], wx[4], wy[4], wz[4];
  const bool any_valid = valid != 0;
  const PgdvsCamera cam = p.cams[J.view];
  const bool lerp = (J.same_time == 0);
  if (any_valid) {
    load4(J.depth1, pix0, vec, HW, d1);
    if (!lerp) {
      load4(J.rgb1, pix0 * 3, vec, HW * 3, c1);
      load4(J.rgb1, pix0 * 3 + 4, vec, HW * 3, c1 + 4);
      load4(J.rgb1, pix0 * 3 + 8, vec, HW * 3, c1 + 8);
    }
    const float4* __restrict__ rgbd2 = reinterpret_cast<const float4*>(J.rgbd2);
    float u2a[4], v2a[4], wgt[4][4], dep2[4];
    float4 tap[4][4];
    int near_tap[4];
    {
      int uu = u0, vv = v0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float u = (float)uu, v = (float)vv;
        if (++uu == p.W) { uu = 0; ++vv; }
        // rays_d = (c2w[:3,:3] @ K^-1) @ [u, v, 1]
        wx[k] = J.o1[0] + (J.M1[0] * u + J.M1[1] * v + J.M1[2]) * d1[k];
        wy[k] = J.o1[1] + (J.M1[3] * u + J.M1[4] * v + J.M1[5]) * d1[k];
        wz[k] = J.o1[2] + (J.M1[6] * u + J.M1[7] * v + J.M1[8]) * d1[k];
        cr[k] = cg[k] = cb[k] = 0.0f;
        dep2[k] = 0.0f;
        if (lerp) {
          const bool ok = (valid >> k) & 1u;
          const float u2 = __fadd_rn(u, fl[2 * k]), v2 = __fadd_rn(v, fl[2 * k + 1]);
          u2a[k] = u2;
          v2a[k] = v2;
          const float ix = grid_unnormalize(u2, (float)p.W);
          const float iy = grid_unnormalize(v2, (float)p.H);
          // depth_2: grid_sample(mode="nearest"): nearbyint (half to even), zeros padding
          const float nx = nearbyintf(ix), ny = nearbyintf(iy);
          // rgb: grid_sample(rgb_2, mode="bilinear"), zeros padding (colour comes from frame 2)
          const float x0f = floorf(ix), y0f = floorf(iy);
          const float tw = __fsub_rn(ix, x0f), te = __fsub_rn(1.0f, tw);
          const float tn = __fsub_rn(iy, y0f), ts = __fsub_rn(1.0f, tn);
          const int x0 = (int)x0f, y0 = (int)y0f;
          wgt[k][0] = __fmul_rn(ts, te);
          wgt[k][1] = __fmul_rn(ts, tw);
          wgt[k][2] = __fmul_rn(tn, te);
          wgt[k][3] = __fmul_rn(tn, tw);
          // the nearest pixel is one of the 4 bilinear taps
          near_tap[k] = ((ny != y0f) ? 2 : 0) + ((nx != x0f) ? 1 : 0);
          if (rgbd2 != nullptr) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int xs = x0 + (t & 1), ys = y0 + (t >> 1);
              const bool inb = ok && xs >= 0 && xs < p.W && ys >= 0 && ys < p.H;
              tap[k][t] = make_float4(0.f, 0.f, 0.f, 0.f);  // zeros padding
              if (inb) tap[k][t] = __ldg(rgbd2 + (int64_t)ys * p.W + xs);
            }
          } else {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int xs = x0 + (t & 1), ys = y0 + (t >> 1);
              const bool inb = ok && xs >= 0 && xs < p.W && ys >= 0 && ys < p.H;
              tap[k][t] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (inb) {
                const int64_t o = (int64_t)ys * p.W + xs;
                tap[k][t] = make_float4(__ldg(J.rgb2 + o * 3), __ldg(J.rgb2 + o * 3 + 1),
                                        __ldg(J.rgb2 + o * 3 + 2), __ldg(J.depth2 + o));
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!lerp) {
        cr[k] = c1[3 * k];
        cg[k] = c1[3 * k + 1];
        cb[k] = c1[3 * k + 2];
      } else {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          cr[k] = __fadd_rn(cr[k], __fmul_rn(tap[k][t].x, wgt[k][t]));
          cg[k] = __fadd_rn(cg[k], __fmul_rn(tap[k][t].y, wgt[k][t]));
          cb[k] = __fadd_rn(cb[k], __fmul_rn(tap[k][t].z, wgt[k][t]));
          if (t == near_tap[k]) dep2[k] = tap[k][t].w;
        }
        // pcl_2 = o_2 + (R_2 @ (K_2^-1 @ [u2, v2, 1])) * depth_2
        const float u2 = u2a[k], v2 = v2a[k];
        const float kx = J.K2inv[0] * u2 + J.K2inv[1] * v2 + J.K2inv[2];
        const float ky = J.K2inv[3] * u2 + J.K2inv[4] * v2 + J.K2inv[5];
        const float kz = J.K2inv[6] * u2 + J.K2inv[7] * v2 + J.K2inv[8];
        const float qx = J.o2[0] + (J.R2[0] * kx + J.R2[1] * ky + J.R2[2] * kz) * dep2[k];
        const float qy = J.o2[1] + (J.R2[3] * kx + J.R2[4] * ky + J.R2[5] * kz) * dep2[k];
        const float qz = J.o2[2] + (J.R2[6] * kx + J.R2[7] * ky + J.R2[8] * kz) * dep2[k];
        wx[k] = J.w1 * wx[k] + J.w2 * qx;
        wy[k] = J.w1 * wy[k] + J.w2 * qy;
        wz[k] = J.w1 * wz[k] + J.w2 * qz;
      }
      const float3 ndc = world_to_ndc(cam, wx[k], wy[k], wz[k]);
      ox[k] = ndc.x;
      oy[k] = ndc.y;
      oz[k] = ndc.z;
    }
  }

  // ------------------------------------------------------------ order between tiles
  if (warp == 0) {
    long long prefix = 0;
    if (tile != 0) {
      int look = tile - 1;
      while (true) {
        const int idx = look - lane;
        unsigned long long st = kUFlagPrefix;  // lanes past the start act as a zero prefix
        if (idx >= 0) {
          do {
            st = *reinterpret_cast<volatile unsigned long long*>(p.state + idx);
          } while ((st >> 32) == 0);
        }
        const unsigned has_prefix = __ballot_sync(0xffffffffu, (st >> 32) == 2);
        int val = (int)(unsigned int)(st & 0xffffffffull);
        if (has_prefix) {
          const int firstp = __ffs(has_prefix) - 1;
          if (lane > firstp) val = 0;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
        prefix += val;
        if (has_prefix) break;
        look -= 32;
      }
      if (lane == 0)
        atomicExch(p.state + tile, kUFlagPrefix | (unsigned int)(prefix + aggregate));
    }
    if (lane == 0) {
      s_prefix = prefix;
      if (jt == 0) p.job_start[job_i] = prefix;
      if ((int64_t)tile == p.n_tiles - 1) p.job_start[p.n_jobs] = prefix + aggregate;
    }
  }
  __syncthreads();
  if (!any_valid) return;
  int64_t out = s_prefix + s_warp[warp] + (inc - cnt);

  // ------------------------------------------------------------ stores, in packed order
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (!((valid >> k) & 1u)) continue;
    if (FUSED) {
      const int cell = point_cell(p.g, J.view, ox[k], oy[k], oz[k]);
      if (cell >= 0) atomicAdd(p.cell_count + cell, 1);  // result unused -> RED
      p.cell_of[out] = cell;
      p.preA[out] = make_float4(ox[k], oy[k], oz[k], __int_as_float((int)out));
      p.preB[out] = make_float4(cr[k], cg[k], cb[k], 0.0f);
    }
    if (p.xyz_ndc) {
      p.xyz_ndc[out * 3 + 0] = ox[k];
      p.xyz_ndc[out * 3 + 1] = oy[k];
      p.xyz_ndc[out * 3 + 2] = oz[k];
    }
    if (p.rgb) {
      p.rgb[out * 3 + 0] = cr[k];
      p.rgb[out * 3 + 1] = cg[k];
      p.rgb[out * 3 + 2] = cb[k];
    }
    if (p.xyz_world) {
      p.xyz_world[out * 3 + 0] = wx[k];
      p.xyz_world[out * 3 + 1] = wy[k];
      p.xyz_world[out * 3 + 2] = wz[k];
    }
    if (p.src_pix) p.src_pix[out] = (int32_t)(pix0 + k);
    ++out;
  }
}

// per-view first index / count from the per-job starts (jobs are sorted by view)
__global__ void k_uwp_finalize(const PgdvsUwpJob* jobs, int n_jobs, int n_views,
                               const int64_t* job_start, int64_t* first_idx, int64_t* num_points,
                               int64_t* total) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v == 0 && total) *total = job_start[n_jobs];
  if (v >= n_views) return;
  int a = n_jobs, b = n_jobs;
  for (int j = n_jobs - 1; j >= 0; --j) {
    const int jv = jobs[j].view;
    if (jv >= v) a = j;
    if (jv >= v + 1) b = j;
  }
  if (first_idx) first_idx[v] = job_start[a];
  if (num_points) num_points[v] = job_start[b] - job_start[a];
}

// (r,g,b) [HW,3] + depth [HW] -> (r,g,b,depth) float4 [HW] for a batch of frames
__global__ void __launch_bounds__(256) k_pack_rgbd(const PgdvsFramePack* frames, int64_t HW) {
  const PgdvsFramePack f = frames[blockIdx.y];
  float4* __restrict__ out = reinterpret_cast<float4*>(f.rgbd);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW;
       i += (int64_t)gridDim.x * blockDim.x) {
    out[i] = make_float4(__ldg(f.rgb + i * 3), __ldg(f.rgb + i * 3 + 1), __ldg(f.rgb + i * 3 + 2),
                         __ldg(f.depth + i));
  }
}

struct UwpLayout {
  int tiles_per_job;
  int64_t n_tiles;
  size_t off_state, off_ticket, off_job_start, total;
};

static inline UwpLayout make_uwp_layout(int n_jobs, int H, int W) {
  UwpLayout L;
  const int64_t HW = (int64_t)H * W;
  L.tiles_per_job = (int)((HW + kUwpTile - 1) / kUwpTile);
  L.n_tiles = (int64_t)L.tiles_per_job * n_jobs;
  size_t o = 0;
  L.off_state = o;
  o = align256(o + sizeof(unsigned long long) * (size_t)(L.n_tiles > 0 ? L.n_tiles : 1));
  L.off_ticket = o;
  o = align256(o + 256);
  L.off_job_start = o;
  o = align256(o + sizeof(int64_t) * (size_t)(n_jobs + 1));
  L.total = o;
  return L;
}

static int run_uwp(const PgdvsUwpJob* jobs, int n_jobs, const PgdvsCamera* cameras, int n_views, int H,
                   int W, float* xyz_ndc, float* rgb, float* xyz_world, int32_t* src_pix,
                   int64_t* first_idx, int64_t* num_points, int64_t* total_points, char* ws,
                   const UwpLayout& L, const CellGrid* grid, int* cell_count, int* cell_of, float4* preA,
                   float4* preB, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(ws, 0, L.total, stream);
  if (e != cudaSuccess) return (int)e;
  if (n_jobs > 0) {
    UwpParams p = {};
    p.jobs = jobs;
    p.cams = cameras;
    p.n_jobs = n_jobs;
    p.H = H;
    p.W = W;
    p.tiles_per_job = L.tiles_per_job;
    p.n_tiles = L.n_tiles;
    p.xyz_ndc = xyz_ndc;
    p.rgb = rgb;
    p.xyz_world = xyz_world;
    p.src_pix = src_pix;
    p.state = reinterpret_cast<unsigned long long*>(ws + L.off_state);
    p.ticket = reinterpret_cast<int*>(ws + L.off_ticket);
    p.job_start = reinterpret_cast<int64_t*>(ws + L.off_job_start);
    if (grid != nullptr) {
      p.g = *grid;
      p.cell_count = cell_count;
      p.cell_of = cell_of;
      p.preA = preA;
      p.preB = preB;
      k_uwp<true><<<(unsigned)L.n_tiles, kUwpThreads, 0, stream>>>(p);
    } else {
      k_uwp<false><<<(unsigned)L.n_tiles, kUwpThreads, 0, stream>>>(p);
    }
    if (int rc = check_launch()) return rc;
  }
  // always run: also publishes total_points (0 when there are no jobs: job_start is zeroed)
  k_uwp_finalize<<<(n_views + 128) / 128, 128, 0, stream>>>(
      jobs, n_jobs, n_views, reinterpret_cast<const int64_t*>(ws + L.off_job_start), first_idx,
      num_points, total_points);
  return check_launch();
}

}  // namespace pgdvs

using namespace pgdvs;

extern "C" int pgdvs_uwp_workspace_bytes(int n_jobs, int H, int W, size_t* bytes) {
  if (!bytes || n_jobs < 0 || H <= 0 || W <= 0) return PGDVS_E_BADARG;
  *bytes = make_uwp_layout(n_jobs, H, W).total;
  return PGDVS_OK;
}

extern "C" int pgdvs_unproject_warp_project(const PgdvsUwpJob* jobs, int n_jobs,
                                            const PgdvsCamera* cameras, int n_views, int H, int W,
                                            float* xyz_ndc, float* rgb, float* xyz_world,
                                            int32_t* src_pix, int64_t* first_idx,
                                            int64_t* num_points, int64_t* total_points,
                                            void* workspace, size_t workspace_bytes, void* stream_) {
  if (n_jobs < 0 || n_views < 0 || H <= 0 || W <= 0 || !workspace) return PGDVS_E_BADARG;
  if (n_views > 0 && (!first_idx || !num_points)) return PGDVS_E_BADARG;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return PGDVS_E_ALIGN;
  UwpLayout L = make_uwp_layout(n_jobs, H, W);
  if (workspace_bytes < L.total) return PGDVS_E_WORKSPACE;
  if ((int64_t)n_jobs * H * W >= (int64_t)INT32_MAX) return PGDVS_E_BADARG;
  if (n_jobs > 0 && (!jobs || !cameras || !xyz_ndc || !rgb)) return PGDVS_E_BADARG;
  return run_uwp(jobs, n_jobs, cameras, n_views, H, W, xyz_ndc, rgb, xyz_world, src_pix, first_idx,
                 num_points, total_points, static_cast<char*>(workspace), L, nullptr, nullptr, nullptr,
                 nullptr, nullptr, (cudaStream_t)stream_);
}

extern "C" int pgdvs_uwp_bin_workspace_bytes(int n_jobs, int n_views, int H, int W, float radius,
                                             size_t* bytes) {
  if (!bytes || n_jobs < 0 || n_views < 0 || H <= 0 || W <= 0 || !(radius >= 0.0f))
    return PGDVS_E_BADARG;
  const int64_t cap = (int64_t)n_jobs * H * W;
  if (cap >= (int64_t)INT32_MAX) return PGDVS_E_BADARG;
  BinLayout B = make_bin_layout(n_views, H, W, cap, radius);
  if (B.cells + kScanTile >= (int64_t)INT32_MAX) return PGDVS_E_BADARG;
  FusedTail T = make_fused_tail(B, cap);
  *bytes = T.total + align256(make_uwp_layout(n_jobs, H, W).total);
  return PGDVS_OK;
}

extern "C" int pgdvs_uwp_bin(const PgdvsUwpJob* jobs, int n_jobs, const PgdvsCamera* cameras,
                             int n_views, int H, int W, float radius, float* xyz_ndc, float* rgb,
                             int64_t* first_idx, int64_t* num_points, int64_t* total_points,
                             void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_jobs < 0 || n_views < 0 || H <= 0 || W <= 0 || !workspace || !(radius >= 0.0f) || !total_points)
    return PGDVS_E_BADARG;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return PGDVS_E_ALIGN;
  const int64_t cap = (int64_t)n_jobs * H * W;
  if (cap >= (int64_t)INT32_MAX) return PGDVS_E_BADARG;
  BinLayout B = make_bin_layout(n_views, H, W, cap, radius);
  if (B.cells + kScanTile >= (int64_t)INT32_MAX) return PGDVS_E_BADARG;
  FusedTail T = make_fused_tail(B, cap);
  UwpLayout U = make_uwp_layout(n_jobs, H, W);
  if (workspace_bytes < T.total + align256(U.total)) return PGDVS_E_WORKSPACE;
  if (n_jobs > 0 && (!jobs || !cameras)) return PGDVS_E_BADARG;
  if (n_views == 0) return PGDVS_OK;
  char* ws = static_cast<char*>(workspace);
  // cell counters, scan state and ticket sit contiguously at the front of the bin layout
  cudaError_t e = cudaMemsetAsync(ws, 0, B.off_zero_end, stream);
  if (e != cudaSuccess) return (int)e;
  const CellGrid g = make_cell_grid(H, W, B.halo);
  int rc = run_uwp(jobs, n_jobs, cameras, n_views, H, W, xyz_ndc, rgb, nullptr, nullptr, first_idx,
                   num_points, total_points, ws + T.total, U, &g,
                   reinterpret_cast<int*>(ws + B.off_cells), reinterpret_cast<int*>(ws + B.off_cell_of),
                   reinterpret_cast<float4*>(ws + T.off_preA), reinterpret_cast<float4*>(ws + T.off_preB),
                   stream);
  if (rc) return rc;
  return bin_scan_fill_fused(ws, B, T, cap, total_points, stream);
}

extern "C" int pgdvs_pack_rgbd(const PgdvsFramePack* frames_dev, int n_frames, int H, int W,
                               void* stream) {
  if (n_frames < 0 || H <= 0 || W <= 0) return PGDVS_E_BADARG;
  if (n_frames == 0) return PGDVS_OK;
  if (!frames_dev) return PGDVS_E_BADARG;
  const int64_t HW = (int64_t)H * W;
  int gx = (int)((HW + 255) / 256);
  if (gx > 148 * 4) gx = 148 * 4;
  dim3 grid(gx, n_frames);
  k_pack_rgbd<<<grid, 256, 0, (cudaStream_t)stream>>>(frames_dev, HW);
  return check_launch();
}
