// Fused unproject -> flow-warp -> time-lerp -> project kernel.
//
// One launch turns a batch of (target view, source-frame pair) jobs into a packed NDC point
// cloud ready for binning.  It replaces ~25 separate torch kernels and six boolean-index
// compactions (each a host sync) of the reference:
//   get_batched_rays                 pgdvs_renderer_base.py:17-57
//   compute_dyn_pcl (geometry part)  pgdvs_renderer_dyn.py:304-388
//   w2c / camera / transform         pgdvs_renderer_dyn.py:676-687 + PointsRasterizer.transform
//
// Each thread owns 4 consecutive source pixels (float4-vectorised, fully coalesced reads of
// depth / mask / occlusion / flow / rgb).  Surviving points are written in the reference's
// order (job-major, row-major pixels): a block-level prefix sum orders points inside a
// 1024-pixel tile and a single-pass chained scan (decoupled look-back, ticketed tiles)
// orders the tiles, so the packed indices — and therefore the rasterizer's idx output —
// are identical to the reference's boolean-mask compaction.
#include "common.cuh"

namespace pgdvs {

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
  float* xyz_ndc;
  float* rgb;
  float* xyz_world;
  int32_t* src_pix;
  unsigned long long* state;  // [n_tiles]
  int* ticket;
  int64_t* job_start;         // [n_jobs + 1]
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

__global__ void __launch_bounds__(kUwpThreads) k_uwp(const __grid_constant__ UwpParams p) {
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

  // ------------------------------------------------------------ validity (cheap loads only)
  float m[4] = {0, 0, 0, 0}, fl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  unsigned valid = 0;
  if (in_range) {
    load4(J.mask1, pix0, vec, HW, m);
    load4(J.flow12, pix0 * 2, vec, HW * 2, fl);
    load4(J.flow12, pix0 * 2 + 4, vec, HW * 2, fl + 4);
    float oc[4] = {0, 0, 0, 0};
    if (J.occ12 != nullptr) load4(J.occ12, pix0, vec, HW, oc);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t pix = pix0 + k;
      if (pix >= HW) break;
      bool ok = (m[k] != 0.0f);                       // dyn_mask.bool()
      if (J.occ12 != nullptr) ok = ok && !(oc[k] > 0.0f);  // ~(occ > 0) & mask
      const float u = (float)(int)(pix % p.W), v = (float)(int)(pix / p.W);
      const float u2 = __fadd_rn(u, fl[2 * k]), v2 = __fadd_rn(v, fl[2 * k + 1]);
      ok = ok && (u2 >= 0.0f) && (u2 <= (float)(p.W - 1)) && (v2 >= 0.0f) && (v2 <= (float)(p.H - 1));
      if (ok && J.keep != nullptr) ok = (J.keep[pix] != 0);
      if (ok) valid |= 1u << k;
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
  if (warp == 0) {
    const int w = (lane < kUwpThreads / 32) ? s_warp[lane] : 0;
    int winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, winc, d);
      if (lane >= d) winc += o;
    }
    if (lane < kUwpThreads / 32) s_warp[lane] = winc - w;
    const int aggregate = __shfl_sync(0xffffffffu, winc, 31);
    // ---------------------------------------------------------- order between tiles
    long long prefix = 0;
    if (tile == 0) {
      if (lane == 0) atomicExch(p.state + tile, kUFlagPrefix | (unsigned int)aggregate);
    } else {
      if (lane == 0) atomicExch(p.state + tile, kUFlagAgg | (unsigned int)aggregate);
      int look = tile - 1;
      while (true) {
        const int idx = look - lane;
        unsigned long long st = kUFlagPrefix;
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
  if (valid == 0) return;
  int64_t out = s_prefix + s_warp[warp] + (inc - cnt);

  // ------------------------------------------------------------ geometry for survivors
  float d1[4], c1[12];
  load4(J.depth1, pix0, vec, HW, d1);
  const PgdvsCamera cam = p.cams[J.view];
  const bool lerp = (J.same_time == 0);
  if (!lerp) {
    load4(J.rgb1, pix0 * 3, vec, HW * 3, c1);
    load4(J.rgb1, pix0 * 3 + 4, vec, HW * 3, c1 + 4);
    load4(J.rgb1, pix0 * 3 + 8, vec, HW * 3, c1 + 8);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (!(valid & (1u << k))) continue;
    const int64_t pix = pix0 + k;
    const float u = (float)(int)(pix % p.W), v = (float)(int)(pix / p.W);
    // rays_d = (c2w[:3,:3] @ K^-1) @ [u, v, 1]
    float wx = J.o1[0] + (J.M1[0] * u + J.M1[1] * v + J.M1[2]) * d1[k];
    float wy = J.o1[1] + (J.M1[3] * u + J.M1[4] * v + J.M1[5]) * d1[k];
    float wz = J.o1[2] + (J.M1[6] * u + J.M1[7] * v + J.M1[8]) * d1[k];
    float cr, cg, cb;
    if (!lerp) {
      cr = c1[3 * k]; cg = c1[3 * k + 1]; cb = c1[3 * k + 2];
    } else {
      const float u2 = __fadd_rn(u, fl[2 * k]), v2 = __fadd_rn(v, fl[2 * k + 1]);
      const float ix = grid_unnormalize(u2, (float)p.W);
      const float iy = grid_unnormalize(v2, (float)p.H);
      // depth_2: grid_sample(mode="nearest"): nearbyint (half to even), zeros padding
      const float nx = nearbyintf(ix), ny = nearbyintf(iy);
      float dep2 = 0.0f;
      if (nx >= 0.0f && nx <= (float)(p.W - 1) && ny >= 0.0f && ny <= (float)(p.H - 1))
        dep2 = __ldg(J.depth2 + (int64_t)ny * p.W + (int64_t)nx);
      // rgb: grid_sample(rgb_2, mode="bilinear"), zeros padding  (colour comes from frame 2)
      const float x0f = floorf(ix), y0f = floorf(iy);
      const float tw = __fsub_rn(ix, x0f), te = __fsub_rn(1.0f, tw);
      const float tn = __fsub_rn(iy, y0f), ts = __fsub_rn(1.0f, tn);
      const int x0 = (int)x0f, y0 = (int)y0f;
      const float wgt[4] = {__fmul_rn(ts, te), __fmul_rn(ts, tw), __fmul_rn(tn, te), __fmul_rn(tn, tw)};
      const int xs[4] = {x0, x0 + 1, x0, x0 + 1};
      const int ys[4] = {y0, y0, y0 + 1, y0 + 1};
      cr = cg = cb = 0.0f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (xs[t] >= 0 && xs[t] < p.W && ys[t] >= 0 && ys[t] < p.H) {
          const float* q = J.rgb2 + ((int64_t)ys[t] * p.W + xs[t]) * 3;
          cr = __fadd_rn(cr, __fmul_rn(__ldg(q + 0), wgt[t]));
          cg = __fadd_rn(cg, __fmul_rn(__ldg(q + 1), wgt[t]));
          cb = __fadd_rn(cb, __fmul_rn(__ldg(q + 2), wgt[t]));
        }
      }
      // pcl_2 = o_2 + (R_2 @ (K_2^-1 @ [u2, v2, 1])) * depth_2
      const float kx = J.K2inv[0] * u2 + J.K2inv[1] * v2 + J.K2inv[2];
      const float ky = J.K2inv[3] * u2 + J.K2inv[4] * v2 + J.K2inv[5];
      const float kz = J.K2inv[6] * u2 + J.K2inv[7] * v2 + J.K2inv[8];
      const float qx = J.o2[0] + (J.R2[0] * kx + J.R2[1] * ky + J.R2[2] * kz) * dep2;
      const float qy = J.o2[1] + (J.R2[3] * kx + J.R2[4] * ky + J.R2[5] * kz) * dep2;
      const float qz = J.o2[2] + (J.R2[6] * kx + J.R2[7] * ky + J.R2[8] * kz) * dep2;
      wx = J.w1 * wx + J.w2 * qx;
      wy = J.w1 * wy + J.w2 * qy;
      wz = J.w1 * wz + J.w2 * qz;
    }
    const float3 ndc = world_to_ndc(cam, wx, wy, wz);
    p.xyz_ndc[out * 3 + 0] = ndc.x;
    p.xyz_ndc[out * 3 + 1] = ndc.y;
    p.xyz_ndc[out * 3 + 2] = ndc.z;
    p.rgb[out * 3 + 0] = cr;
    p.rgb[out * 3 + 1] = cg;
    p.rgb[out * 3 + 2] = cb;
    if (p.xyz_world) {
      p.xyz_world[out * 3 + 0] = wx;
      p.xyz_world[out * 3 + 1] = wy;
      p.xyz_world[out * 3 + 2] = wz;
    }
    if (p.src_pix) p.src_pix[out] = (int32_t)pix;
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
  first_idx[v] = job_start[a];
  num_points[v] = job_start[b] - job_start[a];
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
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_jobs < 0 || n_views < 0 || H <= 0 || W <= 0 || !workspace) return PGDVS_E_BADARG;
  if (n_views > 0 && (!first_idx || !num_points)) return PGDVS_E_BADARG;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return PGDVS_E_ALIGN;
  UwpLayout L = make_uwp_layout(n_jobs, H, W);
  if (workspace_bytes < L.total) return PGDVS_E_WORKSPACE;
  if ((int64_t)n_jobs * H * W >= (int64_t)INT32_MAX) return PGDVS_E_BADARG;
  char* ws = static_cast<char*>(workspace);
  cudaError_t e = cudaMemsetAsync(ws, 0, L.total, stream);
  if (e != cudaSuccess) return (int)e;
  if (n_jobs > 0) {
    if (!jobs || !cameras || !xyz_ndc || !rgb) return PGDVS_E_BADARG;
    UwpParams p;
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
    k_uwp<<<(unsigned)L.n_tiles, kUwpThreads, 0, stream>>>(p);
    if (int rc = check_launch()) return rc;
  }
  if (n_views > 0) {
    k_uwp_finalize<<<(n_views + 127) / 128, 128, 0, stream>>>(
        jobs, n_jobs, n_views, reinterpret_cast<const int64_t*>(ws + L.off_job_start), first_idx,
        num_points, total_points);
    if (int rc = check_launch()) return rc;
  }
  return PGDVS_OK;
}
