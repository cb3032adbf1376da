// Track branch: turn 2-D point tracks into a 3-D cloud at the target time.
//
// Replaces PGDVSDynamicTrackRenderer.compute_pcl_for_tgt up to its KNN filters
// (pgdvs_renderer_dyn_track.py:98-284): keep tracks that are invisible in the temporally
// closest source frames and visible in >= 2 others (:116-127); for each, take the two visible
// frames whose time is closest to the target (:146-155); sample colour (bilinear,
// align_corners=True) and depth (nearest, align_corners=False) at the track position in those
// frames (:204-229); unproject with (c2w @ K^-1) (:239-252); colour = mean of the two (:271-276);
// position = p_a + (p_b - p_a) * (t - t_a) / (t_b - t_a + 1e-8) (:278-284, extrapolates).
// The reference does this with a python loop over the frames (a host sync per iteration); here
// it is one thread per track, with the same ordered compaction as the uwp stage
// (count -> scan -> write) so the output order equals the reference's boolean-mask order.
#include "common.cuh"

namespace pgdvs {

int scan_exclusive_inplace(int* data, int64_t n_tiles, unsigned long long* state, int* ticket,
                           cudaStream_t stream);

constexpr int kTrackThreads = 256;

struct TrackParams {
  const float* tracks;      // [Q,F,2] (col, row)
  const uint8_t* visibles;  // [Q,F]
  int64_t Q;
  int F, H, W;
  const PgdvsTrackFrame* frames;
  uint32_t closest_mask, real_mask;
  float time_tgt;
  int* tile_off;  // [n_tiles + 1]
  float* pcl;
  float* rgb;
  int32_t* track_id;
  int64_t* count;
};

__device__ __forceinline__ uint32_t visible_bits(const TrackParams& p, int64_t q) {
  uint32_t v = 0;
  for (int f = 0; f < p.F; ++f)
    if (p.visibles[q * p.F + f]) v |= 1u << f;
  return v;
}

__device__ __forceinline__ bool track_valid(const TrackParams& p, uint32_t vis) {
  const bool invisible_in_closest = (vis & p.closest_mask) == 0;            // all(~vis[closest])
  const bool visible_enough = __popc(vis & p.real_mask) >= 2;               // sum(vis[real]) >= 2
  return invisible_in_closest && visible_enough;
}

__global__ void __launch_bounds__(kTrackThreads) k_track_count(const __grid_constant__ TrackParams p) {
  __shared__ int s_warp[kTrackThreads / 32];
  const int64_t q = (int64_t)blockIdx.x * kTrackThreads + threadIdx.x;
  const bool ok = (q < p.Q) && track_valid(p, visible_bits(p, q));
  const unsigned m = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = __popc(m);
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < kTrackThreads / 32; ++w) t += s_warp[w];
    p.tile_off[blockIdx.x] = t;
  }
}

// one frame's contribution: world point and colour of the track position in frame f
__device__ __forceinline__ void sample_frame(const TrackParams& p, const PgdvsTrackFrame& fr, float u, float v,
                                             float pt[3], float col[3]) {
  const float Wf = (float)p.W, Hf = (float)p.H;
  // grid = 2*uv/(W,H) - 1  (pgdvs_renderer_dyn_track.py:199-202)
  const float gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, u), Wf), 1.0f);
  const float gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, v), Hf), 1.0f);
  // colour: bilinear, align_corners=True  ->  ix = (g + 1) / 2 * (size - 1), zeros padding
  const float ix = __fmul_rn(__fadd_rn(gx, 1.0f), __fmul_rn(0.5f, Wf - 1.0f));
  const float iy = __fmul_rn(__fadd_rn(gy, 1.0f), __fmul_rn(0.5f, Hf - 1.0f));
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float tw = ix - x0f, te = 1.0f - tw, tn = iy - y0f, ts = 1.0f - tn;
  const float wgt[4] = {ts * te, ts * tw, tn * te, tn * tw};
  const int x0 = (int)x0f, y0 = (int)y0f;
  col[0] = col[1] = col[2] = 0.0f;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int xs = x0 + (t & 1), ys = y0 + (t >> 1);
    if (xs >= 0 && xs < p.W && ys >= 0 && ys < p.H) {
      const float* c = fr.rgb + ((int64_t)ys * p.W + xs) * 3;
      col[0] += __ldg(c) * wgt[t];
      col[1] += __ldg(c + 1) * wgt[t];
      col[2] += __ldg(c + 2) * wgt[t];
    }
  }
  // depth: nearest, align_corners=False  ->  ix = ((g + 1) * size - 1) / 2, round half to even
  const float nx = nearbyintf(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.0f), Wf), 1.0f), 0.5f));
  const float ny = nearbyintf(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.0f), Hf), 1.0f), 0.5f));
  float d = 0.0f;
  if (nx >= 0.0f && nx <= Wf - 1.0f && ny >= 0.0f && ny <= Hf - 1.0f)
    d = __ldg(fr.depth + (int64_t)ny * p.W + (int64_t)nx);
  // rays_d = (c2w[:3,:3] @ K^-1) @ [u, v, 1];  point = o + d * depth
  pt[0] = fr.o[0] + (fr.M[0] * u + fr.M[1] * v + fr.M[2]) * d;
  pt[1] = fr.o[1] + (fr.M[3] * u + fr.M[4] * v + fr.M[5]) * d;
  pt[2] = fr.o[2] + (fr.M[6] * u + fr.M[7] * v + fr.M[8]) * d;
}

__global__ void __launch_bounds__(kTrackThreads) k_track_main(const __grid_constant__ TrackParams p) {
  __shared__ int s_warp[kTrackThreads / 32];
  const int64_t q = (int64_t)blockIdx.x * kTrackThreads + threadIdx.x;
  const uint32_t vis = (q < p.Q) ? visible_bits(p, q) : 0u;
  const bool ok = (q < p.Q) && track_valid(p, vis);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned m = __ballot_sync(0xffffffffu, ok);
  if (lane == 0) s_warp[warp] = __popc(m);
  __syncthreads();
  int off = __ldg(p.tile_off + blockIdx.x);
  for (int w = 0; w < warp; ++w) off += s_warp[w];
  off += __popc(m & ((1u << lane) - 1u));
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0 && p.count) {
    int t = 0;
    for (int w = 0; w < kTrackThreads / 32; ++w) t += s_warp[w];
    *p.count = (int64_t)__ldg(p.tile_off + blockIdx.x) + t;
  }
  if (!ok) return;

  // the two visible frames closest in time to the target (ties: lower frame index first)
  int fa = -1, fb = -1;
  float da = __builtin_huge_valf(), db = __builtin_huge_valf();
  for (int f = 0; f < p.F; ++f) {
    if (!((vis >> f) & 1u)) continue;
    const float d = fabsf(p.frames[f].time - p.time_tgt);
    if (d < da) {
      fb = fa;
      db = da;
      fa = f;
      da = d;
    } else if (d < db) {
      fb = f;
      db = d;
    }
  }
  const float* tr = p.tracks + q * p.F * 2;
  float pa[3], pb[3], ca[3], cb[3];
  sample_frame(p, p.frames[fa], tr[fa * 2], tr[fa * 2 + 1], pa, ca);
  sample_frame(p, p.frames[fb], tr[fb * 2], tr[fb * 2 + 1], pb, cb);
  const float ta = p.frames[fa].time, tb = p.frames[fb].time;
  const float ratio = (p.time_tgt - ta) / (tb - ta + 1e-8f);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    p.pcl[(int64_t)off * 3 + i] = pa[i] + (pb[i] - pa[i]) * ratio;
    p.rgb[(int64_t)off * 3 + i] = (ca[i] + cb[i]) * 0.5f;  // torch.mean over the two frames
  }
  if (p.track_id) p.track_id[off] = (int32_t)q;
}

struct TrackLayout {
  int64_t n_tiles, n_scan_tiles;
  size_t off_tile_off, off_state, off_ticket, total;
};

static inline TrackLayout make_track_layout(int64_t Q) {
  TrackLayout L;
  L.n_tiles = (Q + kTrackThreads - 1) / kTrackThreads;
  if (L.n_tiles < 1) L.n_tiles = 1;
  L.n_scan_tiles = (L.n_tiles + 1 + kScanTile - 1) / kScanTile;
  size_t o = 0;
  L.off_tile_off = o;
  o = align256(o + sizeof(int) * (size_t)(L.n_scan_tiles * kScanTile));
  L.off_state = o;
  o = align256(o + sizeof(unsigned long long) * (size_t)L.n_scan_tiles);
  L.off_ticket = o;
  o = align256(o + 256);
  L.total = o;
  return L;
}

}  // namespace pgdvs

using namespace pgdvs;

extern "C" int pgdvs_track_workspace_bytes(int64_t Q, size_t* bytes) {
  if (!bytes || Q < 0) return PGDVS_E_BADARG;
  *bytes = make_track_layout(Q).total;
  return PGDVS_OK;
}

extern "C" int pgdvs_track_points(const float* tracks, const uint8_t* visibles, int64_t Q, int F,
                                  const PgdvsTrackFrame* frames_dev, uint32_t closest_mask,
                                  uint32_t real_mask, float time_tgt, int H, int W, float* pcl, float* rgb,
                                  int32_t* track_id, int64_t* count_dev, void* workspace,
                                  size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (Q < 0 || F < 1 || F > 32 || H <= 0 || W <= 0 || !workspace || !count_dev) return PGDVS_E_BADARG;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return PGDVS_E_ALIGN;
  TrackLayout L = make_track_layout(Q);
  if (workspace_bytes < L.total) return PGDVS_E_WORKSPACE;
  char* ws = static_cast<char*>(workspace);
  cudaError_t e = cudaMemsetAsync(ws, 0, L.total, stream);
  if (e != cudaSuccess) return (int)e;
  if (Q == 0) {
    e = cudaMemsetAsync(count_dev, 0, sizeof(int64_t), stream);
    return e == cudaSuccess ? PGDVS_OK : (int)e;
  }
  if (!tracks || !visibles || !frames_dev || !pcl || !rgb) return PGDVS_E_BADARG;
  TrackParams p = {};
  p.tracks = tracks;
  p.visibles = visibles;
  p.Q = Q;
  p.F = F;
  p.H = H;
  p.W = W;
  p.frames = frames_dev;
  p.closest_mask = closest_mask;
  p.real_mask = real_mask;
  p.time_tgt = time_tgt;
  p.tile_off = reinterpret_cast<int*>(ws + L.off_tile_off);
  p.pcl = pcl;
  p.rgb = rgb;
  p.track_id = track_id;
  p.count = count_dev;
  k_track_count<<<(unsigned)L.n_tiles, kTrackThreads, 0, stream>>>(p);
  if (int rc = check_launch()) return rc;
  if (int rc = scan_exclusive_inplace(p.tile_off, L.n_scan_tiles,
                                      reinterpret_cast<unsigned long long*>(ws + L.off_state),
                                      reinterpret_cast<int*>(ws + L.off_ticket), stream))
    return rc;
  k_track_main<<<(unsigned)L.n_tiles, kTrackThreads, 0, stream>>>(p);
  return check_launch();
}
