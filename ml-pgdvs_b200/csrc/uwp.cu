// Fused unproject -> flow-warp -> time-lerp -> project stage.
//
// Turns a batch of (target view, source-frame pair) jobs into a packed NDC point cloud ready
// for the rasterizer.  It replaces ~25 separate torch kernels and six boolean-index
// compactions (each a host sync) of the reference:
//   get_batched_rays                 pgdvs_renderer_base.py:17-57
//   compute_dyn_pcl (geometry part)  pgdvs_renderer_dyn.py:304-388
//   w2c / camera / transform         pgdvs_renderer_dyn.py:676-687 + PointsRasterizer.transform
//
// Survivors must come out in the reference's order (job-major, row-major source pixels) so
// that packed indices — and therefore the rasterizer's idx output — equal the reference's
// boolean-mask compaction.  That ordered compaction is done without any spin-waiting:
//   k_uwp_count : per 1024-pixel tile, how many pixels survive (mask / occlusion / in-bounds /
//                 keep tests; float4-vectorised coalesced loads, warp-shuffle reduction)
//   k_scan      : exclusive scan of the tile counts (bin.cu)
//   k_uwp       : each thread owns 4 consecutive pixels; a warp-shuffle + shared-memory prefix
//                 sum orders survivors inside the tile, the scanned tile offset orders tiles.
// Frame-2 colour (bilinear) and depth (nearest) are gathered from a packed (r,g,b,depth)
// float4 plane when the caller provides one: the nearest pixel is always one of the four
// bilinear taps, so 4 x 128-bit loads replace 13 scalar ones.  In the fused mode
// (pgdvs_uwp_bin) the kernel also files every point under its raster cell (one RED atomic),
// which removes the separate counting pass over the cloud, and leaves a 28-byte packed-order
// record (x, y, z, cell | r, g, b) per point for k_fill_pre to scatter into cell order.
#include "common.cuh"

namespace pgdvs {

int bin_scan_fill_fused(char* ws, const BinLayout& L, const FusedTail& T, int64_t capacity,
                        const int64_t* total_dev, const PgdvsUwpJob* jobs, int n_jobs, int n_views,
                        int tiles_per_job, const int* tile_off, cudaStream_t stream);
int scan_exclusive_inplace(int* data, int64_t n_tiles, unsigned long long* state, int* ticket,
                           cudaStream_t stream);

#ifndef PGDVS_UWP_THREADS
#define PGDVS_UWP_THREADS 256
#endif
#ifndef PGDVS_UWP_PIX
#define PGDVS_UWP_PIX 4
#endif
constexpr int kUwpThreads = PGDVS_UWP_THREADS;
constexpr int kUwpPix = PGDVS_UWP_PIX;                   // pixels per thread
constexpr int kUwpTile = kUwpThreads * kUwpPix;          // pixels per tile
static_assert(kUwpTile == kUwpTilePixels, "bin.cu's tile-major fill assumes this tile size");
static_assert((kUwpThreads / 32) * kUwpPix == 32 && (kUwpPix % 2) == 0,
              "one warp scans the (pixel slot, warp) survivor counts: 32 of them; pixels are processed in pairs");

struct UwpParams {
  const PgdvsUwpJob* jobs;
  const PgdvsCamera* cams;
  int n_jobs, H, W;
  int tiles_per_job;
  int64_t n_tiles;
  int* tile_off;     // [n_tiles + 1]: counts, then exclusive offsets; [n_tiles] = total
  const int32_t* group_first;    // [n_groups + 1] or null (every job its own group)
  const int32_t* group_members;  // [n_jobs] job indices, grouped
  float* xyz_ndc;    // [cap,3] or null
  float* rgb;        // [cap,3] or null
  float* xyz_world;  // [cap,3] or null
  int32_t* src_pix;  // [cap] or null
  float* world_by_pixel;  // [n_jobs, H*W, 3] or null: the world point of every SURVIVING source pixel at
                          // its own pixel slot (not compacted; the caller pre-fills the array)
  // fused binning (all null/0 in the plain mode)
  CellGrid g;
  int* cell_count;
  uint32_t* zrange;  // [2 * n_views] per-view range of the filed z patterns (common.cuh)
  float4* preA;      // [cap] (x_ndc, y_ndc, z, cell id)  in packed (reference) order
  float4* preB;      // [cap * 3 floats] (r, g, b), unpadded
};

// torch grid_sample(align_corners=False) source index of pixel coordinate c on an axis of
// `size` pixels, as compute_dyn_pcl builds it (pgdvs_renderer_dyn.py:341):
//   g = 2*c/size - 1 ;  ix = ((g + 1) * size - 1) / 2          (every op rounded to fp32;
// identical to ATen's CPU `(g + 1) * (size/2) - 0.5`, checked in tests)
__device__ __forceinline__ float grid_unnormalize(float c, float size) {
#ifdef PGDVS_EXP_FAST_DIV
  const float g = __fsub_rn(__fdividef(__fmul_rn(2.0f, c), size), 1.0f);
#else
  const float g = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, c), size), 1.0f);
#endif
  return __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.0f), size), 1.0f), 0.5f);
}

// Thread t of a tile owns the 4 pixels  tile0 + k*256 + t  (k = 0..3): for a fixed k the 32
// lanes of a warp touch 32 CONSECUTIVE source pixels, so the streaming reads (mask, flow as
// float2, depth) are perfectly coalesced, the frame-2 taps of neighbouring lanes share
// sectors, and — because survivors keep pixel order — consecutive lanes write consecutive
// float4 records.
struct PixelSet {
  int64_t pix[kUwpPix];
  int u[kUwpPix], v[kUwpPix];
  float2 flow[kUwpPix];
  unsigned valid;  // bit k: pixel k survives
};

// Which of this thread's pixels survive (pgdvs_renderer_dyn.py:304-316, 438-440).
__device__ __forceinline__ void pixel_validity(const UwpParams& p, const PgdvsUwpJob& J, int64_t tile0,
                                               PixelSet& s) {
  const int64_t HW = (int64_t)p.H * p.W;
  s.valid = 0;
  const int64_t first = tile0 + threadIdx.x;
  int v = (int)(first / p.W), u = (int)(first - (int64_t)v * p.W);
#pragma unroll
  for (int k = 0; k < kUwpPix; ++k) {
    const int64_t pix = first + (int64_t)k * kUwpThreads;
    s.pix[k] = pix;
    s.u[k] = u;
    s.v[k] = v;
    s.flow[k] = make_float2(0.f, 0.f);
    if (pix < HW) {
      const float m = __ldg(J.mask1 + pix);
      const float2 fl = __ldg(reinterpret_cast<const float2*>(J.flow12) + pix);
      s.flow[k] = fl;
      bool ok = (m != 0.0f);                                                  // dyn_mask.bool()
      if (J.occ12 != nullptr) ok = ok && !(__ldg(J.occ12 + pix) > 0.0f);      // ~(occ > 0) & mask
      const float u2 = __fadd_rn((float)u, fl.x), v2 = __fadd_rn((float)v, fl.y);
      ok = ok && (u2 >= 0.0f) && (u2 <= (float)(p.W - 1)) && (v2 >= 0.0f) && (v2 <= (float)(p.H - 1));
      if (ok && J.keep != nullptr) ok = (J.keep[pix] != 0);
      if (ok) s.valid |= 1u << k;
    }
    u += kUwpThreads;  // next pixel of this thread is 256 further along the row-major order
    while (u >= p.W) {
      u -= p.W;
      ++v;
    }
  }
}

// Job groups: jobs that share the source pair, the pair geometry and the lerp weights differ
// only in the target camera (the NVIDIA benchmark renders all 12 cameras of a time step:
// datasets/nvidia_eval.py:53).  Validity, frame-2 gathers, bilinear colour and the world point
// are then computed ONCE per source pixel and only projection + filing + stores repeat per
// member.  Without grouping (group_first == NULL) every job is its own group.
__device__ __forceinline__ int group_size(const UwpParams& p, int g) {
  return p.group_first ? (p.group_first[g + 1] - p.group_first[g]) : 1;
}
__device__ __forceinline__ int group_member(const UwpParams& p, int g, int m) {
  return p.group_first ? p.group_members[p.group_first[g] + m] : g;
}

constexpr int kZChunk = 32;  // members of a job group staged / z ranges flushed at a time

// what one member of a job group adds to the shared source pair
struct UwpMember {
  PgdvsCamera cam;  // 16 floats
  int view, job, tile_base, pad;
};
constexpr int kMemberWords = sizeof(UwpMember) / 4;
static_assert(sizeof(PgdvsCamera) == 64 && sizeof(UwpMember) == 80, "word-wise staging below");

// members [m0, m0 + nc) of group g -> shared memory, one word per thread and step
__device__ __forceinline__ void stage_members(const UwpParams& p, int g, int jt, int m0, int nc, UwpMember* s_mem) {
  for (int idx = threadIdx.x; idx < nc * kMemberWords; idx += kUwpThreads) {
    const int mi = idx / kMemberWords, w = idx - mi * kMemberWords;
    const int job_m = group_member(p, g, m0 + mi);
    uint32_t val = 0u;
    if (w < 17) {
      const int view = __ldg(&p.jobs[job_m].view);
      val = (w < 16) ? __ldg(reinterpret_cast<const uint32_t*>(p.cams + view) + w) : (uint32_t)view;
    } else if (w == 17) {
      val = (uint32_t)job_m;
    } else if (w == 18) {
      val = (uint32_t)__ldg(p.tile_off + (int64_t)job_m * p.tiles_per_job + jt);
    }
    reinterpret_cast<uint32_t*>(s_mem + mi)[w] = val;
  }
}

__global__ void __launch_bounds__(kUwpThreads) k_uwp_count(const __grid_constant__ UwpParams p) {
  __shared__ int s_warp[kUwpThreads / 32];
  const int g = blockIdx.x / p.tiles_per_job;
  const int jt = blockIdx.x - g * p.tiles_per_job;
  const PgdvsUwpJob& J = p.jobs[group_member(p, g, 0)];
  PixelSet s;
  pixel_validity(p, J, (int64_t)jt * kUwpTile, s);
  int cnt = __popc(s.valid);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = cnt;
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int w = 0; w < kUwpThreads / 32; ++w) t += s_warp[w];
  // the survivors of a tile are the same for every member of the group
  for (int m = threadIdx.x; m < group_size(p, g); m += kUwpThreads)
    p.tile_off[(int64_t)group_member(p, g, m) * p.tiles_per_job + jt] = t;
}

#ifndef PGDVS_UWP_MINBLOCKS
#define PGDVS_UWP_MINBLOCKS 3
#endif
// PACKED: frame 2 comes as (r, g, b, depth) float4 (pgdvs_pack_rgbd): four 128-bit taps per pixel.
// The two tap paths are two instantiations of the whole body behind one CTA-uniform branch, so
// neither carries the other's (predicated-off) instructions.
template <bool FUSED, bool PACKED>
__device__ __forceinline__ void uwp_body(const UwpParams& p, const int g, const int jt, const int n_members,
                                         int* s_cnt, uint32_t (*s_z)[2], const PgdvsUwpJob& J, UwpMember* s_mem) {
  constexpr int kWarps = kUwpThreads / 32;

  PixelSet s;
  pixel_validity(p, J, (int64_t)jt * kUwpTile, s);
  // (frame-1 depth: issued with the validity reads, consumed after the ordering barrier)
  float d1[kUwpPix];
#pragma unroll
  for (int k = 0; k < kUwpPix; ++k) d1[k] = ((s.valid >> k) & 1u) ? __ldg(J.depth1 + s.pix[k]) : 0.0f;

  // ------------------------------------------------------------ order inside the tile:
  // warp ballots give each survivor its rank inside (k, warp); a 32-entry prefix over the
  // (k, warp) counts gives the rest
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int rank[kUwpPix];
#pragma unroll
  for (int k = 0; k < kUwpPix; ++k) {
    const unsigned m = __ballot_sync(0xffffffffu, (s.valid >> k) & 1u);
    rank[k] = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) s_cnt[k * kWarps + warp] = __popc(m);
  }
  __syncthreads();
  {
    // every warp redundantly scans the 32 counts (one shuffle scan; cheaper than a 2nd barrier)
    const int c = s_cnt[lane];
    int inc = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    const int excl = inc - c;
#pragma unroll
    for (int k = 0; k < kUwpPix; ++k) rank[k] += __shfl_sync(0xffffffffu, excl, k * kWarps + warp);
  }
  // (no early exit for threads without survivors: in the fused mode the CTA meets again at the
  //  barriers of the z-range flush; every load and store below is guarded by the validity bits)
  if (!FUSED && s.valid == 0 && n_members <= kZChunk) return;  // (larger groups meet again at the restaging barriers)

  // ------------------------------------------------------------ world point + colour, once
  const bool lerp = (J.same_time == 0);
  const float4* __restrict__ rgbd2 = reinterpret_cast<const float4*>(J.rgbd2);
  float wx[kUwpPix], wy[kUwpPix], wz[kUwpPix], cr[kUwpPix], cg[kUwpPix], cb[kUwpPix];
  // two pixels at a time: their 8 frame-2 taps are issued back to back before any is consumed.
  // The taps are NAMED scalars, not an array: with tap[kk][t] the compiler turns the nearest-tap
  // select into a dynamically indexed load and parks all eight float4 in local memory (16 STL.128 +
  // LDL per pixel pair in the round-2 build).
#pragma unroll
  for (int k0 = 0; k0 < kUwpPix; k0 += 2) {
    float u2a[2], v2a[2];
    float wg0[2], wg1[2], wg2[2], wg3[2];
    float4 ta0[2], ta1[2], ta2[2], ta3[2];
    bool nx1[2], ny1[2];  // nearest pixel = right column / lower row of the 2x2 taps
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const int k = k0 + kk;
      nx1[kk] = ny1[kk] = false;
      ta0[kk] = ta1[kk] = ta2[kk] = ta3[kk] = make_float4(0.f, 0.f, 0.f, 0.f);  // zeros padding
      if (lerp) {
        const bool ok = (s.valid >> k) & 1u;
        const float u2 = __fadd_rn((float)s.u[k], s.flow[k].x), v2 = __fadd_rn((float)s.v[k], s.flow[k].y);
        u2a[kk] = u2;
        v2a[kk] = v2;
        const float ix = grid_unnormalize(u2, (float)p.W);
        const float iy = grid_unnormalize(v2, (float)p.H);
        // depth_2: grid_sample(mode="nearest"): nearbyint (half to even), zeros padding
        const float nx = nearbyintf(ix), ny = nearbyintf(iy);
        // rgb: grid_sample(rgb_2, mode="bilinear"), zeros padding (colour comes from frame 2)
        const float x0f = floorf(ix), y0f = floorf(iy);
        const float tw = __fsub_rn(ix, x0f), te = __fsub_rn(1.0f, tw);
        const float tn = __fsub_rn(iy, y0f), ts = __fsub_rn(1.0f, tn);
        const int x0 = (int)x0f, y0 = (int)y0f;
        wg0[kk] = __fmul_rn(ts, te);
        wg1[kk] = __fmul_rn(ts, tw);
        wg2[kk] = __fmul_rn(tn, te);
        wg3[kk] = __fmul_rn(tn, tw);
        // the nearest pixel is one of the 4 bilinear taps
        nx1[kk] = (nx != x0f);
        ny1[kk] = (ny != y0f);
        const bool xin0 = x0 >= 0 && x0 < p.W, xin1 = x0 + 1 >= 0 && x0 + 1 < p.W;
        const bool yin0 = ok && y0 >= 0 && y0 < p.H, yin1 = ok && y0 + 1 >= 0 && y0 + 1 < p.H;
        const int64_t o00 = (int64_t)y0 * p.W + x0;
        auto fetch = [&](bool inb, int64_t o) -> float4 {
          if (!inb) return make_float4(0.f, 0.f, 0.f, 0.f);
          if (PACKED) return __ldg(rgbd2 + o);
          return make_float4(__ldg(J.rgb2 + o * 3), __ldg(J.rgb2 + o * 3 + 1), __ldg(J.rgb2 + o * 3 + 2), __ldg(J.depth2 + o));
        };
        ta0[kk] = fetch(yin0 && xin0, o00);
        ta1[kk] = fetch(yin0 && xin1, o00 + 1);
        ta2[kk] = fetch(yin1 && xin0, o00 + p.W);
        ta3[kk] = fetch(yin1 && xin1, o00 + p.W + 1);
      }
    }
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const int k = k0 + kk;
      const float u = (float)s.u[k], v = (float)s.v[k];
      // rays_d = (c2w[:3,:3] @ K^-1) @ [u, v, 1]
      wx[k] = J.o1[0] + (J.M1[0] * u + J.M1[1] * v + J.M1[2]) * d1[k];
      wy[k] = J.o1[1] + (J.M1[3] * u + J.M1[4] * v + J.M1[5]) * d1[k];
      wz[k] = J.o1[2] + (J.M1[6] * u + J.M1[7] * v + J.M1[8]) * d1[k];
      cr[k] = cg[k] = cb[k] = 0.0f;
      if (!((s.valid >> k) & 1u)) continue;
      if (!lerp) {
        const float* c1 = J.rgb1 + s.pix[k] * 3;
        cr[k] = __ldg(c1);
        cg[k] = __ldg(c1 + 1);
        cb[k] = __ldg(c1 + 2);
      } else {
        // (same order of additions as the tap loop t = 0..3 it replaces)
        cr[k] = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(0.0f, __fmul_rn(ta0[kk].x, wg0[kk])), __fmul_rn(ta1[kk].x, wg1[kk])),
                                    __fmul_rn(ta2[kk].x, wg2[kk])), __fmul_rn(ta3[kk].x, wg3[kk]));
        cg[k] = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(0.0f, __fmul_rn(ta0[kk].y, wg0[kk])), __fmul_rn(ta1[kk].y, wg1[kk])),
                                    __fmul_rn(ta2[kk].y, wg2[kk])), __fmul_rn(ta3[kk].y, wg3[kk]));
        cb[k] = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(0.0f, __fmul_rn(ta0[kk].z, wg0[kk])), __fmul_rn(ta1[kk].z, wg1[kk])),
                                    __fmul_rn(ta2[kk].z, wg2[kk])), __fmul_rn(ta3[kk].z, wg3[kk]));
        const float dtop = nx1[kk] ? ta1[kk].w : ta0[kk].w, dbot = nx1[kk] ? ta3[kk].w : ta2[kk].w;
        const float dep2 = ny1[kk] ? dbot : dtop;
        // pcl_2 = o_2 + (R_2 @ (K_2^-1 @ [u2, v2, 1])) * depth_2
        const float u2 = u2a[kk], v2 = v2a[kk];
        const float kx = J.K2inv[0] * u2 + J.K2inv[1] * v2 + J.K2inv[2];
        const float ky = J.K2inv[3] * u2 + J.K2inv[4] * v2 + J.K2inv[5];
        const float kz = J.K2inv[6] * u2 + J.K2inv[7] * v2 + J.K2inv[8];
        const float qx = J.o2[0] + (J.R2[0] * kx + J.R2[1] * ky + J.R2[2] * kz) * dep2;
        const float qy = J.o2[1] + (J.R2[3] * kx + J.R2[4] * ky + J.R2[5] * kz) * dep2;
        const float qz = J.o2[2] + (J.R2[6] * kx + J.R2[7] * ky + J.R2[8] * kz) * dep2;
        wx[k] = J.w1 * wx[k] + J.w2 * qx;
        wy[k] = J.w1 * wy[k] + J.w2 * qy;
        wz[k] = J.w1 * wz[k] + J.w2 * qz;
      }
    }
  }

  // ------------------------------------------------------------ per member: project, file, store
  for (int m0 = 0; m0 < n_members; m0 += kZChunk) {
    const int nc = min(kZChunk, n_members - m0);
    if (m0 > 0) {  // (the barriers of the previous chunk's flush protect s_mem)
      stage_members(p, g, jt, m0, nc, s_mem);
      __syncthreads();
    }
    for (int mi = 0; mi < nc; ++mi) {
      const UwpMember& M = s_mem[mi];
      const int job_m = M.job;
      const int view = M.view;
      const PgdvsCamera& cam = M.cam;
      const int tile_base = M.tile_base;
      uint32_t z_nlo = 0u, z_hi = 0u;
#pragma unroll
      for (int k = 0; k < kUwpPix; ++k) {
        if (!((s.valid >> k) & 1u)) continue;
        const float3 ndc = world_to_ndc(cam, wx[k], wy[k], wz[k]);
        const int64_t out = (int64_t)tile_base + rank[k];
        if (FUSED) {
          const int cell = point_cell(p.g, view, ndc.x, ndc.y, ndc.z);
          if (cell >= 0) {
            atomicAdd(p.cell_count + cell, 1);  // result unused -> RED
            const uint32_t zb = z_pattern(ndc.z);
            z_nlo = max(z_nlo, ~zb);
            z_hi = max(z_hi, zb);
          }
          // packed-order record: (x, y, z, cell) + (r, g, b); the packed index is the position itself
          p.preA[out] = make_float4(ndc.x, ndc.y, ndc.z, __int_as_float(cell));
          float* pb = reinterpret_cast<float*>(p.preB) + out * 3;
          pb[0] = cr[k];
          pb[1] = cg[k];
          pb[2] = cb[k];
        }
        if (p.xyz_ndc) {
          p.xyz_ndc[out * 3 + 0] = ndc.x;
          p.xyz_ndc[out * 3 + 1] = ndc.y;
          p.xyz_ndc[out * 3 + 2] = ndc.z;
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
        if (p.src_pix) p.src_pix[out] = (int32_t)s.pix[k];
        if (p.world_by_pixel) {
          float* wp = p.world_by_pixel + ((int64_t)job_m * p.H * p.W + s.pix[k]) * 3;
          wp[0] = wx[k];
          wp[1] = wy[k];
          wp[2] = wz[k];
        }
      }
      if (FUSED) {
        // z range of this member's view: warp reduction -> shared-memory max
        z_nlo = __reduce_max_sync(0xffffffffu, z_nlo);
        z_hi = __reduce_max_sync(0xffffffffu, z_hi);
        if ((threadIdx.x & 31) == 0 && (z_nlo | z_hi) != 0u) {
          atomicMax(&s_z[mi][0], z_nlo);
          atomicMax(&s_z[mi][1], z_hi);
        }
      }
    }
    if (FUSED) {
      // the CTA publishes the ranges of the chunk's views (a look at the current value, an atomic
      // only if wider)
      __syncthreads();
      if ((int)threadIdx.x < nc) {
        const int vm = s_mem[threadIdx.x].view;
        const uint32_t nlo = s_z[threadIdx.x][0], hi = s_z[threadIdx.x][1];
        if ((nlo | hi) != 0u) {
          const uint2 cur = __ldcg(reinterpret_cast<const uint2*>(p.zrange) + vm);
          if (nlo > cur.x) atomicMax(p.zrange + 2 * vm, nlo);
          if (hi > cur.y) atomicMax(p.zrange + 2 * vm + 1, hi);
        }
        s_z[threadIdx.x][0] = 0u;
        s_z[threadIdx.x][1] = 0u;
      }
      __syncthreads();
    } else if (m0 + kZChunk < n_members) {
      __syncthreads();  // s_mem is about to be restaged
    }
  }
}

template <bool FUSED>
__global__ void __launch_bounds__(kUwpThreads, PGDVS_UWP_MINBLOCKS) k_uwp(const __grid_constant__ UwpParams p) {
  __shared__ int s_cnt[kUwpPix * (kUwpThreads / 32)];  // survivors of (k, warp), k-major == pixel order
  // fused mode: range of the z patterns this CTA files under each member's view (common.cuh),
  // kZChunk members at a time: (max ~bits, max bits), zero == empty
  __shared__ uint32_t s_z[kZChunk][2];
  // The job (pointers, pair geometry, lerp weights: 55 words) and, kZChunk members at a time, what
  // each member of the group adds (target camera, view index, output offset of this tile) are
  // fetched ONCE per CTA by as many threads as there are words and read back as shared-memory
  // broadcasts: one memory round trip instead of a chain of dependent uniform loads per use (job
  // -> view -> camera -> offset, per member), and the 35 geometry constants no longer sit in
  // registers for the whole kernel.
  __shared__ __align__(16) PgdvsUwpJob s_job;
  __shared__ __align__(16) UwpMember s_mem[kZChunk];
  if (FUSED && threadIdx.x < 2 * kZChunk) (&s_z[0][0])[threadIdx.x] = 0u;  // (published by the barrier below)
  const int g = blockIdx.x / p.tiles_per_job;
  const int jt = blockIdx.x - g * p.tiles_per_job;
  const int n_members = group_size(p, g);
  {
    static_assert(sizeof(PgdvsUwpJob) % 4 == 0 && sizeof(PgdvsUwpJob) / 4 <= kUwpThreads, "one word per thread");
    const uint32_t* src = reinterpret_cast<const uint32_t*>(p.jobs + group_member(p, g, 0));
    if (threadIdx.x < sizeof(PgdvsUwpJob) / 4)
      reinterpret_cast<uint32_t*>(&s_job)[threadIdx.x] = __ldg(src + threadIdx.x);
  }
  stage_members(p, g, jt, 0, min(kZChunk, n_members), s_mem);
  __syncthreads();
  if (s_job.rgbd2 != nullptr)  // (CTA-uniform)
    uwp_body<FUSED, true>(p, g, jt, n_members, s_cnt, s_z, s_job, s_mem);
  else
    uwp_body<FUSED, false>(p, g, jt, n_members, s_cnt, s_z, s_job, s_mem);
}

// per-view first index / count / total from the scanned tile offsets (jobs sorted by view)
__global__ void k_uwp_finalize(const PgdvsUwpJob* jobs, int n_jobs, int n_views, const int* tile_off,
                               int tiles_per_job, int64_t* first_idx, int64_t* num_points,
                               int64_t* total) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v == 0 && total) *total = tile_off[(int64_t)n_jobs * tiles_per_job];
  if (v >= n_views) return;
  int a = n_jobs, b = n_jobs;
  for (int j = n_jobs - 1; j >= 0; --j) {
    const int jv = jobs[j].view;
    if (jv >= v) a = j;
    if (jv >= v + 1) b = j;
  }
  const int64_t sa = tile_off[(int64_t)a * tiles_per_job], sb = tile_off[(int64_t)b * tiles_per_job];
  if (first_idx) first_idx[v] = sa;
  if (num_points) num_points[v] = sb - sa;
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
  int64_t n_tiles;       // pixel tiles
  int64_t n_scan_tiles;  // k_scan tiles covering n_tiles + 1 ints
  size_t off_tile_off, off_state, off_ticket, off_zero_end, total;
};

static inline UwpLayout make_uwp_layout(int n_jobs, int H, int W) {
  UwpLayout L;
  const int64_t HW = (int64_t)H * W;
  L.tiles_per_job = (int)((HW + kUwpTile - 1) / kUwpTile);
  L.n_tiles = (int64_t)L.tiles_per_job * n_jobs;
  L.n_scan_tiles = (L.n_tiles + 1 + kScanTile - 1) / kScanTile;
  size_t o = 0;
  L.off_tile_off = o;
  o = align256(o + sizeof(int) * (size_t)(L.n_scan_tiles * kScanTile));
  L.off_state = o;
  o = align256(o + sizeof(unsigned long long) * (size_t)L.n_scan_tiles);
  L.off_ticket = o;
  o = align256(o + 256);
  L.off_zero_end = o;
  L.total = o;
  return L;
}

static int run_uwp(const PgdvsUwpJob* jobs, int n_jobs, const PgdvsCamera* cameras, int n_views, int H,
                   int W, float* xyz_ndc, float* rgb, float* xyz_world, int32_t* src_pix,
                   int64_t* first_idx, int64_t* num_points, int64_t* total_points, char* ws,
                   const UwpLayout& L, const CellGrid* grid, int* cell_count, uint32_t* zrange, float4* preA,
                   float4* preB, const int32_t* group_first, const int32_t* group_members, int n_groups,
                   cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(ws, 0, L.off_zero_end, stream);
  if (e != cudaSuccess) return (int)e;
  int* tile_off = reinterpret_cast<int*>(ws + L.off_tile_off);
  const bool grouped = group_first != nullptr && group_members != nullptr && n_groups > 0;
  const unsigned grid_blocks = (unsigned)((int64_t)(grouped ? n_groups : n_jobs) * L.tiles_per_job);
  if (n_jobs > 0) {
    UwpParams p = {};
    p.group_first = grouped ? group_first : nullptr;
    p.group_members = grouped ? group_members : nullptr;
    p.jobs = jobs;
    p.cams = cameras;
    p.n_jobs = n_jobs;
    p.H = H;
    p.W = W;
    p.tiles_per_job = L.tiles_per_job;
    p.n_tiles = L.n_tiles;
    p.tile_off = tile_off;
    p.xyz_ndc = xyz_ndc;
    p.rgb = rgb;
    p.xyz_world = xyz_world;
    p.src_pix = src_pix;
    k_uwp_count<<<grid_blocks, kUwpThreads, 0, stream>>>(p);
    if (int rc = check_launch()) return rc;
    if (int rc = scan_exclusive_inplace(tile_off, L.n_scan_tiles,
                                        reinterpret_cast<unsigned long long*>(ws + L.off_state),
                                        reinterpret_cast<int*>(ws + L.off_ticket), stream))
      return rc;
    if (grid != nullptr) {
      p.g = *grid;
      p.cell_count = cell_count;
      p.zrange = zrange;
      p.preA = preA;
      p.preB = preB;
      k_uwp<true><<<grid_blocks, kUwpThreads, 0, stream>>>(p);
    } else {
      k_uwp<false><<<grid_blocks, kUwpThreads, 0, stream>>>(p);
    }
    if (int rc = check_launch()) return rc;
  }
  // always run: also publishes total_points (0 when there are no jobs: tile_off is zeroed)
  k_uwp_finalize<<<(n_views + 128) / 128, 128, 0, stream>>>(jobs, n_jobs, n_views, tile_off,
                                                            L.tiles_per_job, first_idx, num_points,
                                                            total_points);
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
                                            const int32_t* group_first, const int32_t* group_members,
                                            int n_groups, void* workspace, size_t workspace_bytes,
                                            void* stream_) {
  if (n_jobs < 0 || n_views < 0 || H <= 0 || W <= 0 || !workspace || n_groups < 0) return PGDVS_E_BADARG;
  if (n_views > 0 && (!first_idx || !num_points)) return PGDVS_E_BADARG;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return PGDVS_E_ALIGN;
  UwpLayout L = make_uwp_layout(n_jobs, H, W);
  if (workspace_bytes < L.total) return PGDVS_E_WORKSPACE;
  if ((int64_t)n_jobs * H * W >= (int64_t)INT32_MAX) return PGDVS_E_BADARG;
  if (n_jobs > 0 && (!jobs || !cameras || !xyz_ndc || !rgb)) return PGDVS_E_BADARG;
  return run_uwp(jobs, n_jobs, cameras, n_views, H, W, xyz_ndc, rgb, xyz_world, src_pix, first_idx,
                 num_points, total_points, static_cast<char*>(workspace), L, nullptr, nullptr, nullptr, nullptr,
                 nullptr, group_first, group_members, n_groups, (cudaStream_t)stream_);
}

extern "C" int pgdvs_uwp_world_by_pixel(const PgdvsUwpJob* jobs, int n_jobs, const PgdvsCamera* cameras,
                                        int n_views, int H, int W, float* world, void* workspace,
                                        size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_jobs < 0 || n_views <= 0 || H <= 0 || W <= 0 || !workspace) return PGDVS_E_BADARG;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return PGDVS_E_ALIGN;
  UwpLayout L = make_uwp_layout(n_jobs, H, W);
  if (workspace_bytes < L.total) return PGDVS_E_WORKSPACE;
  if (n_jobs == 0) return PGDVS_OK;
  if (!jobs || !cameras || !world) return PGDVS_E_BADARG;
  char* ws = static_cast<char*>(workspace);
  // every slot NaN (all-ones pattern) first: the kernel only writes the surviving pixels
  cudaError_t e = cudaMemsetAsync(world, 0xFF, sizeof(float) * 3 * (size_t)n_jobs * H * W, stream);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(ws, 0, L.off_zero_end, stream);  // tile offsets unused (no compaction): all zero
  if (e != cudaSuccess) return (int)e;
  UwpParams p = {};
  p.jobs = jobs;
  p.cams = cameras;
  p.n_jobs = n_jobs;
  p.H = H;
  p.W = W;
  p.tiles_per_job = L.tiles_per_job;
  p.n_tiles = L.n_tiles;
  p.tile_off = reinterpret_cast<int*>(ws + L.off_tile_off);
  p.world_by_pixel = world;
  k_uwp<false><<<(unsigned)((int64_t)n_jobs * L.tiles_per_job), kUwpThreads, 0, stream>>>(p);
  return check_launch();
}

extern "C" int pgdvs_uwp_bin_workspace_bytes(int n_jobs, int n_views, int H, int W, float radius,
                                             size_t* bytes) {
  if (!bytes || n_jobs < 0 || n_views < 0 || H <= 0 || W <= 0 || !(radius >= 0.0f))
    return PGDVS_E_BADARG;
  const int64_t cap = (int64_t)n_jobs * H * W;
  if (cap >= kMaxRecords) return PGDVS_E_BADARG;  // record float4 indices (2 per record) are int32
  BinLayout B = make_bin_layout(n_views, H, W, cap, radius);
  if (B.cells + kScanTile >= (int64_t)INT32_MAX) return PGDVS_E_BADARG;
  FusedTail T = make_fused_tail(B, cap);
  *bytes = T.total + align256(make_uwp_layout(n_jobs, H, W).total);
  return PGDVS_OK;
}

extern "C" int pgdvs_uwp_bin(const PgdvsUwpJob* jobs, int n_jobs, const PgdvsCamera* cameras,
                             int n_views, int H, int W, float radius, float* xyz_ndc, float* rgb,
                             int64_t* first_idx, int64_t* num_points, int64_t* total_points,
                             const int32_t* group_first, const int32_t* group_members, int n_groups,
                             void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_jobs < 0 || n_views < 0 || H <= 0 || W <= 0 || !workspace || !(radius >= 0.0f) || !total_points ||
      n_groups < 0)
    return PGDVS_E_BADARG;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return PGDVS_E_ALIGN;
  const int64_t cap = (int64_t)n_jobs * H * W;
  if (cap >= kMaxRecords) return PGDVS_E_BADARG;  // record float4 indices (2 per record) are int32
  BinLayout B = make_bin_layout(n_views, H, W, cap, radius);
  if (B.cells + kScanTile >= (int64_t)INT32_MAX) return PGDVS_E_BADARG;
  FusedTail T = make_fused_tail(B, cap);
  UwpLayout U = make_uwp_layout(n_jobs, H, W);
  if (workspace_bytes < T.total + align256(U.total)) return PGDVS_E_WORKSPACE;
  if (n_jobs > 0 && (!jobs || !cameras)) return PGDVS_E_BADARG;
  if (n_views == 0) return PGDVS_OK;
  char* ws = static_cast<char*>(workspace);
  // front pad, cell counters, scan state and ticket sit contiguously at the front
  cudaError_t e = cudaMemsetAsync(ws, 0, B.off_zero_end, stream);
  if (e != cudaSuccess) return (int)e;
  const CellGrid g = make_cell_grid(H, W, B.halo);
  int rc = run_uwp(jobs, n_jobs, cameras, n_views, H, W, xyz_ndc, rgb, nullptr, nullptr, first_idx,
                   num_points, total_points, ws + T.total, U, &g,
                   reinterpret_cast<int*>(ws + B.off_cells), reinterpret_cast<uint32_t*>(ws + B.off_zrange),
                   reinterpret_cast<float4*>(ws + T.off_preA), reinterpret_cast<float4*>(ws + T.off_preB),
                   group_first, group_members, n_groups, stream);
  if (rc) return rc;
  return bin_scan_fill_fused(ws, B, T, cap, total_points, jobs, n_jobs, n_views, U.tiles_per_job,
                             reinterpret_cast<const int*>(ws + T.total + U.off_tile_off), stream);
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
