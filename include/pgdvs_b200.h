/*
 * pgdvs_b200.h — C ABI of the B200-native PGDVS dynamic-content point-splat path.
 *
 * Drop-in boundary.  The reference (apple/ml-pgdvs) is pure Python and has no FFI layer of
 * its own; the seam it crosses is the pytorch3d C++ extension (`pytorch3d._C`), reached from
 *   pgdvs/renderers/pgdvs_renderer_dyn.py:684-722   (render_dyn_pcl -> PointsRenderer)
 *   pgdvs/renderers/st_geo_renderer.py:85-120
 * Each entry point below names the reference (or pytorch3d) interface it replaces.  All
 * pointers are DEVICE pointers unless stated otherwise; all work is enqueued on `stream`
 * (a cudaStream_t passed as void*); nothing is allocated behind the caller's back — scratch
 * comes from a caller-provided workspace whose size is queried first.  Every function
 * returns 0 on success, a negative PGDVS_E_* code for argument errors, or a positive
 * cudaError_t value if a launch failed.  No function throws, none synchronises the device.
 *
 * There is no CPU fallback: the library contains sm_100a code only.
 */
#ifndef PGDVS_B200_H_
#define PGDVS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGDVS_B200_ABI_VERSION 1

/* error codes (negative = argument / capacity errors detected on the host side) */
#define PGDVS_OK 0
#define PGDVS_E_BADARG (-1)      /* null pointer, non-positive size, unknown mode */
#define PGDVS_E_K_TOO_LARGE (-2) /* points_per_pixel > PGDVS_MAX_POINTS_PER_PIXEL */
#define PGDVS_E_WORKSPACE (-3)   /* workspace smaller than *_workspace_bytes() */
#define PGDVS_E_CHANNELS (-4)    /* fused compositing supports C <= PGDVS_MAX_FUSED_CHANNELS */
#define PGDVS_E_ALIGN (-5)       /* pointer not aligned as documented */
#define PGDVS_E_KNN_K (-6)       /* knn K > 64 (dyn_pcl_outlier_knn + 1 must be <= 64) */

/* pytorch3d's kMaxPointsPerPixel (rasterize_points.py): hard cap on K */
#define PGDVS_MAX_POINTS_PER_PIXEL 150
#define PGDVS_MAX_FUSED_CHANNELS 4

/* compositor selector — pytorch3d.renderer.compositing.{alpha_composite,norm_weighted_sum,
 * weighted_sum}; PGDVS uses NORM_WEIGHTED (pgdvs_renderer_dyn.py:707-709), ALPHA is the
 * commented-out alternative (:704-706). */
#define PGDVS_COMPOSITE_NONE 0
#define PGDVS_COMPOSITE_ALPHA 1
#define PGDVS_COMPOSITE_NORM_WEIGHTED 2
#define PGDVS_COMPOSITE_WEIGHTED_SUM 3

int pgdvs_abi_version(void);
const char* pgdvs_error_string(int code);
/* out[0..3] = sizeof(PgdvsCamera), sizeof(PgdvsUwpJob), offsetof(PgdvsUwpJob, M1),
 * offsetof(PgdvsUwpJob, view): lets a foreign-language binding verify its struct mirror. */
int pgdvs_struct_layout(int32_t out[4]);

/* --------------------------------------------------------------------------------------
 * 1. Binning: cell-sort a packed batch of NDC point clouds.
 *
 * Replaces nothing in the reference (it forces bin_size=0, pgdvs_renderer_dyn.py:689-695,
 * because pytorch3d's fixed-capacity bins overflow); this is the exact, overflow-free
 * count -> scan -> fill that makes the tiled rasterizer possible.
 *
 *  points        f32 [P,3]   NDC x, NDC y, view-space z  (PointsRasterizer.transform output)
 *  features      f32 [P,C]   per-point features, C <= 4, or NULL (raster-only)
 *  first_idx     i64 [N]     cloud_to_packed_first_idx
 *  num_points    i64 [N]     num_points_per_cloud
 *  radius        f32 [P] or NULL; radius_max = scalar radius (NULL case) or max(radius)
 *  workspace     >= pgdvs_bin_workspace_bytes(N,H,W,P,radius_max), 256-byte aligned
 * ------------------------------------------------------------------------------------ */
int pgdvs_bin_workspace_bytes(int N, int H, int W, int64_t P, float radius_max, size_t* bytes);

int pgdvs_bin_points(const float* points, const float* features, int C, const int64_t* first_idx,
                     const int64_t* num_points, int N, int64_t P, const float* radius,
                     float radius_max, int H, int W, void* workspace, size_t workspace_bytes,
                     void* stream);

/* --------------------------------------------------------------------------------------
 * 2. Tiled rasterize-and-composite over a binned workspace.
 *
 * Replaces `pytorch3d._C.rasterize_points(points, cloud_to_packed_first_idx,
 * num_points_per_cloud, image_size, radius, points_per_pixel, bin_size=0, ...)`
 * + `PointsRenderer.forward` weights (1 - dists/(r*r)) + the compositor
 * + `_add_background_color_to_images`, and the second all-ones render that PGDVS uses for
 * the mask (pgdvs_renderer_dyn.py:717-722), in one pass.
 *
 *  K             points_per_pixel (1..150)
 *  radius_max / per_point_radius   must repeat what pgdvs_bin_points was given (scalar radius,
 *                or 1 if a per-point radius tensor was binned)
 *  rr_weight     the fp32 value of (r*r) that PointsRenderer divides by (r*r evaluated in
 *                double, then rounded); ignored when compositor == NONE
 *  background    f32 [C] HOST pointer or NULL (zeros)
 *  idx/zbuf/dists  i32/f32/f32 [N,H,W,K] or NULL (skip writing fragments), -1 filled
 *  image         f32 [N,H,W,C] or NULL
 *  mask          f32 [N,H,W,1] or NULL: 1.0 where the same compositor applied to all-ones
 *                features is > 0
 *  static_rgb    f32 [N,H,W,C] or NULL.  If given (with image and mask non-NULL) `image`
 *                receives the blend (1-mask)*static + mask*dyn of pgdvs_renderer.py:169-172.
 *  workspace     read, and possibly REORDERED in place: when a pixel's window holds many more
 *                records than K, the records inside each small cell are put in ascending z
 *                order first so that the walk can leave a cell early.  The order of records
 *                within a cell never affects the result, so the workspace stays valid for
 *                further calls (on the same stream).
 * ------------------------------------------------------------------------------------ */
int pgdvs_rasterize_composite(void* workspace, size_t workspace_bytes, int N, int64_t P,
                              int H, int W, int K, float radius_max, int per_point_radius, int C,
                              int compositor,
                              float rr_weight, const float* background, const float* static_rgb,
                              int32_t* idx, float* zbuf, float* dists, float* image, float* mask,
                              void* stream);

/* Extended outputs of the same pass (HOST struct, may be NULL; every member may be NULL):
 *  depth     f32 [N,H,W,1]  the compositor applied to the hits' view-space z as one more feature
 *            channel — the "composited depth" of the north star (the reference never composites
 *            depth: it discards pytorch3d's zbuf, pgdvs_renderer_dyn.py:717); 0 where idx[...,0] < 0.
 *            zbuf[..., 0] is the nearest-hit depth.
 *  image_u8  u8 [N,H,W,C]   the image (after the static blend, if any) as the reference's evaluator /
 *            video writer quantise it (engines/evaluator_pgdvs.py:51-77: NaN -> 0, clamp(0,1),
 *            (x*255).byte()); `image` may then be NULL.  Saves the separate pgdvs_quantize_u8 pass.
 *  mask_u8   u8 [N,H,W,1]   mask * 255 */
typedef struct PgdvsRasterExtra {
  float* depth;
  uint8_t* image_u8;
  uint8_t* mask_u8;
} PgdvsRasterExtra;
int pgdvs_rasterize_composite_ex(void* workspace, size_t workspace_bytes, int N, int64_t P,
                                 int H, int W, int K, float radius_max, int per_point_radius, int C,
                                 int compositor, float rr_weight, const float* background,
                                 const float* static_rgb, int32_t* idx, float* zbuf, float* dists,
                                 float* image, float* mask, const PgdvsRasterExtra* extra, void* stream);

/* Developer switches of the rasterizer (kernel selection only — results are bit-identical in
 * every setting; tests and A/B harnesses use them).  which: 0 = z-sort the cells before the
 * generic kernel, 1 = force the generic kernel, 2 = k_raster_tile instead of k_raster_pair
 * (0 = k_raster_pair whenever it applies, even for launches too small to fill the GPU with it).
 * value: -1 automatic, 0 off, 1 on.  Initial values come from the environment variables
 * PGDVS_SORT_CELLS / PGDVS_RASTER_FORCE_GENERIC / PGDVS_RASTER_NO_PAIR, read once per process. */
int pgdvs_debug_switch(int which, int value);

/* --------------------------------------------------------------------------------------
 * 3. Stand-alone compositors (pytorch3d `_C.accum_alphacomposite`, `_C.accum_weightedsumnorm`,
 *    `_C.accum_weightedsum` forward):  idx i64 [N,K,H,W], alphas f32 [N,K,H,W],
 *    features f32 [C,P]  ->  out f32 [N,C,H,W].
 * ------------------------------------------------------------------------------------ */
int pgdvs_composite(const int64_t* idx, const float* alphas, const float* features, int N, int K,
                    int H, int W, int C, int64_t P, int mode, float* out, void* stream);

/* --------------------------------------------------------------------------------------
 * 4. Fused unproject -> flow-warp -> time-lerp -> project.
 *
 * Replaces PGDVSBaseRenderer.get_batched_rays (pgdvs_renderer_base.py:17-57),
 * PGDVSDynamicRenderer.compute_dyn_pcl's geometry (pgdvs_renderer_dyn.py:304-388) and
 * PointsRasterizer.transform for the camera built at pgdvs_renderer_dyn.py:684-687,
 * for a batch of (target view, source-frame pair) jobs.  Output order is the reference's:
 * view-major, job-major, then row-major source pixels (stable compaction).
 * ------------------------------------------------------------------------------------ */
typedef struct PgdvsCamera { /* pytorch3d PerspectiveCameras(in_ndc=True), row-vector convention */
  float R[9];                /* X_view = X_world @ R + T */
  float T[3];
  float focal[2];
  float p0[2];
} PgdvsCamera;

typedef struct PgdvsUwpJob {
  const float* depth1;  /* [H,W]   source frame 1 depth           */
  const float* rgb1;    /* [H,W,3]                                 */
  const float* mask1;   /* [H,W]   dynamic mask (non-zero = keep)  */
  const float* flow12;  /* [H,W,2] pixels, +u right / +v down      */
  const float* occ12;   /* [H,W] or NULL: >0 = occluded (dyn_render_use_flow_consistency) */
  const float* depth2;  /* [H,W]   frame 2 (nearest-sampled at uv2-0.5) */
  const float* rgb2;    /* [H,W,3] frame 2 (bilinear)              */
  const uint8_t* keep;  /* [H*W] or NULL: per-SOURCE-PIXEL outlier verdict (0 = drop), applied
                           after the validity test (pgdvs_renderer_dyn.py:438-440) */
  const float* rgbd2;   /* [H,W,4] or NULL: frame 2 packed as (r,g,b,depth) by pgdvs_pack_rgbd,
                           16-byte aligned; when given, rgb2/depth2 are not read */
  float M1[9];          /* c2w_1[:3,:3] @ inv(K_1[:3,:3])  (base.py:40-45) */
  float o1[3];          /* c2w_1[:3,3]                              */
  float K2inv[9];       /* inv(K_2[:3,:3])                         (dyn.py:362-365) */
  float R2[9];          /* c2w_2[:3,:3]                             */
  float o2[3];
  float w1, w2;         /* (t2-t)/(t2-t1), (t-t1)/(t2-t1)          (dyn.py:385-386) */
  int32_t same_time;    /* t1 == t2 -> pcl = pcl_1, rgb = rgb_1     (dyn.py:333-337) */
  int32_t view;         /* index into cameras[] and the per-view outputs; non-decreasing */
} PgdvsUwpJob;

int pgdvs_uwp_workspace_bytes(int n_jobs, int H, int W, size_t* bytes);

/*  jobs          device array [n_jobs], sorted by view
 *  cameras       device array [n_views]
 *  xyz_ndc       f32 [n_jobs*H*W, 3] capacity; packed NDC points (x, y, view z)
 *  rgb           f32 [n_jobs*H*W, 3] capacity
 *  xyz_world     f32 [.,3] or NULL;  src_pix i32 [.] or NULL (flat source pixel of each point)
 *  first_idx/num_points  i64 [n_views] outputs (cloud_to_packed_first_idx / num_points_per_cloud)
 *  total_points  i64 [1] device output
 *
 *  Job groups (optional; pass NULL, NULL, 0 for none): jobs that share the source pair, its
 *  geometry and the lerp weights and differ only in the target camera — e.g. the 12 cameras the
 *  NVIDIA benchmark renders per time step (datasets/nvidia_eval.py:53).  The world point and
 *  colour of a source pixel are then computed once per group and only projected per member.
 *  group_first i32 [n_groups+1] (device), group_members i32 [n_jobs] (device, job indices).
 *  Results are identical with and without grouping.
 */
int pgdvs_unproject_warp_project(const PgdvsUwpJob* jobs, int n_jobs, const PgdvsCamera* cameras,
                                 int n_views, int H, int W, float* xyz_ndc, float* rgb,
                                 float* xyz_world, int32_t* src_pix, int64_t* first_idx,
                                 int64_t* num_points, int64_t* total_points,
                                 const int32_t* group_first, const int32_t* group_members,
                                 int n_groups, void* workspace, size_t workspace_bytes, void* stream);

/* Fused variant used by the batched renderer: the same kernel additionally files every point
 * under its raster cell, then the cell counters are scanned and the records scattered, i.e.
 * on return `workspace` is in exactly the state pgdvs_bin_points would leave it in for
 *   N = n_views, P = n_jobs*H*W (capacity), radius_max = radius, features = rgb (C = 3)
 * and can be handed to pgdvs_rasterize_composite with those arguments.
 * xyz_ndc / rgb / first_idx / num_points may be NULL (not materialised). */
int pgdvs_uwp_bin_workspace_bytes(int n_jobs, int n_views, int H, int W, float radius, size_t* bytes);
int pgdvs_uwp_bin(const PgdvsUwpJob* jobs, int n_jobs, const PgdvsCamera* cameras, int n_views, int H,
                  int W, float radius, float* xyz_ndc, float* rgb, int64_t* first_idx,
                  int64_t* num_points, int64_t* total_points, const int32_t* group_first,
                  const int32_t* group_members, int n_groups, void* workspace, size_t workspace_bytes,
                  void* stream);

/* Pack source frames as (r,g,b,depth) float4 planes so that the warp stage fetches frame-2
 * colour (bilinear, pgdvs_renderer_dyn.py:350-356) and depth (nearest, :342-348) with four
 * 128-bit loads.  frames_dev: device array [n_frames]. */
typedef struct PgdvsFramePack {
  const float* rgb;   /* [H,W,3] */
  const float* depth; /* [H,W]   */
  float* rgbd;        /* [H,W,4] out, 16-byte aligned */
} PgdvsFramePack;
int pgdvs_pack_rgbd(const PgdvsFramePack* frames_dev, int n_frames, int H, int W, void* stream);

/* World -> NDC only (PointsRasterizer.transform) for an already-built cloud, e.g. the one
 * handed to render_dyn_pcl (pgdvs_renderer_dyn.py:671-724) or StaticGeoPointRenderer. */
int pgdvs_project_points(const float* xyz_world, int64_t P, const PgdvsCamera* camera_dev,
                         float* xyz_ndc, void* stream);

/* Projector.compute_projections (pgdvs/models/gnt/projector.py:41-73) — the `proj_func` the
 * dynamic renderer is constructed with (pgdvs_renderer.py:78) and calls for the softsplat flow
 * (pgdvs_renderer_dyn.py:470-473):  h = (K @ inv(c2w)) @ [xyz, 1];  uv = h.xy / clamp(h.z, min=1e-8),
 * clamped to [-1e6, 1e6];  mask = h.z > 0.
 *  xyz f32 [P,3];  proj f32 [n_cams,12] = rows 0..2 of K @ inv(c2w) (device);
 *  uv f32 [n_cams,P,2];  mask u8 [n_cams,P] (nullable). */
int pgdvs_compute_projections(const float* xyz, int64_t P, const float* proj, int n_cams, float* uv,
                              uint8_t* mask, void* stream);

/* --------------------------------------------------------------------------------------
 * 5. dyn/track merge + static blend, channels-first like the reference:
 *    mask_for_track = ~(dyn_mask>0) & (track_mask>0); rgb = (1-m)*dyn + m*track;
 *    mask = dyn|track  (pgdvs_renderer_dyn.py:229-235);  then, if static_rgb != NULL,
 *    combined = (1-mask)*static + mask*rgb (pgdvs_renderer.py:169-172).
 *    dyn_rgb/track_rgb/static_rgb [B,3,H,W]; masks [B,1,H,W]; track_* may be NULL.
 * ------------------------------------------------------------------------------------ */
int pgdvs_merge_blend(const float* dyn_rgb, const float* dyn_mask, const float* track_rgb,
                      const float* track_mask, const float* static_rgb, int B, int H, int W,
                      float* out_rgb, float* out_mask, float* out_combined, void* stream);

/* 8-bit frames as the reference's evaluator / visualizer consume them
 * (pgdvs/engines/evaluator_pgdvs.py:51-77): NaN -> 0, clamp to [0,1], (x*255).byte().
 * in f32 [n] (16-byte aligned), out u8 [n].  Used before gathering frames across GPUs. */
int pgdvs_quantize_u8(const float* in, uint8_t* out, int64_t n, void* stream);

/* --------------------------------------------------------------------------------------
 * 5b. Track branch: 2-D point tracks -> 3-D cloud at the target time.  Replaces
 *     PGDVSDynamicTrackRenderer.compute_pcl_for_tgt up to its KNN filters
 *     (pgdvs_renderer_dyn_track.py:98-284).  Tracker inference itself (TAPIR / CoTracker) is
 *     out of scope: tracks and visibilities are inputs.
 *
 *  tracks f32 [Q,F,2] (col,row); visibles u8 [Q,F]; frames_dev device array [F] (F <= 32)
 *  closest_mask / real_mask: bit f set <=> frame f is in data_for_track["idx_temporal_closest"] /
 *  ["idx_real_track"];  outputs pcl/rgb f32 [Q,3] capacity, track_id i32 [Q] or NULL, count i64 [1]
 * ------------------------------------------------------------------------------------ */
typedef struct PgdvsTrackFrame {
  const float* rgb;   /* [H,W,3] */
  const float* depth; /* [H,W]   */
  float M[9];         /* c2w[:3,:3] @ inv(K[:3,:3]) */
  float o[3];         /* c2w[:3,3] */
  float time;
  float _pad;
} PgdvsTrackFrame;
int pgdvs_track_workspace_bytes(int64_t Q, size_t* bytes);
int pgdvs_track_points(const float* tracks, const uint8_t* visibles, int64_t Q, int F,
                       const PgdvsTrackFrame* frames_dev, uint32_t closest_mask, uint32_t real_mask,
                       float time_tgt, int H, int W, float* pcl, float* rgb, int32_t* track_id,
                       int64_t* count_dev, void* workspace, size_t workspace_bytes, void* stream);

/* --------------------------------------------------------------------------------------
 * 6. Statistical outlier support ("next" row 1 of the scope table): mean squared distance
 *    from each query point to its K nearest reference points, dropping the first
 *    `skip_first` of them.  Replaces `pytorch3d.ops.knn_points(q, r, K, return_nn=True)` +
 *    `torch.mean(nn_dists[:, skip_first:], dim=1)` at pgdvs_renderer_dyn.py:405-419 and
 *    pgdvs_renderer_dyn_track.py:303-318, 345-361.  K <= 64 (PGDVS_E_KNN_K above that).  As upstream,
 *    the mean always divides by K - skip_first (knn_points zero-pads when R < K).
 *    query f32 [Q,3], ref f32 [R,3] -> mean_out f32 [Q].
 * ------------------------------------------------------------------------------------ */
int pgdvs_knn_workspace_bytes(int64_t Q, int64_t R, size_t* bytes);
int pgdvs_knn_mean_dist(const float* query, int64_t Q, const float* ref, int64_t R, int K,
                        int skip_first, float* mean_out, void* workspace, size_t workspace_bytes,
                        void* stream);
/* Full `pytorch3d.ops.knn_points` result for one cloud pair: squared distances f32 [Q,K]
 * (ascending) and reference indices i64 [Q,K]; slots beyond R hold (0, 0) like pytorch3d's
 * zero-initialised outputs.  K <= 64. */
int pgdvs_knn_points(const float* query, int64_t Q, const float* ref, int64_t R, int K,
                     float* dists_out, int64_t* idx_out, void* stream);

/* Batched statistical outlier filter without host syncs (pgdvs_renderer_dyn.py:401-457 for many
 * source pairs at once).  pgdvs_uwp_world_by_pixel runs the unproject -> warp -> lerp part of
 * pgdvs_unproject_warp_project and leaves the world point of every surviving source pixel at its
 * own pixel slot: world f32 [n_jobs, H*W, 3], NaN where the pixel does not survive (workspace as
 * pgdvs_uwp_workspace_bytes).  Each job's slice is then a fixed-size cloud for
 * pgdvs_knn_mean_dist(query = ref = slice, K = knn + 1, skip_first = 1), which gives +inf for the
 * NaN slots, and pgdvs_outlier_keep turns the statistics into the per-source-pixel verdict
 * PgdvsUwpJob.keep consumes:  thres = median(avg) + std(avg) * std_thres over the finite entries
 * (lower median and unbiased std, like torch.median / torch.std at :420-423), keep = avg < thres.
 *  avg f32 [n_clouds, n_slots]; keep u8 [n_clouds, n_slots]; thres_out f32 [n_clouds] or NULL. */
int pgdvs_uwp_world_by_pixel(const PgdvsUwpJob* jobs, int n_jobs, const PgdvsCamera* cameras, int n_views,
                             int H, int W, float* world, void* workspace, size_t workspace_bytes,
                             void* stream);
int pgdvs_outlier_keep(const float* avg, int n_clouds, int64_t n_slots, float std_thres, uint8_t* keep,
                       float* thres_out, void* stream);

/* --------------------------------------------------------------------------------------
 * 7. Softmax splatting ("next" row 3 of the scope table; the reference's default
 *    dyn_render_type).
 *    pgdvs_softsplat_forward replaces `softsplat_func.forward` (pgdvs/utils/softsplat.py:342-427,
 *    a cupy-compiled kernel upstream): tenIn f32 [N,C,H,W], tenFlow f32 [N,2,H,W] ->
 *    tenOut f32 [N,C,H,W]; every source pixel is added to the four pixels around
 *    (x + flow_x, y + flow_y) with bilinear weights; non-finite targets are skipped.
 *    pgdvs_softsplat_dyn fuses the whole dynamic branch of PGDVSDynamicRenderer.forward
 *    (pgdvs_renderer_dyn.py:157-209 with PGDVSBaseRenderer.softsplat_img,
 *    pgdvs_renderer_base.py:59-138): static regions of frame 1 replaced by `noise` (NULL = 0),
 *    back-warp of frame 2 along flow_12 (grid_sample bilinear/zeros/align_corners=True),
 *    metric = mean_c |rgb1 - warp|, weight exp(clip(-alpha*metric, -alpha, alpha)), "soft" splat
 *    of rgb and of the mask along flow_1_to_tgt, division by (sum + 1e-7), mask > 1e-3,
 *    rgb * mask.  Inputs channels-last f32 (rgb1/rgb2/noise [B,H,W,3], mask1 [B,H,W,1], flows
 *    [B,H,W,2]); outputs as upstream: out_rgb [B,3,H,W], out_mask [B,1,H,W], out_metric
 *    [B,1,H,W] (nullable).  workspace: pgdvs_softsplat_workspace_bytes, 32-byte aligned.
 * ------------------------------------------------------------------------------------ */
int pgdvs_softsplat_forward(const float* ten_in, const float* ten_flow, int N, int C, int H, int W,
                            float* ten_out, void* stream);
int pgdvs_softsplat_workspace_bytes(int B, int H, int W, size_t* bytes);
int pgdvs_softsplat_dyn(const float* rgb1, const float* mask1, const float* noise, const float* rgb2,
                        const float* flow_1_to_tgt, const float* flow_12, float alpha, int B, int H,
                        int W, float* out_rgb, float* out_mask, float* out_metric, void* workspace,
                        size_t workspace_bytes, void* stream);

/* --------------------------------------------------------------------------------------
 * 8. Mesh mode ("next" row 4 of the scope table; dyn_render_type = mesh).
 *    Replaces pytorch3d MeshRasterizer (blur_radius = 0, faces_per_pixel = 1, bin_size = 0,
 *    cull_backfaces = False, clip_barycentric_coords = False) + the reference's SimpleShader
 *    (pgdvs/utils/pytorch3d_utils.py:50-67) + the second all-ones render for the mask, as called
 *    by PGDVSDynamicRenderer.render_dyn_mesh (pgdvs_renderer_dyn.py:606-655), for ONE mesh.
 *    verts_ndc f32 [V,3] (x_ndc, y_ndc, z_view: pgdvs_project_points), faces i32 [F,3],
 *    vert_rgb f32 [V,3] (nullable when image == NULL).  Outputs (each nullable):
 *    pix_to_face i32 [H,W] (-1 = background), zbuf f32 [H,W], bary f32 [H,W,3] (-1 filled),
 *    image f32 [H,W,3] (black background), mask f32 [H,W,1].
 *    workspace: pgdvs_mesh_workspace_bytes, 8-byte aligned.
 * ------------------------------------------------------------------------------------ */
int pgdvs_mesh_workspace_bytes(int H, int W, size_t* bytes);
int pgdvs_rasterize_mesh(const float* verts_ndc, int64_t V, const int32_t* faces, int64_t F, int H,
                         int W, int perspective_correct, const float* vert_rgb, int32_t* pix_to_face,
                         float* zbuf, float* bary, float* image, float* mask, void* workspace,
                         size_t workspace_bytes, void* stream);

/* --------------------------------------------------------------------------------------
 * 9. Multi-GPU frame sink (NCCL / NVLink is used only to gather rendered frames).
 *    The gathering rank exports a device buffer over CUDA IPC; every other rank (one process per
 *    GPU) maps it and delivers its frames with copy-engine writes over NVLink — no kernel on
 *    either side.  The reference has no counterpart (ranks write PNGs, rank 0 globs them:
 *    pgdvs/engines/visualizer_pgdvs.py:160-177).
 *    pgdvs_ipc_alloc: cudaMalloc + export; handle_out = PGDVS_IPC_HANDLE_BYTES HOST bytes to ship
 *    to the peers (any byte channel).  pgdvs_ipc_open maps a peer's buffer (peer access enabled
 *    lazily), pgdvs_ipc_close unmaps it, pgdvs_ipc_free releases the owner's allocation.
 *    pgdvs_copy_async: stream-ordered device-to-device copy (local or peer), done by a DMA engine.
 * ------------------------------------------------------------------------------------ */
#define PGDVS_IPC_HANDLE_BYTES 64
int pgdvs_ipc_alloc(size_t bytes, void** dev_ptr, unsigned char* handle_out);
int pgdvs_ipc_open(const unsigned char* handle, void** dev_ptr);
int pgdvs_ipc_close(void* dev_ptr);
int pgdvs_ipc_free(void* dev_ptr);
int pgdvs_copy_async(void* dst, const void* src, size_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PGDVS_B200_H_ */
