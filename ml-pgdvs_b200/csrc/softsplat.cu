// Softmax splatting forward (pgdvs/utils/softsplat.py:280-427, the reference's default
// `dyn_render_type`), as used by PGDVSDynamicRenderer.forward (pgdvs_renderer_dyn.py:157-209)
// through PGDVSBaseRenderer.softsplat_img (pgdvs_renderer_base.py:59-138).
//
//   k_splat_nchw            softsplat_func.forward itself: every source pixel adds its value to the
//                           four pixels around (x + flow_x, y + flow_y) with bilinear weights
//                           (the reference's cupy kernel, one thread per (n, y, x) instead of per
//                           (n, c, y, x): the taps and weights are computed once for all channels).
//   k_softsplat_dyn_scatter the whole dynamic branch fused for channels-last inputs: static
//                           regions replaced by noise, back-warp of frame 2 (grid_sample,
//                           bilinear, zeros, align_corners=True), importance metric
//                           exp(clip(-alpha * mean|rgb1 - warp|)), and ONE splat of
//                           (rgb*e, mask*e, e) — the reference splats rgb and the mask in two
//                           passes with the same metric.  One 128-bit vector atomic + one scalar
//                           atomic per tap instead of 4 + 2 scalar atomics.
//   k_softsplat_dyn_resolve normalisation by (sum e + 1e-7), mask > 1e-3, rgb * mask, NCHW output.
//
// Floating-point atomics make the summation order, hence the last bits, run-dependent — exactly
// as in the reference; parity is tolerance-based (tests/test_gpu_softsplat.py).
#include "common.cuh"

namespace pgdvs {

struct SplatTaps {
  int x0, y0;        // north-west tap
  float w[4];        // nw, ne, sw, se  (softsplat.py:371-374)
  bool ok;           // finite target position
};

__device__ __forceinline__ SplatTaps splat_taps(float fx, float fy) {
  SplatTaps t;
  t.ok = isfinite(fx) && isfinite(fy);
  const float flx = floorf(fx), fly = floorf(fy);
  // clamp before the conversion: positions far outside the image have no valid tap anyway
  t.x0 = (int)fminf(fmaxf(flx, -2.0f), 1.0e9f);
  t.y0 = (int)fminf(fmaxf(fly, -2.0f), 1.0e9f);
  const float nwx = flx, nwy = fly, sex = flx + 1.0f, sey = fly + 1.0f;
  t.w[0] = __fmul_rn(__fsub_rn(sex, fx), __fsub_rn(sey, fy));
  t.w[1] = __fmul_rn(__fsub_rn(fx, nwx), __fsub_rn(sey, fy));
  t.w[2] = __fmul_rn(__fsub_rn(sex, fx), __fsub_rn(fy, nwy));
  t.w[3] = __fmul_rn(__fsub_rn(fx, nwx), __fsub_rn(fy, nwy));
  if (!(flx >= -2.0f && flx <= 1.0e9f && fly >= -2.0f && fly <= 1.0e9f)) t.ok = false;
  return t;
}

// tenIn [N,C,H,W], tenFlow [N,2,H,W] -> tenOut [N,C,H,W] (zeroed by the caller)
__global__ void __launch_bounds__(256) k_splat_nchw(const float* __restrict__ in, const float* __restrict__ flow,
                                                    float* __restrict__ out, int N, int C, int H, int W) {
  const int64_t HW = (int64_t)H * W;
  const int64_t total = (int64_t)N * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / HW);
    const int64_t px = i % HW;
    const int y = (int)(px / W), x = (int)(px % W);
    const float fx = __fadd_rn((float)x, __ldg(flow + ((int64_t)n * 2 + 0) * HW + px));
    const float fy = __fadd_rn((float)y, __ldg(flow + ((int64_t)n * 2 + 1) * HW + px));
    const SplatTaps t = splat_taps(fx, fy);
    if (!t.ok) continue;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int tx = t.x0 + (k & 1), ty = t.y0 + (k >> 1);
      if (tx < 0 || tx >= W || ty < 0 || ty >= H) continue;
      const int64_t o = (int64_t)ty * W + tx;
      for (int c = 0; c < C; ++c) {
        const float v = __ldg(in + ((int64_t)n * C + c) * HW + px);
        atomicAdd(out + ((int64_t)n * C + c) * HW + o, __fmul_rn(v, t.w[k]));
      }
    }
  }
}

// torch.linspace(-1, 1, steps) element i in fp32 (ATen: symmetric evaluation from both ends)
__device__ __forceinline__ float linspace_pm1(int i, int steps) {
  if (steps == 1) return -1.0f;
  const float step = __fdiv_rn(2.0f, (float)(steps - 1));
  return (i < steps / 2) ? __fadd_rn(-1.0f, __fmul_rn(step, (float)i))
                         : __fsub_rn(1.0f, __fmul_rn(step, (float)(steps - 1 - i)));
}

struct SoftsplatDynParams {
  const float* rgb1;    // [B,H,W,3]
  const float* mask1;   // [B,H,W,1]  valid dynamic mask of frame 1
  const float* noise;   // [B,H,W,3]  clamp(randn, 0, 1) for the static regions, or null (zeros)
  const float* rgb2;    // [B,H,W,3]
  const float* flow_t;  // [B,H,W,2]  flow frame 1 -> target view
  const float* flow_12; // [B,H,W,2]  flow frame 1 -> frame 2
  float alpha;
  int B, H, W;
  float* acc;           // [B,H,W,8]  (r e, g e, b e, m e, e, -, -, -)
  float* out_rgb;       // [B,3,H,W]
  float* out_mask;      // [B,1,H,W]
  float* out_metric;    // [B,1,H,W] or null
};

__global__ void __launch_bounds__(256) k_softsplat_dyn_scatter(const SoftsplatDynParams p) {
  const int H = p.H, W = p.W;
  const int64_t HW = (int64_t)H * W;
  const int64_t total = (int64_t)p.B * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const int64_t px = i % HW;
    const int y = (int)(px / W), x = (int)(px % W);
    const float m = __ldg(p.mask1 + i);
    float c1[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      // rgb_src_1 * mask + clamp(randn) * (1 - mask)           (pgdvs_renderer_dyn.py:181-185)
      const float nz = p.noise ? __ldg(p.noise + i * 3 + c) : 0.0f;
      c1[c] = __fadd_rn(__fmul_rn(__ldg(p.rgb1 + i * 3 + c), m), __fmul_rn(nz, __fsub_rn(1.0f, m)));
    }
    // ---- back-warp frame 2 to frame 1 (pgdvs_renderer_base.py:100-138): grid = linspace + flow /
    //      ((size - 1) / 2), grid_sample(bilinear, zeros, align_corners=True)
    const float f12x = __ldg(p.flow_12 + i * 2), f12y = __ldg(p.flow_12 + i * 2 + 1);
    const float gx = __fadd_rn(linspace_pm1(x, W), __fdiv_rn(f12x, __fdiv_rn(__fsub_rn((float)W, 1.0f), 2.0f)));
    const float gy = __fadd_rn(linspace_pm1(y, H), __fdiv_rn(f12y, __fdiv_rn(__fsub_rn((float)H, 1.0f), 2.0f)));
    const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.0f), 2.0f), (float)(W - 1));
    const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.0f), 2.0f), (float)(H - 1));
    const float x0f = floorf(ix), y0f = floorf(iy);
    const float wx1 = __fsub_rn(ix, x0f), wy1 = __fsub_rn(iy, y0f);
    const float wx0 = __fsub_rn(__fadd_rn(x0f, 1.0f), ix), wy0 = __fsub_rn(__fadd_rn(y0f, 1.0f), iy);
    const float bw[4] = {__fmul_rn(wx0, wy0), __fmul_rn(wx1, wy0), __fmul_rn(wx0, wy1), __fmul_rn(wx1, wy1)};
    float warp[3] = {0.f, 0.f, 0.f};
    if (x0f >= -1.0f && x0f <= (float)W && y0f >= -1.0f && y0f <= (float)H) {
      const int xb = (int)x0f, yb = (int)y0f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int tx = xb + (k & 1), ty = yb + (k >> 1);
        if (tx < 0 || tx >= W || ty < 0 || ty >= H) continue;
        const float* src = p.rgb2 + ((int64_t)b * HW + (int64_t)ty * W + tx) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) warp[c] = __fadd_rn(warp[c], __fmul_rn(__ldg(src + c), bw[k]));
      }
    }
    // l1_loss(reduction="none").mean(dim=1)
    const float l1 = __fdiv_rn(__fadd_rn(__fadd_rn(fabsf(__fsub_rn(c1[0], warp[0])), fabsf(__fsub_rn(c1[1], warp[1]))),
                                         fabsf(__fsub_rn(c1[2], warp[2]))), 3.0f);
    if (p.out_metric) p.out_metric[i] = l1;
    const float e = expf(fminf(fmaxf(__fmul_rn(-p.alpha, l1), -p.alpha), p.alpha));
    // ---- splat (softsplat.py:355-393)
    const float fx = __fadd_rn((float)x, __ldg(p.flow_t + i * 2));
    const float fy = __fadd_rn((float)y, __ldg(p.flow_t + i * 2 + 1));
    const SplatTaps t = splat_taps(fx, fy);
    if (!t.ok) continue;
    const float v0 = __fmul_rn(c1[0], e), v1 = __fmul_rn(c1[1], e), v2 = __fmul_rn(c1[2], e), v3 = __fmul_rn(m, e);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int tx = t.x0 + (k & 1), ty = t.y0 + (k >> 1);
      if (tx < 0 || tx >= W || ty < 0 || ty >= H) continue;
      float* a = p.acc + ((int64_t)b * HW + (int64_t)ty * W + tx) * 8;
      const float w = t.w[k];
      atomicAdd(reinterpret_cast<float4*>(a),
                make_float4(__fmul_rn(v0, w), __fmul_rn(v1, w), __fmul_rn(v2, w), __fmul_rn(v3, w)));
      atomicAdd(a + 4, __fmul_rn(e, w));
    }
  }
}

__global__ void __launch_bounds__(256) k_softsplat_dyn_resolve(const SoftsplatDynParams p) {
  const int64_t HW = (int64_t)p.H * p.W;
  const int64_t total = (int64_t)p.B * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const int64_t px = i % HW;
    const float4 a = *reinterpret_cast<const float4*>(p.acc + i * 8);
    const float den = __fadd_rn(p.acc[i * 8 + 4], 0.0000001f);  // softsplat.py:313-316
    const float mk = (__fdiv_rn(a.w, den) > 1e-3f) ? 1.0f : 0.0f;  // pgdvs_renderer_dyn.py:203
    p.out_mask[i] = mk;
    p.out_rgb[((int64_t)b * 3 + 0) * HW + px] = __fmul_rn(__fdiv_rn(a.x, den), mk);
    p.out_rgb[((int64_t)b * 3 + 1) * HW + px] = __fmul_rn(__fdiv_rn(a.y, den), mk);
    p.out_rgb[((int64_t)b * 3 + 2) * HW + px] = __fmul_rn(__fdiv_rn(a.z, den), mk);
  }
}

static inline int splat_grid(int64_t total) {
  int64_t g = (total + 255) / 256;
  if (g < 1) g = 1;
  if (g > 148 * 16) g = 148 * 16;
  return (int)g;
}

}  // namespace pgdvs

using namespace pgdvs;

extern "C" int pgdvs_softsplat_forward(const float* ten_in, const float* ten_flow, int N, int C, int H, int W,
                                       float* ten_out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (N < 0 || C < 1 || H <= 0 || W <= 0) return PGDVS_E_BADARG;
  if (N == 0) return PGDVS_OK;
  if (!ten_in || !ten_flow || !ten_out) return PGDVS_E_BADARG;
  const int64_t total = (int64_t)N * H * W;
  cudaError_t e = cudaMemsetAsync(ten_out, 0, sizeof(float) * (size_t)total * C, stream);
  if (e != cudaSuccess) return (int)e;
  k_splat_nchw<<<splat_grid(total), 256, 0, stream>>>(ten_in, ten_flow, ten_out, N, C, H, W);
  return check_launch();
}

extern "C" int pgdvs_softsplat_workspace_bytes(int B, int H, int W, size_t* bytes) {
  if (!bytes || B < 0 || H <= 0 || W <= 0) return PGDVS_E_BADARG;
  *bytes = sizeof(float) * 8 * (size_t)(B > 0 ? B : 1) * H * W;
  return PGDVS_OK;
}

extern "C" int pgdvs_softsplat_dyn(const float* rgb1, const float* mask1, const float* noise, const float* rgb2,
                                   const float* flow_1_to_tgt, const float* flow_12, float alpha, int B, int H,
                                   int W, float* out_rgb, float* out_mask, float* out_metric, void* workspace,
                                   size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (B < 0 || H <= 0 || W <= 0 || !(alpha >= 0.0f)) return PGDVS_E_BADARG;
  if (B == 0) return PGDVS_OK;
  if (!rgb1 || !mask1 || !rgb2 || !flow_1_to_tgt || !flow_12 || !out_rgb || !out_mask || !workspace)
    return PGDVS_E_BADARG;
  const size_t need = sizeof(float) * 8 * (size_t)B * H * W;
  if (workspace_bytes < need) return PGDVS_E_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(workspace) & 31) != 0) return PGDVS_E_ALIGN;
  SoftsplatDynParams p;
  p.rgb1 = rgb1;
  p.mask1 = mask1;
  p.noise = noise;
  p.rgb2 = rgb2;
  p.flow_t = flow_1_to_tgt;
  p.flow_12 = flow_12;
  p.alpha = alpha;
  p.B = B;
  p.H = H;
  p.W = W;
  p.acc = static_cast<float*>(workspace);
  p.out_rgb = out_rgb;
  p.out_mask = out_mask;
  p.out_metric = out_metric;
  cudaError_t e = cudaMemsetAsync(workspace, 0, need, stream);
  if (e != cudaSuccess) return (int)e;
  const int64_t total = (int64_t)B * H * W;
  k_softsplat_dyn_scatter<<<splat_grid(total), 256, 0, stream>>>(p);
  k_softsplat_dyn_resolve<<<splat_grid(total), 256, 0, stream>>>(p);
  return check_launch();
}
