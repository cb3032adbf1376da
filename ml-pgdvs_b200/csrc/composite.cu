// Stand-alone compositors (pytorch3d layout), dyn/track merge + static blend, and the
// world -> NDC projection used when a caller hands over an already-built cloud.
#include "common.cuh"
#include <cstddef>

namespace pgdvs {

// pytorch3d csrc/compositing/{alpha_composite,norm_weighted_sum,weighted_sum}.cu forward:
// one thread per (n, c, pixel); idx i64 [N,K,H,W], alphas [N,K,H,W], features [C,P].
// The arithmetic order restates the CPU loops (oracle/raster_cpu.cpp) with explicitly
// rounded ops so the result is bit-identical to the oracle.
__global__ void __launch_bounds__(256) k_composite(const int64_t* __restrict__ idx,
                                                   const float* __restrict__ alphas,
                                                   const float* __restrict__ features, int N, int K,
                                                   int64_t HW, int C, int64_t P, int mode,
                                                   float* __restrict__ out) {
  const int64_t total = (int64_t)N * C * HW;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t px = t % HW;
    const int c = (int)((t / HW) % C);
    const int n = (int)(t / (HW * C));
    const int64_t base = (int64_t)n * K * HW + px;
    const float* f = features + (int64_t)c * P;
    float res = 0.0f;
    if (mode == PGDVS_COMPOSITE_ALPHA) {
      float cum = 1.0f;
      for (int k = 0; k < K; ++k) {
        const int64_t l = __ldg(idx + base + k * HW);
        if (l < 0) continue;
        const float a = __ldg(alphas + base + k * HW);
        res = __fadd_rn(res, __fmul_rn(__fmul_rn(cum, a), __ldg(f + l)));
        cum = __fmul_rn(cum, __fsub_rn(1.0f, a));
      }
    } else if (mode == PGDVS_COMPOSITE_NORM_WEIGHTED) {
      float t_alpha = 0.0f;
      for (int k = 0; k < K; ++k) {
        const int64_t l = __ldg(idx + base + k * HW);
        if (l < 0) continue;
        t_alpha = __fadd_rn(t_alpha, __ldg(alphas + base + k * HW));
      }
      t_alpha = fmaxf(t_alpha, 1e-4f);
      for (int k = 0; k < K; ++k) {
        const int64_t l = __ldg(idx + base + k * HW);
        if (l < 0) continue;
        const float a = __ldg(alphas + base + k * HW);
        res = __fadd_rn(res, __fdiv_rn(__fmul_rn(a, __ldg(f + l)), t_alpha));
      }
    } else {
      for (int k = 0; k < K; ++k) {
        const int64_t l = __ldg(idx + base + k * HW);
        if (l < 0) continue;
        const float a = __ldg(alphas + base + k * HW);
        res = __fadd_rn(res, __fmul_rn(a, __ldg(f + l)));
      }
    }
    out[t] = res;
  }
}

// pgdvs_renderer_dyn.py:229-235 and pgdvs_renderer.py:169-172, channels-first [B,3,H,W].
__global__ void __launch_bounds__(256) k_merge_blend(const float* __restrict__ dyn_rgb,
                                                     const float* __restrict__ dyn_mask,
                                                     const float* __restrict__ track_rgb,
                                                     const float* __restrict__ track_mask,
                                                     const float* __restrict__ static_rgb, int B,
                                                     int64_t HW, float* __restrict__ out_rgb,
                                                     float* __restrict__ out_mask,
                                                     float* __restrict__ out_combined) {
  const int64_t total = (int64_t)B * HW;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / HW, px = t % HW;
    const float dm = __ldg(dyn_mask + t);
    const float tm = track_mask ? __ldg(track_mask + t) : 0.0f;
    const float m_track = (!(dm > 0.0f) && (tm > 0.0f)) ? 1.0f : 0.0f;
    const float m = ((dm > 0.0f) || (tm > 0.0f)) ? 1.0f : 0.0f;
    if (out_mask) out_mask[t] = m;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int64_t o = (b * 3 + c) * HW + px;
      const float d = __ldg(dyn_rgb + o);
      const float tr = track_rgb ? __ldg(track_rgb + o) : 0.0f;
      const float rgb = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, m_track), d), __fmul_rn(m_track, tr));
      if (out_rgb) out_rgb[o] = rgb;
      if (out_combined) {
        const float st = __ldg(static_rgb + o);
        out_combined[o] = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, m), st), __fmul_rn(m, rgb));
      }
    }
  }
}

// 8-bit frames as the reference's evaluator / video writer consume them
// (engines/evaluator_pgdvs.py:51-77): NaN -> 0, clamp to [0,1], (x * 255).byte() (truncation).
__global__ void __launch_bounds__(256) k_quantize_u8(const float4* __restrict__ in, uchar4* __restrict__ out,
                                                     int64_t n4, const float* __restrict__ tail_in,
                                                     uint8_t* __restrict__ tail_out, int n_tail) {
  auto q = [](float v) -> unsigned char {
    v = (v != v) ? 0.0f : fminf(fmaxf(v, 0.0f), 1.0f);
    return (unsigned char)(int)(v * 255.0f);
  };
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(in + i);
    out[i] = make_uchar4(q(v.x), q(v.y), q(v.z), q(v.w));
  }
  if (blockIdx.x == 0 && threadIdx.x < n_tail) tail_out[threadIdx.x] = q(tail_in[threadIdx.x]);
}

__global__ void __launch_bounds__(256) k_project(const float* __restrict__ xyz, int64_t P,
                                                 const PgdvsCamera* __restrict__ cam,
                                                 float* __restrict__ out) {
  const PgdvsCamera c = *cam;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float3 o = world_to_ndc(c, __ldg(xyz + i * 3), __ldg(xyz + i * 3 + 1), __ldg(xyz + i * 3 + 2));
    out[i * 3 + 0] = o.x;
    out[i * 3 + 1] = o.y;
    out[i * 3 + 2] = o.z;
  }
}

// Projector.compute_projections (models/gnt/projector.py:41-73) for n_cams cameras
__global__ void __launch_bounds__(256) k_compute_projections(const float* __restrict__ xyz, int64_t P,
                                                             const float* __restrict__ proj, int n_cams,
                                                             float* __restrict__ uv, uint8_t* __restrict__ mask) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float x = __ldg(xyz + i * 3), y = __ldg(xyz + i * 3 + 1), z = __ldg(xyz + i * 3 + 2);
    for (int c = 0; c < n_cams; ++c) {
      const float* m = proj + c * 12;
      const float hx = m[0] * x + m[1] * y + m[2] * z + m[3];
      const float hy = m[4] * x + m[5] * y + m[6] * z + m[7];
      const float hz = m[8] * x + m[9] * y + m[10] * z + m[11];
      const float d = fmaxf(hz, 1e-8f);  // torch.clamp(min=1e-8)
      const float u = fminf(fmaxf(__fdiv_rn(hx, d), -1e6f), 1e6f);
      const float v = fminf(fmaxf(__fdiv_rn(hy, d), -1e6f), 1e6f);
      reinterpret_cast<float2*>(uv)[(int64_t)c * P + i] = make_float2(u, v);
      if (mask) mask[(int64_t)c * P + i] = hz > 0.0f ? 1 : 0;
    }
  }
}

static inline int grid_for(int64_t total) {
  int64_t g = (total + 255) / 256;
  if (g < 1) g = 1;
  if (g > 148 * 32) g = 148 * 32;
  return (int)g;
}

}  // namespace pgdvs

using namespace pgdvs;

extern "C" int pgdvs_composite(const int64_t* idx, const float* alphas, const float* features,
                               int N, int K, int H, int W, int C, int64_t P, int mode, float* out,
                               void* stream) {
  if (N < 0 || K < 1 || H <= 0 || W <= 0 || C < 1 || P < 0) return PGDVS_E_BADARG;
  if (mode < PGDVS_COMPOSITE_ALPHA || mode > PGDVS_COMPOSITE_WEIGHTED_SUM) return PGDVS_E_BADARG;
  if (N == 0) return PGDVS_OK;
  if (!idx || !alphas || !out || (P > 0 && !features)) return PGDVS_E_BADARG;
  const int64_t HW = (int64_t)H * W;
  k_composite<<<grid_for((int64_t)N * C * HW), 256, 0, (cudaStream_t)stream>>>(
      idx, alphas, features, N, K, HW, C, P, mode, out);
  return check_launch();
}

extern "C" int pgdvs_merge_blend(const float* dyn_rgb, const float* dyn_mask, const float* track_rgb,
                                 const float* track_mask, const float* static_rgb, int B, int H,
                                 int W, float* out_rgb, float* out_mask, float* out_combined,
                                 void* stream) {
  if (B < 0 || H <= 0 || W <= 0 || !dyn_rgb || !dyn_mask) return PGDVS_E_BADARG;
  if ((track_rgb == nullptr) != (track_mask == nullptr)) return PGDVS_E_BADARG;
  if (out_combined != nullptr && static_rgb == nullptr) return PGDVS_E_BADARG;
  if (B == 0) return PGDVS_OK;
  const int64_t HW = (int64_t)H * W;
  k_merge_blend<<<grid_for((int64_t)B * HW), 256, 0, (cudaStream_t)stream>>>(
      dyn_rgb, dyn_mask, track_rgb, track_mask, static_rgb, B, HW, out_rgb, out_mask, out_combined);
  return check_launch();
}

extern "C" int pgdvs_quantize_u8(const float* in, uint8_t* out, int64_t n, void* stream) {
  if (n < 0) return PGDVS_E_BADARG;
  if (n == 0) return PGDVS_OK;
  if (!in || !out) return PGDVS_E_BADARG;
  if ((reinterpret_cast<uintptr_t>(in) & 15) != 0 || (reinterpret_cast<uintptr_t>(out) & 3) != 0) return PGDVS_E_ALIGN;
  const int64_t n4 = n / 4;
  const int n_tail = (int)(n - n4 * 4);
  k_quantize_u8<<<grid_for(n4 > 0 ? n4 : 1), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(in), reinterpret_cast<uchar4*>(out), n4, in + n4 * 4, out + n4 * 4, n_tail);
  return check_launch();
}

extern "C" int pgdvs_project_points(const float* xyz_world, int64_t P, const PgdvsCamera* camera_dev,
                                    float* xyz_ndc, void* stream) {
  if (P < 0 || !camera_dev) return PGDVS_E_BADARG;
  if (P == 0) return PGDVS_OK;
  if (!xyz_world || !xyz_ndc) return PGDVS_E_BADARG;
  k_project<<<grid_for(P), 256, 0, (cudaStream_t)stream>>>(xyz_world, P, camera_dev, xyz_ndc);
  return check_launch();
}

extern "C" int pgdvs_compute_projections(const float* xyz, int64_t P, const float* proj, int n_cams,
                                         float* uv, uint8_t* mask, void* stream) {
  if (P < 0 || n_cams < 0) return PGDVS_E_BADARG;
  if (P == 0 || n_cams == 0) return PGDVS_OK;
  if (!xyz || !proj || !uv) return PGDVS_E_BADARG;
  if ((reinterpret_cast<uintptr_t>(uv) & 7) != 0) return PGDVS_E_ALIGN;
  k_compute_projections<<<grid_for(P), 256, 0, (cudaStream_t)stream>>>(xyz, P, proj, n_cams, uv, mask);
  return check_launch();
}

extern "C" int pgdvs_abi_version(void) { return PGDVS_B200_ABI_VERSION; }

extern "C" int pgdvs_struct_layout(int32_t out[4]) {
  if (!out) return PGDVS_E_BADARG;
  out[0] = (int32_t)sizeof(PgdvsCamera);
  out[1] = (int32_t)sizeof(PgdvsUwpJob);
  out[2] = (int32_t)offsetof(PgdvsUwpJob, M1);
  out[3] = (int32_t)offsetof(PgdvsUwpJob, view);
  return PGDVS_OK;
}

extern "C" const char* pgdvs_error_string(int code) {
  switch (code) {
    case PGDVS_OK: return "ok";
    case PGDVS_E_BADARG: return "bad argument (null pointer, non-positive size or unknown mode)";
    case PGDVS_E_K_TOO_LARGE: return "points_per_pixel exceeds kMaxPointsPerPixel (150)";
    case PGDVS_E_WORKSPACE: return "workspace smaller than *_workspace_bytes()";
    case PGDVS_E_CHANNELS: return "fused compositing supports at most 4 feature channels";
    case PGDVS_E_ALIGN: return "pointer is not aligned as documented";
    case PGDVS_E_KNN_K: return "knn K exceeds 64 (dyn_pcl_outlier_knn + 1 must be <= 64)";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown error";
}
