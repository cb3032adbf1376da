// Mesh mode of the dynamic renderer (`dyn_render_type = mesh`): the grid-topology
// triangles that PGDVSDynamicRenderer.render_dyn_mesh (pgdvs_renderer_dyn.py:542-669) builds from
// the dynamic mask are rasterized with pytorch3d's MeshRasterizer semantics for
//   blur_radius = 0, faces_per_pixel = 1, bin_size = 0 (naive), cull_backfaces = False,
//   clip_barycentric_coords = False, perspective_correct = True
// and shaded like the reference's SimpleShader (pgdvs/utils/pytorch3d_utils.py:50-67): vertex
// colours interpolated with the barycentric coordinates, hard blend on a black background; the
// mask (a second full render with all-ones colours upstream) comes out of the same pass.
//
//   k_mesh_scatter  one thread per face: the face's pixel bounding box is walked and every
//                   covered pixel takes atomicMin(z bits << 32 | face) — nearest face, ties to
//                   the smaller face index, exactly the order of the CPU rasterizer's queue.
//   k_mesh_resolve  one thread per pixel: barycentrics of the winning face recomputed with the
//                   same code (bit-identical), fragments, colour, mask.
// All coverage arithmetic is explicitly rounded (no FMA contraction) so that pix_to_face / zbuf /
// bary are bit-exact against the oracle restatement (oracle/raster_cpu.cpp).
#include "common.cuh"

namespace pgdvs {

constexpr float kMeshEps = 1e-8f;  // pytorch3d csrc/utils/geometry_utils.h kEpsilon

struct MeshParams {
  const float* verts;    // [V,3] (x_ndc, y_ndc, z_view)
  const int32_t* faces;  // [F,3]
  const float* vert_rgb; // [V,3] or null
  int64_t F;
  int H, W, perspective_correct;
  NdcAxis ax, ay;
  float xf0, yf0, inv_pix;
  unsigned long long* keys;  // [H*W]
  int32_t* pix_to_face;
  float* zbuf;
  float* bary;
  float* image;
  float* mask;
};

__device__ __forceinline__ float edge_rn(float px, float py, float ax, float ay, float bx, float by) {
  return __fsub_rn(__fmul_rn(__fsub_rn(px, ax), __fsub_rn(by, ay)), __fmul_rn(__fsub_rn(py, ay), __fsub_rn(bx, ax)));
}

struct FaceEval {
  bool covered;
  float b0, b1, b2, pz;
};

// RasterizeMeshesNaiveCpu's per-(pixel, face) arithmetic for blur_radius = 0
__device__ __forceinline__ FaceEval eval_face(const float v[9], float xf, float yf, int perspective_correct) {
  FaceEval r;
  r.covered = false;
  const float x0 = v[0], y0 = v[1], z0 = v[2], x1 = v[3], y1 = v[4], z1 = v[5], x2 = v[6], y2 = v[7], z2 = v[8];
  const float xmin = fminf(x0, fminf(x1, x2)), xmax = fmaxf(x0, fmaxf(x1, x2));
  const float ymin = fminf(y0, fminf(y1, y2)), ymax = fmaxf(y0, fmaxf(y1, y2));
  const float zmax = fmaxf(z0, fmaxf(z1, z2));
  if (xf < xmin || xf > xmax || yf < ymin || yf > ymax || zmax < kMeshEps) return r;
  const float face_area = edge_rn(x2, y2, x0, y0, x1, y1);
  if (face_area <= kMeshEps && face_area >= -kMeshEps) return r;
  const float area = __fadd_rn(face_area, kMeshEps);
  const float w0 = __fdiv_rn(edge_rn(xf, yf, x1, y1, x2, y2), area);
  const float w1 = __fdiv_rn(edge_rn(xf, yf, x2, y2, x0, y0), area);
  const float w2 = __fdiv_rn(edge_rn(xf, yf, x0, y0, x1, y1), area);
  if (!(w0 > 0.0f && w1 > 0.0f && w2 > 0.0f)) return r;  // blur_radius = 0: inside only
  r.b0 = w0;
  r.b1 = w1;
  r.b2 = w2;
  if (perspective_correct) {
    const float t0 = __fmul_rn(__fmul_rn(w0, z1), z2);
    const float t1 = __fmul_rn(__fmul_rn(z0, w1), z2);
    const float t2 = __fmul_rn(__fmul_rn(z0, z1), w2);
    const float denom = fmaxf(__fadd_rn(__fadd_rn(t0, t1), t2), kMeshEps);
    r.b0 = __fdiv_rn(t0, denom);
    r.b1 = __fdiv_rn(t1, denom);
    r.b2 = __fdiv_rn(t2, denom);
  }
  r.pz = __fadd_rn(__fadd_rn(__fmul_rn(r.b0, z0), __fmul_rn(r.b1, z1)), __fmul_rn(r.b2, z2));
  if (!(r.pz >= 0.0f)) return r;  // pz < 0 (or NaN) -> skipped
  r.covered = true;
  return r;
}

__device__ __forceinline__ void load_face(const MeshParams& p, int64_t f, float v[9]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int32_t vi = __ldg(p.faces + f * 3 + k);
    v[k * 3 + 0] = __ldg(p.verts + (int64_t)vi * 3 + 0);
    v[k * 3 + 1] = __ldg(p.verts + (int64_t)vi * 3 + 1);
    v[k * 3 + 2] = __ldg(p.verts + (int64_t)vi * 3 + 2);
  }
}

__global__ void __launch_bounds__(256) k_mesh_scatter(const MeshParams p) {
  for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < p.F; f += (int64_t)gridDim.x * blockDim.x) {
    float v[9];
    load_face(p, f, v);
    const float xmin = fminf(v[0], fminf(v[3], v[6])), xmax = fmaxf(v[0], fmaxf(v[3], v[6]));
    const float ymin = fminf(v[1], fminf(v[4], v[7])), ymax = fmaxf(v[1], fmaxf(v[4], v[7]));
    if (!(xmin == xmin && xmax == xmax && ymin == ymin && ymax == ymax)) continue;  // NaN vertex
    // pixel centres run from xf0 (column 0) downwards by 1 / inv_pix per column; a conservative
    // column / row range, the exact box test is in eval_face
    const float c_lo = (p.xf0 - xmax) * p.inv_pix - 1.0f, c_hi = (p.xf0 - xmin) * p.inv_pix + 1.0f;
    const float r_lo = (p.yf0 - ymax) * p.inv_pix - 1.0f, r_hi = (p.yf0 - ymin) * p.inv_pix + 1.0f;
    if (!(c_hi >= 0.0f && c_lo <= (float)(p.W - 1) && r_hi >= 0.0f && r_lo <= (float)(p.H - 1))) continue;
    const int x_a = (int)fmaxf(floorf(c_lo), 0.0f), x_b = (int)fminf(ceilf(c_hi), (float)(p.W - 1));
    const int y_a = (int)fmaxf(floorf(r_lo), 0.0f), y_b = (int)fminf(ceilf(r_hi), (float)(p.H - 1));
    for (int y = y_a; y <= y_b; ++y) {
      const float yf = pixel_center_ndc(p.ay, y);
      for (int x = x_a; x <= x_b; ++x) {
        const float xf = pixel_center_ndc(p.ax, x);
        const FaceEval e = eval_face(v, xf, yf, p.perspective_correct);
        if (!e.covered) continue;
        const unsigned long long key =
            ((unsigned long long)__float_as_uint(__fadd_rn(e.pz, 0.0f)) << 32) | (unsigned long long)(unsigned)f;
        atomicMin(p.keys + (int64_t)y * p.W + x, key);
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_mesh_resolve(const MeshParams p) {
  const int64_t HW = (int64_t)p.H * p.W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long key = p.keys[i];
    int f = -1;
    float b[3] = {-1.0f, -1.0f, -1.0f}, pz = -1.0f;
    float col[3] = {0.f, 0.f, 0.f}, ones = 0.f;
    if (key != ~0ull) {
      f = (int)(unsigned)(key & 0xffffffffull);
      const int y = (int)(i / p.W), x = (int)(i % p.W);
      float v[9];
      load_face(p, f, v);
      const FaceEval e = eval_face(v, pixel_center_ndc(p.ax, x), pixel_center_ndc(p.ay, y), p.perspective_correct);
      b[0] = e.b0;
      b[1] = e.b1;
      b[2] = e.b2;
      pz = e.pz;
      // interpolate_face_attributes: sum_i bary_i * attribute_i
      ones = __fadd_rn(__fadd_rn(b[0], b[1]), b[2]);
      if (p.vert_rgb) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int32_t vi = __ldg(p.faces + (int64_t)f * 3 + k);
            acc = (k == 0) ? __fmul_rn(b[0], __ldg(p.vert_rgb + (int64_t)vi * 3 + c))
                           : __fadd_rn(acc, __fmul_rn(b[k], __ldg(p.vert_rgb + (int64_t)vi * 3 + c)));
          }
          col[c] = acc;
        }
      }
    }
    if (p.pix_to_face) p.pix_to_face[i] = f;
    if (p.zbuf) p.zbuf[i] = pz;
    if (p.bary) {
      p.bary[i * 3 + 0] = b[0];
      p.bary[i * 3 + 1] = b[1];
      p.bary[i * 3 + 2] = b[2];
    }
    if (p.image) {
      p.image[i * 3 + 0] = col[0];
      p.image[i * 3 + 1] = col[1];
      p.image[i * 3 + 2] = col[2];
    }
    if (p.mask) p.mask[i] = (ones > 0.0f) ? 1.0f : 0.0f;
  }
}

}  // namespace pgdvs

using namespace pgdvs;

extern "C" int pgdvs_mesh_workspace_bytes(int H, int W, size_t* bytes) {
  if (!bytes || H <= 0 || W <= 0) return PGDVS_E_BADARG;
  *bytes = sizeof(unsigned long long) * (size_t)H * W;
  return PGDVS_OK;
}

extern "C" int pgdvs_rasterize_mesh(const float* verts_ndc, int64_t V, const int32_t* faces, int64_t F, int H,
                                    int W, int perspective_correct, const float* vert_rgb,
                                    int32_t* pix_to_face, float* zbuf, float* bary, float* image, float* mask,
                                    void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (V < 0 || F < 0 || H <= 0 || W <= 0 || !workspace) return PGDVS_E_BADARG;
  if (F >= (int64_t)INT32_MAX) return PGDVS_E_BADARG;
  if (F > 0 && (!verts_ndc || !faces)) return PGDVS_E_BADARG;
  if (image != nullptr && vert_rgb == nullptr) return PGDVS_E_BADARG;
  const size_t need = sizeof(unsigned long long) * (size_t)H * W;
  if (workspace_bytes < need) return PGDVS_E_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(workspace) & 7) != 0) return PGDVS_E_ALIGN;
  MeshParams p;
  p.verts = verts_ndc;
  p.faces = faces;
  p.vert_rgb = vert_rgb;
  p.F = F;
  p.H = H;
  p.W = W;
  p.perspective_correct = perspective_correct ? 1 : 0;
  p.ax = make_ndc_axis(W, H);
  p.ay = make_ndc_axis(H, W);
  const CellGrid g = make_cell_grid(H, W, 0);
  p.xf0 = g.xf0;
  p.yf0 = g.yf0;
  p.inv_pix = g.inv_pix;
  p.keys = static_cast<unsigned long long*>(workspace);
  p.pix_to_face = pix_to_face;
  p.zbuf = zbuf;
  p.bary = bary;
  p.image = image;
  p.mask = mask;
  cudaError_t e = cudaMemsetAsync(workspace, 0xFF, need, stream);
  if (e != cudaSuccess) return (int)e;
  if (F > 0) {
    int64_t gb = (F + 255) / 256;
    if (gb > 148 * 32) gb = 148 * 32;
    k_mesh_scatter<<<(unsigned)gb, 256, 0, stream>>>(p);
  }
  int64_t gr = ((int64_t)H * W + 255) / 256;
  if (gr > 148 * 32) gr = 148 * 32;
  k_mesh_resolve<<<(unsigned)gr, 256, 0, stream>>>(p);
  return check_launch();
}
