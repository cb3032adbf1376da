// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of the arithmetic the PGDVS dynamic renderer delegates to
// pytorch3d 0.7.4 (pinned in /root/reference/README.md:38, NOT vendored, NOT
// installed here).  Call sites in the reference:
//   pgdvs/renderers/pgdvs_renderer_dyn.py:684-722   (render_dyn_pcl)
//   pgdvs/renderers/st_geo_renderer.py:85-120
// with bin_size=0, i.e. pytorch3d's *naive* rasterizer, and
// NormWeightedCompositor (AlphaCompositor is the commented-out alternative,
// pgdvs_renderer_dyn.py:704-709).
//
// PARITY UNPINNED for this file: the reference ships no tests / golden vectors
// for this path and pytorch3d's source is absent, so the functions below restate
// the *published* pytorch3d 0.7.4 algorithm:
//   csrc/rasterize_points/rasterize_points_cpu.cpp  RasterizePointsNaiveCpu
//   csrc/utils/.../rasterization_utils.h            NonSquareNdcRange, PixToNonSquareNdc
//   csrc/compositing/{alpha_composite,norm_weighted_sum,weighted_sum}_cpu.cpp
// They are anchored on the reference's call sites above and on hand-derived
// known-answer tests (tests/test_oracle_kat.py).
//
// Build: g++ -O2 -ffp-contract=off -fno-fast-math  (no FMA contraction: the x86
// build of pytorch3d evaluates dx*dx + dy*dy with two roundings; the CUDA kernel
// reproduces that with __fmul_rn/__fadd_rn).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <tuple>
#include <vector>

namespace {

// pytorch3d rasterization_utils.h: NDC range of the longer image side is scaled so
// that the shorter side spans [-1, 1].
inline float NonSquareNdcRange(int S1, int S2) {
  float range = 2.0f;
  if (S1 > S2) {
    range = (S1 * range) / S2;
  }
  return range;
}

// NDC coordinate of the centre of pixel i along an axis of S1 pixels (S2 = other axis).
inline float PixToNonSquareNdc(int i, int S1, int S2) {
  float range = NonSquareNdcRange(S1, S2);
  const float offset = (range / 2.0f);
  return -offset + (range * i + offset) / S1;
}

// (z, idx, dist2) ordered lexicographically, exactly like
// std::priority_queue<std::tuple<float,int,float>> in RasterizePointsNaiveCpu.
struct Hit {
  float z;
  int idx;
  float d2;
};
inline bool hit_less(const Hit& a, const Hit& b) {
  return std::tie(a.z, a.idx, a.d2) < std::tie(b.z, b.idx, b.d2);
}

// A bounded max-heap of at most K hits: push, and if size > K pop the largest.
// Final order = ascending (z, idx, d2).  Equivalent to the priority_queue use in
// RasterizePointsNaiveCpu (emplace; if size > K pop; then pop all back to front).
struct KHeap {
  std::vector<Hit> h;
  int K;
  explicit KHeap(int k) : K(k) { h.reserve(k + 1); }
  void clear() { h.clear(); }
  void push(const Hit& x) {
    h.push_back(x);
    std::push_heap(h.begin(), h.end(), hit_less);
    if ((int)h.size() > K) {
      std::pop_heap(h.begin(), h.end(), hit_less);
      h.pop_back();
    }
  }
  // Writes ascending order into out arrays (slots beyond size keep their fill).
  void drain(int32_t* idx, float* z, float* d2) {
    while (!h.empty()) {
      std::pop_heap(h.begin(), h.end(), hit_less);
      Hit t = h.back();
      h.pop_back();
      int i = (int)h.size();
      z[i] = t.z;
      idx[i] = t.idx;
      d2[i] = t.d2;
    }
  }
};

template <typename F>
void parallel_rows(int64_t n_rows, int n_threads, F&& fn) {
  if (n_threads <= 1 || n_rows <= 1) {
    for (int64_t r = 0; r < n_rows; ++r) fn(r);
    return;
  }
  std::atomic<int64_t> next(0);
  std::vector<std::thread> pool;
  n_threads = (int)std::min<int64_t>(n_threads, n_rows);
  for (int t = 0; t < n_threads; ++t) {
    pool.emplace_back([&]() {
      for (;;) {
        int64_t r = next.fetch_add(1);
        if (r >= n_rows) break;
        fn(r);
      }
    });
  }
  for (auto& th : pool) th.join();
}

}  // namespace

extern "C" {

// RasterizePointsNaiveCpu restatement.
//  points  [P,3] NDC (x,y) + view-space z;  first_idx/num_pts [N] int64
//  radius  [P] fp32 (pytorch3d broadcasts a float radius to a [P] tensor)
//  outputs [N,H,W,K], pre-filled here with -1 (int32 / fp32 / fp32)
// n_threads > 1 distributes image rows over threads; pixels are independent so the
// result is bitwise identical to the single-threaded (upstream) order.
int oracle_rasterize_points_naive(const float* points, const int64_t* first_idx,
                                  const int64_t* num_pts, int N, int H, int W,
                                  const float* radius, int K, int32_t* idx_out,
                                  float* zbuf_out, float* dists_out, int n_threads) {
  if (K <= 0 || H <= 0 || W <= 0 || N < 0) return 1;
  const int64_t total = (int64_t)N * H * W * K;
  for (int64_t i = 0; i < total; ++i) {
    idx_out[i] = -1;
    zbuf_out[i] = -1.0f;
    dists_out[i] = -1.0f;
  }
  parallel_rows((int64_t)N * H, n_threads, [&](int64_t row) {
    const int n = (int)(row / H);
    const int yi = (int)(row % H);
    const int64_t p0 = first_idx[n];
    const int64_t p1 = p0 + num_pts[n];
    // Reverse the order of yi so that +Y points up in the image.
    const int yidx = H - 1 - yi;
    const float yf = PixToNonSquareNdc(yidx, H, W);
    KHeap q(K);
    for (int xi = 0; xi < W; ++xi) {
      // Reverse the order of xi so that +X points left in the image.
      const int xidx = W - 1 - xi;
      const float xf = PixToNonSquareNdc(xidx, W, H);
      q.clear();
      for (int64_t p = p0; p < p1; ++p) {
        const float px = points[p * 3 + 0];
        const float py = points[p * 3 + 1];
        const float pz = points[p * 3 + 2];
        const float pr = radius[p];
        const float radius2 = pr * pr;
        if (pz < 0) continue;  // behind the camera
        const float dx = px - xf;
        const float dy = py - yf;
        const float dist2 = dx * dx + dy * dy;
        if (dist2 < radius2) {
          q.push(Hit{pz, (int)p, dist2});
        }
      }
      const int64_t o = (((int64_t)n * H + yi) * W + xi) * K;
      q.drain(idx_out + o, zbuf_out + o, dists_out + o);
    }
  });
  return 0;
}

// Same loop nest restricted to image rows [y0, y1) (outputs are [N, y1-y0, W, K]).  Used by
// bench.py to time a BOUNDED sample of the naive CPU rasterizer: its cost is exactly
// proportional to the number of rows because every pixel visits every point.
int oracle_rasterize_points_naive_rows(const float* points, const int64_t* first_idx,
                                       const int64_t* num_pts, int N, int H, int W,
                                       const float* radius, int K, int y0, int y1,
                                       int32_t* idx_out, float* zbuf_out, float* dists_out,
                                       int n_threads) {
  if (K <= 0 || H <= 0 || W <= 0 || N < 0 || y0 < 0 || y1 > H || y0 >= y1) return 1;
  const int R = y1 - y0;
  const int64_t total = (int64_t)N * R * W * K;
  for (int64_t i = 0; i < total; ++i) {
    idx_out[i] = -1;
    zbuf_out[i] = -1.0f;
    dists_out[i] = -1.0f;
  }
  parallel_rows((int64_t)N * R, n_threads, [&](int64_t row) {
    const int n = (int)(row / R);
    const int yi = y0 + (int)(row % R);
    const int64_t p0 = first_idx[n];
    const int64_t p1 = p0 + num_pts[n];
    const float yf = PixToNonSquareNdc(H - 1 - yi, H, W);
    KHeap q(K);
    for (int xi = 0; xi < W; ++xi) {
      const float xf = PixToNonSquareNdc(W - 1 - xi, W, H);
      q.clear();
      for (int64_t p = p0; p < p1; ++p) {
        const float px = points[p * 3 + 0];
        const float py = points[p * 3 + 1];
        const float pz = points[p * 3 + 2];
        const float pr = radius[p];
        const float radius2 = pr * pr;
        if (pz < 0) continue;
        const float dx = px - xf;
        const float dy = py - yf;
        const float dist2 = dx * dx + dy * dy;
        if (dist2 < radius2) q.push(Hit{pz, (int)p, dist2});
      }
      const int64_t o = (((int64_t)n * R + (yi - y0)) * W + xi) * K;
      q.drain(idx_out + o, zbuf_out + o, dists_out + o);
    }
  });
  return 0;
}

// Accelerated checker: identical per-(pixel,point) arithmetic and identical
// (z, idx, d2) ordering as oracle_rasterize_points_naive, but each pixel row only
// visits points whose y lies within a conservative band of the row (points are
// bucketed by row first).  Validated bitwise against the naive restatement in
// tests/test_oracle_kat.py; exists so that parity can be checked at
// BASELINE.json's full image sizes in seconds.
int oracle_rasterize_points_banded(const float* points, const int64_t* first_idx,
                                   const int64_t* num_pts, int N, int H, int W,
                                   const float* radius, int K, int32_t* idx_out,
                                   float* zbuf_out, float* dists_out, int n_threads) {
  if (K <= 0 || H <= 0 || W <= 0 || N < 0) return 1;
  const int64_t total = (int64_t)N * H * W * K;
  for (int64_t i = 0; i < total; ++i) {
    idx_out[i] = -1;
    zbuf_out[i] = -1.0f;
    dists_out[i] = -1.0f;
  }
  const float pix = NonSquareNdcRange(H, W) / (float)H;  // NDC size of one pixel
  for (int n = 0; n < N; ++n) {
    const int64_t p0 = first_idx[n];
    const int64_t p1 = p0 + num_pts[n];
    // bucket[r] = points that may touch row r (conservative: +1 row of slack each side,
    // plus a second pass over x inside the row with +-1 column of slack).
    std::vector<std::vector<int64_t>> bucket(H);
    const float y_top = PixToNonSquareNdc(H - 1, H, W);  // yf of yi = 0
    for (int64_t p = p0; p < p1; ++p) {
      const float py = points[p * 3 + 1];
      const float pz = points[p * 3 + 2];
      const float pr = radius[p];
      if (!(pz >= 0) && !(pz != pz)) continue;  // pz < 0 (NaN falls through like upstream)
      if (!(py == py) || !(pr == pr)) continue;  // NaN y / radius can never satisfy dist2 < r2
      // yi such that yf(yi) ~= py:  yf(yi) = y_top - yi * pix
      const double c = ((double)y_top - (double)py) / (double)pix;
      const double rp = std::fabs((double)pr) / (double)pix + 1.5;
      int64_t lo = (int64_t)std::floor(c - rp);
      int64_t hi = (int64_t)std::ceil(c + rp);
      if (hi < 0 || lo > H - 1) continue;
      lo = std::max<int64_t>(lo, 0);
      hi = std::min<int64_t>(hi, H - 1);
      for (int64_t r = lo; r <= hi; ++r) bucket[r].push_back(p);
    }
    const float x_left = PixToNonSquareNdc(W - 1, W, H);  // xf of xi = 0
    const float pixx = NonSquareNdcRange(W, H) / (float)W;
    parallel_rows(H, n_threads, [&](int64_t yi) {
      const int yidx = H - 1 - (int)yi;
      const float yf = PixToNonSquareNdc(yidx, H, W);
      const std::vector<int64_t>& b = bucket[yi];
      // column buckets for this row
      std::vector<std::vector<int64_t>> col(W);
      for (int64_t p : b) {
        const float px = points[p * 3 + 0];
        const float pr = radius[p];
        if (!(px == px)) continue;
        const double c = ((double)x_left - (double)px) / (double)pixx;
        const double rp = std::fabs((double)pr) / (double)pixx + 1.5;
        int64_t lo = (int64_t)std::floor(c - rp);
        int64_t hi = (int64_t)std::ceil(c + rp);
        if (hi < 0 || lo > W - 1) continue;
        lo = std::max<int64_t>(lo, 0);
        hi = std::min<int64_t>(hi, W - 1);
        for (int64_t x = lo; x <= hi; ++x) col[x].push_back(p);
      }
      KHeap q(K);
      for (int xi = 0; xi < W; ++xi) {
        const int xidx = W - 1 - xi;
        const float xf = PixToNonSquareNdc(xidx, W, H);
        q.clear();
        for (int64_t p : col[xi]) {  // ascending p, like the naive loop
          const float px = points[p * 3 + 0];
          const float py = points[p * 3 + 1];
          const float pz = points[p * 3 + 2];
          const float pr = radius[p];
          const float radius2 = pr * pr;
          if (pz < 0) continue;
          const float dx = px - xf;
          const float dy = py - yf;
          const float dist2 = dx * dx + dy * dy;
          if (dist2 < radius2) q.push(Hit{pz, (int)p, dist2});
        }
        const int64_t o = (((int64_t)n * H + yi) * W + xi) * K;
        q.drain(idx_out + o, zbuf_out + o, dists_out + o);
      }
    });
  }
  return 0;
}

// NDC pixel-centre helper exported for tests (xf of column xi / yf of row yi).
float oracle_pixel_center_x(int xi, int H, int W) { return PixToNonSquareNdc(W - 1 - xi, W, H); }
float oracle_pixel_center_y(int yi, int H, int W) { return PixToNonSquareNdc(H - 1 - yi, H, W); }

// ---- compositors: idx int64 [N,K,H,W], alphas f32 [N,K,H,W], features f32 [C,P] -> [N,C,H,W]
// mode 1 = alpha_composite, 2 = norm_weighted_sum, 3 = weighted_sum
// (pytorch3d csrc/compositing/*_cpu.cpp forward loops; kEps = 1e-4 for the norm).
int oracle_composite(const int64_t* idx, const float* alphas, const float* features, int N,
                     int K, int H, int W, int C, int64_t P, int mode, float* out) {
  (void)P;
  const float kEps = 1e-4f;
  const int64_t HW = (int64_t)H * W;
  for (int n = 0; n < N; ++n) {
    for (int c = 0; c < C; ++c) {
      for (int64_t px = 0; px < HW; ++px) {
        float res = 0.0f;
        const int64_t base = (int64_t)n * K * HW + px;
        if (mode == 1) {
          float cum_alpha = 1.0f;
          for (int k = 0; k < K; ++k) {
            const int64_t l = idx[base + k * HW];
            if (l < 0) continue;
            const float alpha = alphas[base + k * HW];
            res += cum_alpha * alpha * features[(int64_t)c * P + l];
            cum_alpha = cum_alpha * (1 - alpha);
          }
        } else if (mode == 2) {
          float t_alpha = 0.0f;
          for (int k = 0; k < K; ++k) {
            const int64_t l = idx[base + k * HW];
            if (l < 0) continue;
            t_alpha += alphas[base + k * HW];
          }
          t_alpha = std::max(t_alpha, kEps);
          for (int k = 0; k < K; ++k) {
            const int64_t l = idx[base + k * HW];
            if (l < 0) continue;
            const float alpha = alphas[base + k * HW];
            res += alpha * features[(int64_t)c * P + l] / t_alpha;
          }
        } else if (mode == 3) {
          for (int k = 0; k < K; ++k) {
            const int64_t l = idx[base + k * HW];
            if (l < 0) continue;
            const float alpha = alphas[base + k * HW];
            res += alpha * features[(int64_t)c * P + l];
          }
        } else {
          return 2;
        }
        out[((int64_t)n * C + c) * HW + px] = res;
      }
    }
  }
  return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// Mesh rasterization, faces_per_pixel = 1 semantics generalised to K (PARITY UNPINNED:
// restated from the published pytorch3d 0.7.4 algorithm, csrc/rasterize_meshes/
// rasterize_meshes_cpu.cpp `RasterizeMeshesNaiveCpu` and csrc/utils/geometry_utils.h).
// Reached from PGDVSDynamicRenderer.render_dyn_mesh (pgdvs_renderer_dyn.py:542-669) through
// MeshRasterizer with blur_radius = 0, bin_size = 0, cull_backfaces = False,
// clip_barycentric_coords = False, perspective_correct = True (PerspectiveCameras).
//   face_verts f32 [F,3,3] (x_ndc, y_ndc, z_view); one mesh.
//   outputs -1 filled: pix_to_face i32 [H,W,K], zbuf f32 [H,W,K], bary f32 [H,W,K,3],
//   dists f32 [H,W,K] (signed squared distance to the nearest edge; negative inside).
// ---------------------------------------------------------------------------------------
namespace {
constexpr float kMeshEps = 1e-8f;

inline float EdgeFunction(float px, float py, float ax, float ay, float bx, float by) {
  return (px - ax) * (by - ay) - (py - ay) * (bx - ax);
}

inline float PointLineDist2(float px, float py, float ax, float ay, float bx, float by) {
  // squared distance from p to segment ab (geometry_utils.h PointLineDistanceForward)
  const float abx = bx - ax, aby = by - ay;
  const float l2 = abx * abx + aby * aby;
  if (l2 <= kMeshEps) return (px - bx) * (px - bx) + (py - by) * (py - by);
  float t = (abx * (px - ax) + aby * (py - ay)) / l2;
  t = std::min(std::max(t, 0.0f), 1.0f);
  const float qx = ax + t * abx, qy = ay + t * aby;
  return (px - qx) * (px - qx) + (py - qy) * (py - qy);
}

struct FaceHit {
  float z;
  int f;
  float dist, b0, b1, b2;
};
inline bool face_less(const FaceHit& a, const FaceHit& b) {
  return std::make_tuple(a.z, a.f, a.dist, a.b0, a.b1, a.b2) < std::make_tuple(b.z, b.f, b.dist, b.b0, b.b1, b.b2);
}
}  // namespace

extern "C" int oracle_rasterize_meshes_naive(const float* face_verts, int64_t F, int H, int W, int K,
                                             float blur_radius, int perspective_correct, int32_t* pix_to_face,
                                             float* zbuf, float* bary, float* dists) {
  if (H <= 0 || W <= 0 || K < 1 || F < 0) return 1;
  const float blur = std::sqrt(blur_radius);
  for (int yi = 0; yi < H; ++yi) {
    const float yf = PixToNonSquareNdc(H - 1 - yi, H, W);
    for (int xi = 0; xi < W; ++xi) {
      const float xf = PixToNonSquareNdc(W - 1 - xi, W, H);
      std::vector<FaceHit> q;  // kept sorted ascending, at most K entries (== the priority queue)
      for (int64_t f = 0; f < F; ++f) {
        const float* v = face_verts + f * 9;
        const float x0 = v[0], y0 = v[1], z0 = v[2], x1 = v[3], y1 = v[4], z1 = v[5], x2 = v[6], y2 = v[7], z2 = v[8];
        const float xmin = std::min(x0, std::min(x1, x2)), xmax = std::max(x0, std::max(x1, x2));
        const float ymin = std::min(y0, std::min(y1, y2)), ymax = std::max(y0, std::max(y1, y2));
        const float zmax = std::max(z0, std::max(z1, z2));
        // CheckPointOutsideBoundingBox (+ faces entirely behind the camera)
        if (xf < xmin - blur || xf > xmax + blur || yf < ymin - blur || yf > ymax + blur || zmax < kMeshEps) continue;
        const float face_area = EdgeFunction(x2, y2, x0, y0, x1, y1);
        if (face_area <= kMeshEps && face_area >= -kMeshEps) continue;  // degenerate
        const float area = face_area + kMeshEps;
        const float w0 = EdgeFunction(xf, yf, x1, y1, x2, y2) / area;
        const float w1 = EdgeFunction(xf, yf, x2, y2, x0, y0) / area;
        const float w2 = EdgeFunction(xf, yf, x0, y0, x1, y1) / area;
        float b0 = w0, b1 = w1, b2 = w2;
        if (perspective_correct) {
          const float t0 = w0 * z1 * z2, t1 = z0 * w1 * z2, t2 = z0 * z1 * w2;
          const float denom = std::max(t0 + t1 + t2, kMeshEps);
          b0 = t0 / denom;
          b1 = t1 / denom;
          b2 = t2 / denom;
        }
        const float pz = b0 * z0 + b1 * z1 + b2 * z2;
        if (pz < 0) continue;
        const float e0 = PointLineDist2(xf, yf, x0, y0, x1, y1);
        const float e1 = PointLineDist2(xf, yf, x0, y0, x2, y2);
        const float e2 = PointLineDist2(xf, yf, x1, y1, x2, y2);
        const float dist = std::min(e0, std::min(e1, e2));
        const bool inside = w0 > 0.0f && w1 > 0.0f && w2 > 0.0f;
        if (!inside && dist >= blur_radius) continue;
        FaceHit h{pz, (int)f, inside ? -dist : dist, b0, b1, b2};
        auto it = std::upper_bound(q.begin(), q.end(), h, face_less);
        q.insert(it, h);
        if ((int)q.size() > K) q.pop_back();
      }
      const int64_t base = ((int64_t)yi * W + xi) * K;
      for (int k = 0; k < K; ++k) {
        const bool has = k < (int)q.size();
        pix_to_face[base + k] = has ? q[k].f : -1;
        zbuf[base + k] = has ? q[k].z : -1.0f;
        dists[base + k] = has ? q[k].dist : -1.0f;
        bary[(base + k) * 3 + 0] = has ? q[k].b0 : -1.0f;
        bary[(base + k) * 3 + 1] = has ? q[k].b1 : -1.0f;
        bary[(base + k) * 3 + 2] = has ? q[k].b2 : -1.0f;
      }
    }
  }
  return 0;
}
