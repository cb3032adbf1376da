// Statistical-outlier support: K nearest neighbours by brute force.
//
// Replaces `pytorch3d.ops.knn_points(p1, p2, K=knn+1, return_nn=True)` followed by
// `torch.mean(nn_dists[..., skip:], dim=1)` at pgdvs_renderer_dyn.py:405-419 and
// pgdvs_renderer_dyn_track.py:303-318, 345-361 (squared L2, ascending).  The reference only
// consumes the mean of the K smallest squared distances (it computes and discards the indices
// and neighbours), so the fast entry point keeps a sorted K-list of distances in registers
// while reference points stream through shared memory in float4 tiles.  The full
// (dists, idx) variant backs the pytorch3d-compatible `knn_points` facade.
#include "common.cuh"

namespace pgdvs {

constexpr int kKnnThreads = 128;
constexpr int kKnnTile = 1024;
constexpr int kKnnMaxK = 64;

template <bool WITH_IDX>
__global__ void __launch_bounds__(kKnnThreads) k_knn(const float* __restrict__ query, int64_t Q,
                                                     const float* __restrict__ ref, int64_t R, int K,
                                                     int skip, float* __restrict__ mean_out,
                                                     float* __restrict__ dists_out,
                                                     int64_t* __restrict__ idx_out) {
  __shared__ float4 tile[kKnnTile];
  const int64_t qi = (int64_t)blockIdx.x * kKnnThreads + threadIdx.x;
  const bool active = qi < Q;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (active) {
    qx = __ldg(query + qi * 3 + 0);
    qy = __ldg(query + qi * 3 + 1);
    qz = __ldg(query + qi * 3 + 2);
  }
  float best[kKnnMaxK];
  int bidx[WITH_IDX ? kKnnMaxK : 1];
#pragma unroll
  for (int i = 0; i < kKnnMaxK; ++i) {
    best[i] = __int_as_float(0x7f800000);
    if (WITH_IDX) bidx[i] = -1;
  }
  // only the first K slots matter; kth = best[K-1] is tracked in a scalar
  float kth = __int_as_float(0x7f800000);

  for (int64_t base = 0; base < R; base += kKnnTile) {
    const int n = (int)((R - base) < kKnnTile ? (R - base) : kKnnTile);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kKnnThreads) {
      const float* r = ref + (base + i) * 3;
      tile[i] = make_float4(__ldg(r), __ldg(r + 1), __ldg(r + 2), 0.f);
    }
    __syncthreads();
    if (!active) continue;
    for (int i = 0; i < n; ++i) {
      const float4 r = tile[i];
      const float dx = qx - r.x, dy = qy - r.y, dz = qz - r.z;
      const float d = dx * dx + dy * dy + dz * dz;
      if (d < kth) {
        // sorted insert into best[0..K-1] (strict <: earlier index first among equal distances)
        float c = d;
        int ci = (int)(base + i);
#pragma unroll
        for (int j = 0; j < kKnnMaxK; ++j) {
          if (j < K) {
            const float b = best[j];
            const bool lt = c < b;
            best[j] = lt ? c : b;
            c = lt ? b : c;
            if (WITH_IDX) {
              const int bi = bidx[j];
              bidx[j] = lt ? ci : bi;
              ci = lt ? bi : ci;
            }
          }
        }
        float k2 = best[0];
#pragma unroll
        for (int j = 1; j < kKnnMaxK; ++j)
          if (j == K - 1) k2 = best[j];
        kth = k2;
      }
    }
  }
  if (!active) return;
  const int kk = (int)(R < (int64_t)K ? R : (int64_t)K);
  if (mean_out) {
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kKnnMaxK; ++j)
      if (j >= skip && j < kk) sum += best[j];
    // knn_points zero-pads the columns beyond R and the reference's torch.mean divides by all of them
    const int cnt = K - skip;
    mean_out[qi] = cnt > 0 ? sum / (float)cnt : 0.f;
  }
  if (WITH_IDX) {
#pragma unroll
    for (int j = 0; j < kKnnMaxK; ++j) {
      if (j < K) {
        // slots beyond R hold (0, 0) like pytorch3d's zero-initialised outputs
        dists_out[qi * K + j] = (j < kk) ? best[j] : 0.f;
        idx_out[qi * K + j] = (j < kk) ? (int64_t)bidx[j] : (int64_t)0;
      }
    }
  }
}

// knn_grid.cu
size_t knn_grid_workspace_bytes(int64_t R);
int knn_grid_mean_dist(const float* query, int64_t Q, const float* ref, int64_t R, int K, int skip,
                       float* mean_out, void* workspace, cudaStream_t stream);
constexpr int64_t kKnnGridMinPoints = 4096;  // below this the brute-force kernel is faster

}  // namespace pgdvs

using namespace pgdvs;

extern "C" int pgdvs_knn_workspace_bytes(int64_t Q, int64_t R, size_t* bytes) {
  if (!bytes || Q < 0 || R < 0) return PGDVS_E_BADARG;
  // scratch of the uniform-grid search (knn_grid.cu); small clouds use the brute-force kernel,
  // which needs none
  *bytes = (R >= kKnnGridMinPoints) ? knn_grid_workspace_bytes(R) : 256;
  return PGDVS_OK;
}

extern "C" int pgdvs_knn_mean_dist(const float* query, int64_t Q, const float* ref, int64_t R, int K,
                                   int skip_first, float* mean_out, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  if (Q < 0 || R < 0 || K < 1 || skip_first < 0) return PGDVS_E_BADARG;
  if (K > kKnnMaxK) return PGDVS_E_KNN_K;
  if (Q == 0) return PGDVS_OK;
  if (!query || !mean_out || (R > 0 && !ref)) return PGDVS_E_BADARG;
  // exact uniform-grid search when the caller provides the scratch (pgdvs_knn_workspace_bytes);
  // without it, or for small clouds, the brute-force kernel below gives the same answer
  if (R >= kKnnGridMinPoints && R >= K && R < (int64_t)INT32_MAX && workspace != nullptr &&
      (reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && workspace_bytes >= knn_grid_workspace_bytes(R))
    return knn_grid_mean_dist(query, Q, ref, R, K, skip_first, mean_out, workspace, (cudaStream_t)stream);
  const unsigned grid = (unsigned)((Q + kKnnThreads - 1) / kKnnThreads);
  k_knn<false><<<grid, kKnnThreads, 0, (cudaStream_t)stream>>>(query, Q, ref, R, K, skip_first, mean_out,
                                                               nullptr, nullptr);
  return check_launch();
}

extern "C" int pgdvs_knn_points(const float* query, int64_t Q, const float* ref, int64_t R, int K,
                                float* dists_out, int64_t* idx_out, void* stream) {
  if (Q < 0 || R < 0 || K < 1) return PGDVS_E_BADARG;
  if (K > kKnnMaxK) return PGDVS_E_KNN_K;
  if (Q == 0) return PGDVS_OK;
  if (!query || !dists_out || !idx_out || (R > 0 && !ref)) return PGDVS_E_BADARG;
  const unsigned grid = (unsigned)((Q + kKnnThreads - 1) / kKnnThreads);
  k_knn<true><<<grid, kKnnThreads, 0, (cudaStream_t)stream>>>(query, Q, ref, R, K, 0, nullptr, dists_out,
                                                              idx_out);
  return check_launch();
}
