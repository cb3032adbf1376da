// Statistical outlier verdict on the device (pgdvs_renderer_dyn.py:413-440):
//   thres = median(avg) + std(avg) * std_thres   (torch.median: lower median; torch.std: unbiased)
//   keep  = avg < thres
// over the per-point mean squared distance to the K nearest neighbours (`avg`, knn_grid.cu), one
// cloud per CTA.  The reference does this with torch ops on a compacted cloud, i.e. behind a host
// sync per source pair; here the clouds stay at a fixed capacity (one slot per source pixel, +inf
// = no point) and the number of points never leaves the device.
#include "common.cuh"

namespace pgdvs {

constexpr int kStatThreads = 1024;

__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < kStatThreads / 32; ++w) t += s_red[w];
  return t;
}

__global__ void __launch_bounds__(kStatThreads) k_outlier_keep(const float* __restrict__ avg_all, int64_t n_slots,
                                                              float std_thres, uint8_t* __restrict__ keep_all,
                                                              float* __restrict__ thres_out) {
  __shared__ double s_red[kStatThreads / 32];
  __shared__ unsigned s_hist[256];
  __shared__ unsigned s_prefix, s_rank;
  const float* __restrict__ avg = avg_all + (int64_t)blockIdx.x * n_slots;
  uint8_t* __restrict__ keep = keep_all + (int64_t)blockIdx.x * n_slots;
  const float inf = __int_as_float(0x7f800000);
  // count and mean of the finite entries
  double cnt = 0.0, sum = 0.0;
  for (int64_t i = threadIdx.x; i < n_slots; i += kStatThreads) {
    const float v = avg[i];
    if (v < inf) {  // (false for NaN as well)
      cnt += 1.0;
      sum += (double)v;
    }
  }
  const double n = block_sum(cnt, s_red);
  const double mean = block_sum(sum, s_red) / (n > 0.0 ? n : 1.0);
  double ss = 0.0;
  for (int64_t i = threadIdx.x; i < n_slots; i += kStatThreads) {
    const float v = avg[i];
    if (v < inf) ss += ((double)v - mean) * ((double)v - mean);
  }
  ss = block_sum(ss, s_red);
  // lower median = element of rank (n - 1) / 2: radix select over the bit patterns (avg >= 0, so
  // the pattern orders like the value), 8 bits per round
  if (threadIdx.x == 0) {
    s_prefix = 0u;
    s_rank = (unsigned)((n > 0.0 ? n - 1.0 : 0.0) * 0.5);
  }
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (threadIdx.x < 256) s_hist[threadIdx.x] = 0u;
    __syncthreads();
    const unsigned prefix = s_prefix;
    const unsigned hi_mask = (shift == 24) ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int64_t i = threadIdx.x; i < n_slots; i += kStatThreads) {
      const float v = avg[i];
      if (v < inf) {
        const unsigned b = __float_as_uint(v + 0.0f);
        if ((b & hi_mask) == prefix) atomicAdd(&s_hist[(b >> shift) & 255u], 1u);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned r = s_rank, bin = 255u;
      for (unsigned k = 0; k < 256u; ++k) {
        if (r < s_hist[k]) {
          bin = k;
          break;
        }
        r -= s_hist[k];
      }
      s_rank = r;
      s_prefix = prefix | (bin << shift);
    }
    __syncthreads();
  }
  const float median = __uint_as_float(s_prefix);
  // n < 2: torch.std is NaN, the comparison below is false for every point (like upstream)
  const float sd = (n >= 2.0) ? (float)sqrt(ss / (n - 1.0)) : __int_as_float(0x7fc00000);
  const float thres = (n >= 1.0) ? median + sd * std_thres : __int_as_float(0x7fc00000);
  if (threadIdx.x == 0 && thres_out) thres_out[blockIdx.x] = thres;
  for (int64_t i = threadIdx.x; i < n_slots; i += kStatThreads) keep[i] = (avg[i] < thres) ? 1 : 0;
}

}  // namespace pgdvs

using namespace pgdvs;

extern "C" int pgdvs_outlier_keep(const float* avg, int n_clouds, int64_t n_slots, float std_thres,
                                  uint8_t* keep, float* thres_out, void* stream) {
  if (n_clouds < 0 || n_slots <= 0) return PGDVS_E_BADARG;
  if (n_clouds == 0) return PGDVS_OK;
  if (!avg || !keep) return PGDVS_E_BADARG;
  k_outlier_keep<<<n_clouds, kStatThreads, 0, (cudaStream_t)stream>>>(avg, n_slots, std_thres, keep, thres_out);
  return check_launch();
}
