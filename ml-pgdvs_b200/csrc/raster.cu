// Tiled rasterize-and-composite over the cell-sorted records produced by bin.cu / uwp.cu.
//
// Replaces pytorch3d's RasterizePointsNaiveCudaKernel (O(N*H*W*P), reached from
// pgdvs_renderer_dyn.py:690-717 with bin_size=0), PointsRenderer's weight computation,
// the NormWeighted/Alpha compositors, the background fill and PGDVS's second all-ones
// render for the mask (pgdvs_renderer_dyn.py:719-722) with ONE pass:
//   * each thread walks one pixel and keeps its K nearest hits sorted in registers — as ONE
//     32-bit key per slot (z pattern + candidate ordinal, see KeyList) in the tile kernel, as
//     (z, slot) pairs (PairList) in the generic kernel;
//   * candidates are only the records filed under cells within `halo` of the pixel — for
//     every window row that is one contiguous run of the cell-sorted record array;
//   * the hit test reproduces the reference arithmetic bit for bit (dist2_rn, strict <);
//   * insertion is a min/max chain without payload moves (tile kernel: branch-free, two
//     candidates per step); exact fp32 z ties (which the CPU rasterizer's (z, idx, dist2)
//     priority queue orders by the smaller packed index) and keys that agree on their whole z
//     part are DETECTED and the few affected pixels are redone by an exact (z, idx) rescan, so
//     idx/zbuf/dists stay deterministic and bit-exact.
//
// Two kernels share that logic:
//   k_raster_tile  (halo 1..3, scalar radius, K <= 32 — the PGDVS configurations): a 32x8-pixel
//                  CTA fetches the 8 + 2*halo row runs its pixels can touch into shared memory
//                  with 1-D TMA bulk copies (cp.async.bulk + mbarrier; the runs ARE contiguous
//                  because records are sorted by cell), re-assigns its pixels to threads in
//                  order of candidate count, walks them with shared-memory reads only, hands the
//                  winners back to each pixel's own thread and runs the epilogue in raster order.
//   k_raster_cells (any halo, per-point radii, K <= 150): candidates are read through L1.  When a
//                  window holds many more records than K, k_sort_cells first puts the records of
//                  every small cell in ascending z order (in place) and the walk — one flattened
//                  loop, a cursor per lane — leaves a cell at the first record behind the list's
//                  last element.
// DESIGN.md 4.1 walks through the tile kernel step by step.
#include "common.cuh"

#include <stdlib.h>

namespace pgdvs {

struct RasterParams {
  const int* cell_end;  // per-cell END offsets; start(c) = cell_end[c - 1] (cell_end[-1] == 0)
  const float4* recA;   // 32-byte records; part A (x_ndc, y_ndc, z, packed idx) at [rec_a(j)],
                        // part B (f0, f1, f2, f3 | radius) at [rec_b(j)]  (common.cuh)
  int N, H, W, K, C, halo, GW, GH;
  NdcAxis ax, ay;
  float r2;          // scalar radius^2 (fp32 r*r) or < 0: per-point radius in recB.w
  float rr_weight;   // divisor of PointsRenderer's  1 - dists/(r*r)
  float inv_rr_weight;
  int compositor;
  float bg[4];
  const float* static_rgb;
  int32_t* idx;
  float* zbuf;
  float* dists;
  float* image;
  float* mask;
  int smem_records;  // capacity of the staging buffer of k_raster_tile, in records
  double density;    // host-side hint: mean points per pixel (sizes the staging buffer)
  int cells_sorted;  // 1: every cell of at most kSortCap records is in ascending z order (k_sort_cells)
  int force_generic; // developer switch (PGDVS_RASTER_FORCE_GENERIC): skip the tile kernel
};

// Cells holding up to kSortCap records can be put in ascending z order (k_sort_cells, run by the
// generic path when a pixel sees many more candidates than it keeps): the walk then leaves a cell
// at the first record that lies behind the K-th hit so far, because the rest of the cell does too.
// Measured on C5 (1080p, 8 points per pixel): K = 8, r = 0.01 (121-cell window) 21.7 -> 11.1 ms per
// 4 views, K = 32, r = 0.02 (529 cells) 144 -> 71 ms per 2 views, the sort itself 0.9 / 0.5 ms.
constexpr int kSortCap = 16;

constexpr float kInf = __builtin_huge_valf();

// candidate (z, idx) strictly before list element (ze, slot se)?  Total order (z, idx).
__device__ __forceinline__ bool cand_less(float z, int idx, float ze, int se,
                                          const float4* __restrict__ recA) {
  bool lt = z < ze;
  if (z == ze) {  // exact fp32 z tie -> smaller packed index first
    lt = (se < 0) ? true : (idx < __float_as_int(recA[rec_a(se)].w));
  }
  return lt;
}

// Fast path: the K nearest hits of a pixel as ONE sorted 32-bit key per slot.
//   key = (((bits(z) - base) >> sh) << bits) | t
// t = ordinal of the candidate in the pixel's own walk (`bits` = ceil(log2(#candidates)), 5 or 6
// at PGDVS densities).  z >= 0, so its bit pattern orders like the value and so does the
// integer difference to `base`.  The tile kernel takes base = the smallest z pattern among the
// records it staged: whenever the tile's z range spans fewer than 2^(32-bits) patterns (z_max
// / z_min up to ~256 with 6 payload bits) sh is 0 and the key order is the EXACT z order;
// otherwise (and in the unstaged/generic paths, base = 0) the low `sh` bits are dropped.
// Insertion is branch-free and carries no payload:
//   k'[i] = min(max(c, k[i-1]), k[i])        2 integer min/max per slot, all slots independent
// (a compare-exchange on separate (z, slot) registers costs 5 ALU ops per slot and a serial
// carry).  A candidate that misses the radius test is inserted as kEmpty, which leaves the
// list unchanged, so the candidate loop has no divergent branch at all.
// Keys can only misorder hits whose z agree on all kept bits.  `ambiguous()` detects every
// such case that could matter — two neighbouring kept keys, or the last kept key and the
// smallest rejected/evicted one (`rej`), with equal z part — and those pixels (exact fp32 ties
// included) are redone by rescan_exact() in the full (z, idx) order of the CPU rasterizer's
// priority queue.  Everything else is provably identical to that order.
constexpr uint32_t kEmpty = 0xFFFFFFFFu;
#ifndef PGDVS_RASTER_BRANCHFREE_MAXK
#define PGDVS_RASTER_BRANCHFREE_MAXK 8
#endif

struct KeyCode {
  uint32_t base, mask;
  int sh, bits;
  // n_candidates: most candidates any pixel using this code walks; [zlo, zhi]: range of the z
  // bit patterns it will see (0 .. 0x7fffffff when unknown)
  __device__ __forceinline__ void init(int n_candidates, uint32_t zlo, uint32_t zhi) {
    bits = 32 - __clz(max(n_candidates - 1, 0));
    mask = (1u << bits) - 1u;  // n_candidates < 2^31
    base = zlo;
    const uint32_t range = zhi - zlo;
    sh = max(0, bits - __clz(range));
    // the all-ones z part with t == mask would collide with kEmpty
    if (((range >> sh) << bits | mask) == kEmpty) ++sh;
  }
  // hit -> key; +0.0f maps -0.0 (which the reference does not cull) onto +0.0
  __device__ __forceinline__ uint32_t encode(bool hit, float z, uint32_t t) const {
    const uint32_t zb = __float_as_uint(__fadd_rn(z, 0.0f));
    return hit ? ((((zb - base) >> sh) << bits) | t) : kEmpty;
  }
  __device__ __forceinline__ bool same_z(uint32_t a, uint32_t b) const { return ((a ^ b) & ~mask) == 0u; }
};

template <int KP>
struct KeyList {
  uint32_t k[KP];
  uint32_t rej;  // smallest key that was rejected or evicted
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < KP; ++i) k[i] = kEmpty;
    rej = kEmpty;
  }
  __device__ __forceinline__ void insert(uint32_t c) {
    uint32_t prev = k[0];
    k[0] = min(c, prev);
#pragma unroll
    for (int i = 1; i < KP; ++i) {
      const uint32_t cur = k[i];
      k[i] = min(max(c, prev), cur);
      prev = cur;
    }
    rej = min(rej, max(c, prev));
  }
  // Two candidates at once: the i-th smallest of list + {lo, hi} is
  //   min(k[i], max(k[i-1], lo), max(k[i-2], hi))
  // i.e. 3 ops per slot per PAIR with the 3-input integer min (VIMNMX3) instead of 2 per
  // slot per candidate.
  __device__ __forceinline__ void insert2(uint32_t c1, uint32_t c2) {
    static_assert(KP >= 2, "insert2 needs two slots");
    const uint32_t lo = min(c1, c2), hi = max(c1, c2);
    uint32_t p2 = k[0], p1 = k[1];  // old k[i-2], k[i-1]
    k[0] = min(p2, lo);
    k[1] = min(min(p1, max(p2, lo)), hi);
#pragma unroll
    for (int i = 2; i < KP; ++i) {
      const uint32_t cur = k[i];
      k[i] = min(min(cur, max(p1, lo)), max(p2, hi));
      p2 = p1;
      p1 = cur;
    }
    // the two that fall off: the smaller of them is the KP-th smallest of the union
    rej = min(min(rej, max(p1, lo)), max(p2, hi));
  }
  // small K: always the branch-free chain (a miss is kEmpty and changes nothing);
  // large K: skip the 2*KP-op chain for keys that cannot enter the list
  __device__ __forceinline__ void push(uint32_t c) {
    if (KP <= PGDVS_RASTER_BRANCHFREE_MAXK) {
      insert(c);
    } else if (c < k[KP - 1]) {
      insert(c);
    } else {
      rej = min(rej, c);
    }
  }
  // K == KP
  __device__ __forceinline__ bool ambiguous_full(const KeyCode& kc) const {
    bool amb = (rej != kEmpty) & kc.same_z(rej, k[KP - 1]);
#pragma unroll
    for (int i = 1; i < KP; ++i) amb |= (k[i] != kEmpty) & kc.same_z(k[i], k[i - 1]);
    return amb;
  }
  __device__ __forceinline__ bool ambiguous(int K, const KeyCode& kc) const {
    bool amb = false;
#pragma unroll
    for (int i = 1; i < KP; ++i)
      if (i <= K) amb = amb || (k[i] != kEmpty && kc.same_z(k[i], k[i - 1]));
    if (K >= KP) amb = amb || (rej != kEmpty && kc.same_z(rej, k[KP - 1]));
    return amb;
  }
};

// General path (long lists, unknown z range): (z, record slot) pairs ordered by z alone with a
// compare-exchange chain that only runs for candidates nearer than the current last element.
// Exact fp32 z ties that could change the result are DETECTED, not resolved: `tie` is set when a
// rejected or evicted candidate had the z of the (then) last kept element — which is implied
// whenever it ties with the final one — and ambiguous() adds ties between kept neighbours.
// Flagged pixels are redone by rescan_exact().
template <int KP>
struct PairList {
  float z[KP];
  int s[KP];
  bool tie;
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < KP; ++i) {
      z[i] = kInf;
      s[i] = -1;
    }
    tie = false;
  }
  __device__ __forceinline__ void push(bool hit, float cz, int cs) {
    if (!hit) return;
    if (cz < z[KP - 1]) {
#pragma unroll
      for (int i = 0; i < KP; ++i) {
        const bool p = cz < z[i];
        const float tz = z[i];
        const int ts = s[i];
        z[i] = p ? cz : tz;
        s[i] = p ? cs : ts;
        cz = p ? tz : cz;
        cs = p ? ts : cs;
      }
      tie = tie | ((cs >= 0) & (cz == z[KP - 1]));  // (cz, cs) is now the evicted element
    } else {
      tie = tie | (cz == z[KP - 1]);
    }
  }
  __device__ __forceinline__ bool ambiguous(int K) const {
    bool amb = (K >= KP) & tie;
#pragma unroll
    for (int i = 1; i < KP; ++i)
      if (i <= K) amb = amb | ((s[i] >= 0) & (z[i] == z[i - 1]));
    return amb;
  }
};

// Exact path: full (z, idx) order, (z, slot) pairs.
template <int KP>
struct ExactList {
  float z[KP];
  int s[KP];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < KP; ++i) {
      z[i] = kInf;
      s[i] = -1;
    }
  }
  __device__ __forceinline__ void insert_exact(float cz, int cidx, int cslot,
                                               const float4* __restrict__ recA) {
    if (!cand_less(cz, cidx, z[KP - 1], s[KP - 1], recA)) return;
    bool p_hi = true;
#pragma unroll
    for (int i = KP - 1; i > 0; --i) {
      const bool p_lo = cand_less(cz, cidx, z[i - 1], s[i - 1], recA);
      z[i] = p_lo ? z[i - 1] : (p_hi ? cz : z[i]);
      s[i] = p_lo ? s[i - 1] : (p_hi ? cslot : s[i]);
      p_hi = p_lo;
    }
    if (p_hi) {
      z[0] = cz;
      s[0] = cslot;
    }
  }
};

struct PixelCtx {
  float xf, yf, r2;
};

// PPR = per-point radius (recB.w); the scalar-radius instantiation carries no extra load/branch
template <bool PPR>
__device__ __forceinline__ bool hit_test(const PixelCtx& c, const float4 a,
                                         const float4* __restrict__ rec, int j) {
  const float d2 = dist2_rn(a.x, a.y, c.xf, c.yf);
  float r2 = c.r2;
  if (PPR) {
    const float r = __ldg(&rec[rec_b(j)].w);
    r2 = __fmul_rn(r, r);
  }
  return d2 < r2;
}

// Exact rescan of one pixel over the GLOBAL records (rare: only pixels KeyList::ambiguous()
// flags).  Out of line so that the fast path stays small; returns global record slots.
template <int KP, bool PPR>
__device__ __noinline__ void rescan_exact(const RasterParams& p, const PixelCtx& c, int n, int x,
                                          int y, int* sout) {
  ExactList<KP> q;
  q.init();
  const int span = 2 * p.halo + 1;
  for (int ry = 0; ry < span; ++ry) {
    const int64_t cell0 = ((int64_t)n * p.GH + (y + ry)) * p.GW + x;
    const int s = __ldg(p.cell_end + cell0 - 1);
    const int e = __ldg(p.cell_end + cell0 + span - 1);
    for (int j = s; j < e; ++j) {
      const float4 a = __ldg(p.recA + rec_a(j));
      if (hit_test<PPR>(c, a, p.recA, j)) q.insert_exact(a.z, __float_as_int(a.w), j, p.recA);
    }
  }
#pragma unroll
  for (int i = 0; i < KP; ++i) sout[i] = q.s[i];
}

// ---------------------------------------------------------------------------------------
// Epilogue shared by both kernels.  `rec` points at the records the slots refer to (global
// array, or the CTA's shared-memory staging buffer — the call sites keep the two apart so that
// the staged case compiles to LDS); s[k] < 0 = empty.  FULLK: K == KP (every loop bound is a
// compile-time constant and fragments are written with 128-bit stores).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

template <int KP>
struct Slots {
  int s[KP];
};

// Where the epilogue reads the winners' records from.  The shared-memory source addresses the
// staging buffer with 32-bit shared addresses and three integer ops per access
// (base + 32 j, then OR in bit 4 = bit 2 of j; the buffer is 32-byte aligned).
struct GlobalRecords {
  const float4* p;
  __device__ __forceinline__ float4 a(int j) const { return p[rec_a(j)]; }
  __device__ __forceinline__ float4 b(int j) const { return p[rec_b(j)]; }
};
struct StagedRecords {
  uint32_t base;  // shared-state-space address of the staging buffer
  static __device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
  }
  __device__ __forceinline__ uint32_t addr_a(int j) const {
    // spelled in PTX: the C++ form gets "simplified" into twice as many instructions
    uint32_t addr;
    asm("{\n\t.reg .b32 t, a;\n\t"
        "shl.b32 t, %1, 2;\n\t"
        "mad.lo.u32 a, %1, 32, %2;\n\t"
        "lop3.b32 %0, a, t, 16, 0xF8;\n\t}"  // a | (t & 16)
        : "=r"(addr)
        : "r"(j), "r"(base));
    return addr;
  }
  __device__ __forceinline__ float4 a(int j) const { return lds128(addr_a(j)); }
  __device__ __forceinline__ float4 b(int j) const { return lds128(addr_a(j) ^ 16u); }
};

// SPEC: the PGDVS configuration (NormWeightedCompositor over rgb) with compositor and channel
// count as compile-time constants — no per-winner mode branches, accumulators stay in registers.
template <int KP, bool FULLK, typename Records, bool SPEC = false>
__device__ __forceinline__ void pixel_epilogue(const RasterParams& p, const Slots<KP>& sl,
                                               const PixelCtx& c, int n, int x, int y,
                                               const Records rec) {
  const int K = FULLK ? KP : p.K;
  const int64_t pix = ((int64_t)n * p.H + y) * p.W + x;
  const int mode = SPEC ? (int)PGDVS_COMPOSITE_NORM_WEIGHTED : p.compositor;
  const int nch = SPEC ? 3 : p.C;
  // the static frame is only needed at the very end: fetch it now, use it after the K loop
  float st[4] = {0.f, 0.f, 0.f, 0.f};
  const bool blend = (p.static_rgb != nullptr) && (p.image != nullptr) && (mode != PGDVS_COMPOSITE_NONE);
  if (blend) {
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
      if (ch < nch) st[ch] = __ldg(p.static_rgb + pix * nch + ch);
  }
  // One pass over the K winners: fragments (idx / zbuf / dists, bit-exact), the PointsRenderer
  // weight 1 - dists/(r*r) (as a multiply by the fp32 reciprocal, like torch's CUDA division by
  // a scalar) and the K-ordered compositor sums.  Norm-weighted sums are accumulated
  // unnormalised and scaled once by 1/max(sum w, 1e-4) at the end; images are
  // tolerance-matched (|delta| <= 1e-5), not bit-matched.
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float wsum = 0.f;       // the same compositor applied to all-ones features (mask render)
  float cum_alpha = 1.0f;
  constexpr bool vec4 = FULLK && (KP % 4 == 0);  // K*4 B per pixel is a multiple of 16 B
#pragma unroll
  for (int k0 = 0; k0 < KP; k0 += 4) {
    int o_idx[4];
    float o_z[4], o_d[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int k = k0 + kk;
      o_idx[kk] = -1;
      o_z[kk] = -1.0f;
      o_d[kk] = -1.0f;
      if (k < KP) {
        const int s = (k < K) ? sl.s[k] : -1;
        if (s >= 0) {
          const float4 a = rec.a(s);
          o_d[kk] = dist2_rn(a.x, a.y, c.xf, c.yf);
          o_idx[kk] = __float_as_int(a.w);
          o_z[kk] = a.z;
          if (mode != PGDVS_COMPOSITE_NONE) {
            const float4 f4 = rec.b(s);
            const float w = __fsub_rn(1.0f, __fmul_rn(o_d[kk], p.inv_rr_weight));
            float wk = w;
            if (mode == PGDVS_COMPOSITE_ALPHA) {
              wk = __fmul_rn(cum_alpha, w);
              cum_alpha = __fmul_rn(cum_alpha, __fsub_rn(1.0f, w));
            }
            acc[0] = __fadd_rn(acc[0], __fmul_rn(wk, f4.x));
            acc[1] = __fadd_rn(acc[1], __fmul_rn(wk, f4.y));
            acc[2] = __fadd_rn(acc[2], __fmul_rn(wk, f4.z));
            acc[3] = __fadd_rn(acc[3], __fmul_rn(wk, f4.w));
            wsum = __fadd_rn(wsum, wk);
          }
        }
      }
    }
    if (vec4) {
      const int64_t o = pix * KP + k0;
      if (p.idx) *reinterpret_cast<int4*>(p.idx + o) = make_int4(o_idx[0], o_idx[1], o_idx[2], o_idx[3]);
      if (p.zbuf) *reinterpret_cast<float4*>(p.zbuf + o) = make_float4(o_z[0], o_z[1], o_z[2], o_z[3]);
      if (p.dists) *reinterpret_cast<float4*>(p.dists + o) = make_float4(o_d[0], o_d[1], o_d[2], o_d[3]);
    } else {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = k0 + kk;
        if (k < K) {
          if (p.idx) p.idx[pix * K + k] = o_idx[kk];
          if (p.zbuf) p.zbuf[pix * K + k] = o_z[kk];
          if (p.dists) p.dists[pix * K + k] = o_d[kk];
        }
      }
    }
  }
  if (mode == PGDVS_COMPOSITE_NONE) return;
  float ones_acc = wsum;
  if (mode == PGDVS_COMPOSITE_NORM_WEIGHTED) {
    if (wsum < 0.25f && sl.s[0] >= 0) {
      // ill-conditioned normalisation (every hit sits near the rim of its splat): the 1-ulp
      // slack of the reciprocal would be amplified by 1/sum(w), so redo these few pixels with
      // the true division
      wsum = 0.f;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) acc[ch] = 0.f;
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        const int s = (k < K) ? sl.s[k] : -1;
        if (s >= 0) {
          const float4 a = rec.a(s);
          const float4 f4 = rec.b(s);
          const float w = __fsub_rn(1.0f, __fdiv_rn(dist2_rn(a.x, a.y, c.xf, c.yf), p.rr_weight));
          acc[0] = __fadd_rn(acc[0], __fmul_rn(w, f4.x));
          acc[1] = __fadd_rn(acc[1], __fmul_rn(w, f4.y));
          acc[2] = __fadd_rn(acc[2], __fmul_rn(w, f4.z));
          acc[3] = __fadd_rn(acc[3], __fmul_rn(w, f4.w));
          wsum = __fadd_rn(wsum, w);
        }
      }
    }
    const float inv_t = __frcp_rn(fmaxf(wsum, 1e-4f));
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) acc[ch] = __fmul_rn(acc[ch], inv_t);
    ones_acc = __fmul_rn(wsum, inv_t);
  }
  const bool is_bg = sl.s[0] < 0;  // _add_background_color_to_images: idx[:, 0] < 0
  const float m = (ones_acc > 0.0f) ? 1.0f : 0.0f;
  if (p.mask) p.mask[pix] = m;
  if (p.image) {
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      if (ch < nch) {
        float v = is_bg ? p.bg[ch] : acc[ch];
        // combined = (1 - mask) * static + mask * dyn   (pgdvs_renderer.py:169-172)
        if (blend) v = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, m), st[ch]), __fmul_rn(m, v));
        p.image[pix * nch + ch] = v;
      }
    }
  }
}

// Out-of-line finish over the GLOBAL records: pixels whose keys were ambiguous (rescan = true,
// exact (z, idx) rescan first) and tiles that could not be staged (slots are global already).
// Keeps the rare paths, and their generic-address loads, out of the hot kernels.
template <int KP, bool PPR>
__device__ __noinline__ void finish_global(const RasterParams& p, const PixelCtx c, int n, int x,
                                           int y, Slots<KP> sl, bool rescan) {
  if (rescan) rescan_exact<KP, PPR>(p, c, n, x, y, sl.s);
  if (KP <= 32 && p.K == KP)
    pixel_epilogue<KP, true>(p, sl, c, n, x, y, GlobalRecords{p.recA});
  else
    pixel_epilogue<KP, false>(p, sl, c, n, x, y, GlobalRecords{p.recA});
}

// Resident CTAs per SM the generic kernel is compiled for.  Its walk is a chain of dependent
// loads, so it wants warps more than registers (the K-list itself takes 2*KP of them): 48 / 64 /
// 128 registers for KP <= 8 / 16 / 32 instead of the 72 / 96 / 168 ptxas takes when left alone
// (3 / 2 / 1 CTAs).  Measured on C5 (same bits): K = 8 11.1 -> 9.1 ms, K = 16 15.3 -> 11.4 ms,
// K = 32 71.3 -> 50.7 ms; one step further (40 / 48 registers, or 80 for KP = 32) spills the list.
#ifndef PGDVS_RASTER_MINBLOCKS
#define PGDVS_RASTER_MINBLOCKS ((KP <= 8) ? 5 : ((KP <= 16) ? 4 : ((KP <= 32) ? 2 : 1)))
#endif

// ---------------------------------------------------------------------------------------
// Generic kernel: any halo, optional per-point radii, candidates read through L1.
// ---------------------------------------------------------------------------------------
template <int KP, bool PPR>
__global__ void __launch_bounds__(256, PGDVS_RASTER_MINBLOCKS) k_raster_cells(const __grid_constant__ RasterParams p) {
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  const int n = blockIdx.z;
  if (x >= p.W || y >= p.H) return;
  PixelCtx c;
  c.xf = pixel_center_ndc(p.ax, x);
  c.yf = pixel_center_ndc(p.ay, y);
  c.r2 = p.r2;
  const float4* __restrict__ recA = p.recA;

  const int span = 2 * p.halo + 1;
  // cs[c] = start of cell (x + c) of the first window row = cell_end[... - 1]
  const int* __restrict__ cs = p.cell_end + ((int64_t)n * p.GH + y) * p.GW + x - 1;
  PairList<KP> q;
  q.init();
  if (p.cells_sorted) {
    // z-sorted cells, one flattened loop: every lane keeps its own cursor (window row, cell,
    // record) and leaves a sorted cell at the first record behind its list's last element, so a
    // warp runs for as long as its busiest lane has records to look at — not for the longest
    // cell of every step, which is what a per-cell loop nest costs once lanes exit early
    const int* __restrict__ row = cs;  // boundaries of the cells of window row ry: row[0 .. span]
    int ry = 0, cx = 0;
    int j = __ldg(row), e = j;         // nothing open yet: the first trip opens cell 1
    bool sorted = false;
    for (;;) {
      if (j >= e) {  // open the next cell (j == start of it: the cells of a row are contiguous)
        if (++cx > span) {
          if (++ry >= span) break;
          row += p.GW;
          cx = 1;
          j = __ldg(row);
        }
        e = __ldg(row + cx);
        sorted = (e - j) <= kSortCap;
      }
      if (j < e) {
        const float4 a = __ldg(recA + rec_a(j));
        const int slot = j++;
        if (a.z <= q.z[KP - 1])
          q.push(hit_test<PPR>(c, a, recA, slot), a.z, slot);
        else if (sorted)
          j = e;
      }
    }
  } else {
    for (int ry = 0; ry < span; ++ry) {
      const int s = __ldg(cs + (int64_t)ry * p.GW);
      const int e = __ldg(cs + (int64_t)ry * p.GW + span);
      for (int j = s; j < e; ++j) {
        const float4 a = __ldg(recA + rec_a(j));
        // depth first: once the list is full most candidates lie behind its last element and
        // need no distance test at all (equal z still goes through: it may be a tie)
        if (a.z <= q.z[KP - 1]) q.push(hit_test<PPR>(c, a, recA, j), a.z, j);
      }
    }
  }
  Slots<KP> sl;
#pragma unroll
  for (int i = 0; i < KP; ++i) sl.s[i] = q.s[i];
  if (q.ambiguous(p.K)) {
    finish_global<KP, PPR>(p, c, n, x, y, sl, true);
    return;
  }
  if (KP <= 32 && p.K == KP)
    pixel_epilogue<KP, true>(p, sl, c, n, x, y, GlobalRecords{recA});
  else
    pixel_epilogue<KP, false>(p, sl, c, n, x, y, GlobalRecords{recA});
}

// ---------------------------------------------------------------------------------------
// TMA-staged tile kernel: small halos (compile-time HALO = 1..3), scalar radius.
// ---------------------------------------------------------------------------------------
constexpr int kTileW = 32, kTileH = 8;

// tile kernel: its lists (K <= 32) are sorted 32-bit keys, branch-free up to
// PGDVS_RASTER_BRANCHFREE_MAXK and with a "can it enter?" pre-test above that
#ifndef PGDVS_TILE_KEYS_MAXK
#define PGDVS_TILE_KEYS_MAXK 32
#endif
#ifndef PGDVS_TILE_MINBLOCKS_K8
#define PGDVS_TILE_MINBLOCKS_K8 5
#endif

template <int KP, int HALO>
__global__ void __launch_bounds__(256, (KP <= 8) ? PGDVS_TILE_MINBLOCKS_K8 : ((KP <= 16) ? 2 : 1)) k_raster_tile(const __grid_constant__ RasterParams p) {
  constexpr int SPAN = 2 * HALO + 1;     // window rows / cells per pixel
  constexpr int ROWS = kTileH + 2 * HALO;  // extended-grid rows the tile's pixels can touch
  static_assert(ROWS <= 32, "one lane per tile row");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* s_rec = reinterpret_cast<float4*>(smem_raw);  // staged records (32 bytes each)
  const StagedRecords staged_rec{smem_u32(smem_raw)};
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ int s_delta[ROWS];  // smem record index = global record index + s_delta[row]
  __shared__ int2 s_row[ROWS];   // (first shared slot, records) of every staged row
  __shared__ int s_staged;       // records staged (> 0: the tile's runs fit and are being copied)
  __shared__ uint32_t s_zlo, s_zhi;  // range of the staged z bit patterns (KeyCode)
  __shared__ int s_max;          // most candidates any pixel of the tile walks
  __shared__ int s_hist[64];
  __shared__ unsigned char s_perm[256];
  __shared__ float s_xf[kTileW], s_yf[kTileH];  // NDC pixel centres of the tile's columns / rows
  // one scratch area, two lives: before the walk, per pixel the (start, length) of each window-row
  // run; after it (separated by barriers), per pixel its K winners as 16-bit staged slots
  constexpr int kScratchBytes = (SPAN * 256 * 8 > 256 * KP * 2) ? SPAN * 256 * 8 : 256 * KP * 2;
  __shared__ __align__(16) unsigned char s_scratch[kScratchBytes];
  int2(*s_runs)[256] = reinterpret_cast<int2(*)[256]>(s_scratch);
  uint16_t* s_win = reinterpret_cast<uint16_t*>(s_scratch);

  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  const int n = blockIdx.z;
  int x = x0 + threadIdx.x, y = y0 + threadIdx.y;

  // ---- every global read of the prologue is issued first; the set-up work below (mbarrier, pixel
  //      centre tables, first barrier) runs under their latency
  int rs[SPAN], rl[SPAN];  // start (global record index) and length of each window-row run
#pragma unroll
  for (int r = 0; r < SPAN; ++r) {
    rs[r] = 0;
    rl[r] = 0;
  }
  if (x < p.W && y < p.H) {
    // coalesced read of the runs of the thread's own (identity) pixel; the thread that ends up
    // walking the pixel picks them up from shared memory
    const int* __restrict__ cs = p.cell_end + ((int64_t)n * p.GH + y) * p.GW + x - 1;
#pragma unroll
    for (int r = 0; r < SPAN; ++r) {
      rs[r] = __ldg(cs + r * p.GW);
      rl[r] = __ldg(cs + r * p.GW + SPAN);  // end of the run for now
    }
  }
  int row_gs = 0, row_ge = 0;  // warp 0, lane r: bounds of staged row r
  if (threadIdx.y == 0 && threadIdx.x < ROWS && y0 + (int)threadIdx.x < p.GH) {
    // extended-grid row (y0 + lane) holds image row y0 + lane - HALO; cells x0 .. x0+31+2*HALO
    const int64_t rb = ((int64_t)n * p.GH + y0 + threadIdx.x) * p.GW;
    const int xe = min(x0 + kTileW - 1 + 2 * HALO, p.GW - 1);
    row_gs = __ldg(p.cell_end + rb + x0 - 1);
    row_ge = __ldg(p.cell_end + rb + xe);
  }

  if (tid == 0) {
    // one arrival (the expect_tx below); the copies complete the transaction count
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_zlo = 0xFFFFFFFFu;
    s_zhi = 0u;
    s_max = 0;
  }
  if (tid < 64) s_hist[tid] = 0;
  // pixel centres once per CTA (each costs an IEEE division); warps 1 and 2, warp 0 is busy below
  if (tid >= 32 && tid < 32 + kTileW) s_xf[tid - 32] = pixel_center_ndc(p.ax, x0 + tid - 32);
  if (tid >= 64 && tid < 64 + kTileH) s_yf[tid - 64] = pixel_center_ndc(p.ay, y0 + tid - 64);
  __syncthreads();

  // ---- warp 0 first: size the row runs, decide, arm the barrier, issue the bulk copies; they
  //      are in flight while the CTA sorts its pixels below
  if (threadIdx.y == 0) {
    const int lane = threadIdx.x;
    const int gs = row_gs, ge = row_ge;
    const int len = ge - gs;
    // Row r is staged at shared slot base_r + (gs & 7), base_r a multiple of 8: shared slot and
    // global slot of a record then differ by a multiple of 8, so rec_a()/rec_b() pick the same
    // half of the 32-byte record on both sides.  pad_len = footprint incl. that lead-in, x8.
    const int pad_len = (len > 0) ? (((gs & 7) + len + 7) & ~7) : 0;
    int inc = pad_len;  // inclusive prefix over the rows -> smem offsets
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    const int total = __shfl_sync(0xffffffffu, inc, ROWS - 1);  // padded footprint of the tile
    int cnt = len;  // records actually copied
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    const bool fits = total <= p.smem_records;
    const int dst_slot = (inc - pad_len) + (gs & 7);
    if (lane < ROWS) {
      s_delta[lane] = fits ? dst_slot - gs : 0;
      s_row[lane] = make_int2(dst_slot, len);
    }
    if (lane == 0) {
      s_staged = fits ? cnt : 0;
      if (fits) {
        const uint32_t bytes = (uint32_t)cnt * 32u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(bytes)
                     : "memory");
      }
    }
    __syncwarp();
    if (fits && lane < ROWS && len > 0) {
      const float4* src = p.recA + (int64_t)2 * gs;
      float4* dst = s_rec + 2 * dst_slot;
      const uint32_t bytes = (uint32_t)len * 32u;
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
              smem_u32(dst)),
          "l"(src), "r"(bytes), "r"(smem_u32(&s_bar))
          : "memory");
    }
  }

  // ---- balance the warps: hand the tile's pixels to threads in order of their candidate count
  //      (counting sort in shared memory) so that the lanes of a warp walk runs of similar
  //      length; the largest count also sizes the payload field of the keys
  int mine = tid;          // tile-local index (row * 32 + column) of the pixel this thread walks
  {
    int work = 0;
#pragma unroll
    for (int r = 0; r < SPAN; ++r) {
      rl[r] -= rs[r];  // end -> length
      work += rl[r];
    }
#ifndef PGDVS_RASTER_NO_SORT
#pragma unroll
    for (int r = 0; r < SPAN; ++r) s_runs[r][tid] = make_int2(rs[r], rl[r]);
#endif
    const int wmax = __reduce_max_sync(0xffffffffu, work);
    if ((tid & 31) == 0) atomicMax(&s_max, wmax);
#ifdef PGDVS_TILE_FIXED_BINS
    // (experiment, DESIGN.md 9 item 1a: one CTA barrier less) bin width from the host's density
    // hint instead of the tile's own maximum, which is then first needed after the later barriers
    const int width = (int)(p.density * (double)(SPAN * SPAN) * (2.5 / 64.0)) + 1;
    const int bin = max(0, 63 - work / width);  // heaviest pixels first
#else
    __syncthreads();  // s_max; also s_staged / s_delta of warp 0
#endif
#ifndef PGDVS_RASTER_NO_SORT
#ifndef PGDVS_TILE_FIXED_BINS
    const int width = s_max / 64 + 1;
    const int bin = 63 - work / width;  // heaviest pixels first
#endif
    const int my_rank = atomicAdd(&s_hist[bin], 1);
    __syncthreads();
    if (tid < 32) {  // exclusive scan of the 64 bins by one warp (2 per lane)
      const int a0 = s_hist[2 * tid], a1 = s_hist[2 * tid + 1];
      int inc = a0 + a1;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (tid >= d) inc += o;
      }
      s_hist[2 * tid] = inc - a0 - a1;
      s_hist[2 * tid + 1] = inc - a1;
    }
    __syncthreads();
    s_perm[s_hist[bin] + my_rank] = (unsigned char)tid;
    __syncthreads();
    mine = s_perm[tid];
    x = x0 + (mine & 31);
    y = y0 + (mine >> 5);
#pragma unroll
    for (int r = 0; r < SPAN; ++r) {
      const int2 run = s_runs[r][mine];
      rs[r] = run.x;
      rl[r] = run.y;
    }
#endif
  }
  const int ly = y - y0;  // local row of this thread's pixel

  // ---- the pixel this thread walks
  const bool inside = (x < p.W) && (y < p.H);
  PixelCtx c;
  c.xf = s_xf[x - x0];
  c.yf = s_yf[ly];
  c.r2 = p.r2;
  const int n_staged = s_staged;
  const bool staged = n_staged != 0;
  if (staged) {
    // wait for the bulk copies (phase 0 of the barrier)
    const uint32_t bar = smem_u32(&s_bar);
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred P1;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 0;\n\t"
          "selp.u32 %0, 1, 0, P1;\n\t}"
          : "=r"(done)
          : "r"(bar)
          : "memory");
    }
#pragma unroll
    for (int r = 0; r < SPAN; ++r) rs[r] += s_delta[ly + r];
    // z range of the tile -> keys that order exactly like z (see KeyCode)
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    for (int r = threadIdx.y; r < ROWS; r += kTileH) {  // one warp per staged row
      const int2 row = s_row[r];
      for (int i = threadIdx.x; i < row.y; i += 32) {
        const uint32_t zb = __float_as_uint(__fadd_rn(staged_rec.a(row.x + i).z, 0.0f));
        lo = min(lo, zb);
        hi = max(hi, zb);
      }
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((tid & 31) == 0) {
      atomicMin(&s_zlo, lo);
      atomicMax(&s_zhi, hi);
    }
  }
  __syncthreads();  // (uniform: `staged` is per CTA)
  // short lists: sorted 32-bit keys (KeyList); long lists: the general pair list
  Slots<KP> sl;
  bool amb;
  if constexpr (KP <= PGDVS_TILE_KEYS_MAXK) {
    KeyCode kc;  // the same code for every pixel of the tile
    kc.init(s_max, staged ? s_zlo : 0u, staged ? s_zhi : 0x7fffffffu);

    int total = 0;
  #pragma unroll
    for (int r = 0; r < SPAN; ++r) total += rl[r];

    KeyList<KP> q;
    q.init();
    if (HALO == 1) {
      // the three row runs are walked by ONE flattened loop so that lanes with uneven rows do not
      // wait for each other three times; the body is branch-free (see KeyList)
      const int c0 = rl[0], c01 = rl[0] + rl[1];
      const int s0 = rs[0], o1 = rs[1] - c0, o2 = rs[2] - c01;
      if (staged) {
        if (kc.sh == 0) {
          // exact keys: key = (zb - base) * 2^bits + t as ONE multiply-add (modulo 2^32)
          const uint32_t mul = 1u << kc.bits;
          uint32_t tk = 0u - kc.base * mul;
          auto key_at = [&](int t, uint32_t tkey) {
            const int j = t + (t < c0 ? s0 : (t < c01 ? o1 : o2));
            const float4 a = staged_rec.a(j);
            uint32_t key;  // one IMAD on the FMA pipe (C++ would turn it into shift + add on the ALU pipe)
            asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(key) : "r"(__float_as_uint(__fadd_rn(a.z, 0.0f))), "r"(mul), "r"(tkey));
            return hit_test<false>(c, a, nullptr, j) ? key : kEmpty;
          };
          int t = 0;
          if constexpr (KP >= 2 && KP <= PGDVS_RASTER_BRANCHFREE_MAXK) {
            for (; t + 1 < total; t += 2, tk += 2) q.insert2(key_at(t, tk), key_at(t + 1, tk + 1));
          }
          for (; t < total; ++t, ++tk) q.push(key_at(t, tk));
        } else {
          for (int t = 0; t < total; ++t) {
            const int j = t + (t < c0 ? s0 : (t < c01 ? o1 : o2));
            const float4 a = staged_rec.a(j);
            q.push(kc.encode(hit_test<false>(c, a, nullptr, j), a.z, (uint32_t)t));
          }
        }
      } else {
        // software-pipelined global reads: record t+1 is in flight while t is processed
        int j = (0 < c0 ? s0 : (0 < c01 ? o1 : o2));
        float4 a = (total > 0) ? __ldg(p.recA + rec_a(j)) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = 0; t < total; ++t) {
          const int tn = t + 1;
          const int jn = tn + (tn < c0 ? s0 : (tn < c01 ? o1 : o2));
          float4 an = a;
          if (tn < total) an = __ldg(p.recA + rec_a(jn));
          q.push(kc.encode(hit_test<false>(c, a, nullptr, j), a.z, (uint32_t)t));
          a = an;
          j = jn;
        }
      }
    } else if (staged) {
      // wider windows: the same flattened walk, the run of ordinal t found with SPAN-1 selects
      int cum[SPAN], off[SPAN];  // cum[r] = candidates in rows 0..r, off[r] = rs[r] - cum[r-1]
      int acc = 0;
  #pragma unroll
      for (int r = 0; r < SPAN; ++r) {
        off[r] = rs[r] - acc;
        acc += rl[r];
        cum[r] = acc;
      }
      auto key_at = [&](int t) {
        int o = off[SPAN - 1];
  #pragma unroll
        for (int r = SPAN - 2; r >= 0; --r) o = (t < cum[r]) ? off[r] : o;
        const int j = t + o;
        const float4 a = staged_rec.a(j);
        return kc.encode(hit_test<false>(c, a, nullptr, j), a.z, (uint32_t)t);
      };
      int t = 0;
      if constexpr (KP >= 2 && KP <= PGDVS_RASTER_BRANCHFREE_MAXK) {
        for (; t + 1 < total; t += 2) q.insert2(key_at(t), key_at(t + 1));
      }
      for (; t < total; ++t) q.push(key_at(t));
    } else {
      // unstaged tile with a wide window: row by row over the global records
      uint32_t t = 0;
  #pragma unroll
      for (int r = 0; r < SPAN; ++r) {
        const int s = rs[r], e = rs[r] + rl[r];
        for (int j = s; j < e; ++j, ++t) {
          const float4 a = __ldg(p.recA + rec_a(j));
          q.push(kc.encode(hit_test<false>(c, a, nullptr, j), a.z, t));
        }
      }
    }
    // ordinal -> record slot (shared-memory slot when staged: rs[] already carries s_delta)
    if (HALO == 1) {
      const int c0 = rl[0], c01 = rl[0] + rl[1];
      const int s0 = rs[0], o1 = rs[1] - c0, o2 = rs[2] - c01;
  #pragma unroll
      for (int i = 0; i < KP; ++i) {
        const int o = (int)(q.k[i] & kc.mask);
        sl.s[i] = (q.k[i] != kEmpty) ? o + (o < c0 ? s0 : (o < c01 ? o1 : o2)) : -1;
      }
    } else {
  #pragma unroll
      for (int i = 0; i < KP; ++i) {
        int o = (int)(q.k[i] & kc.mask);
        int j = -1;
  #pragma unroll
        for (int r = 0; r < SPAN; ++r) {
          if (j < 0 && o < rl[r]) j = rs[r] + o;
          o -= rl[r];
        }
        sl.s[i] = (q.k[i] != kEmpty) ? j : -1;
      }
    }
    amb = inside && ((KP <= 32 && p.K == KP) ? q.ambiguous_full(kc) : q.ambiguous(p.K, kc));
  } else {
    static_assert(KP <= PGDVS_TILE_KEYS_MAXK, "longer lists use k_raster_cells");
  }
  if (!staged) {  // (uniform) slots are global indices: finish in walk order
    if (inside) finish_global<KP, false>(p, c, n, x, y, sl, amb);
    return;
  }
#ifndef PGDVS_RASTER_NO_SORT
  // ---- hand the winners back to the pixel's own thread: the epilogue then runs in raster order,
  //      so fragment / image / mask stores and the static-frame loads are coalesced again and
  //      neighbouring lanes mostly read the same staged records
  constexpr uint16_t kNone = 0xFFFFu, kDone = 0xFFFEu;  // staged slots are < smem_records < 0xFFFE
  if (amb) finish_global<KP, false>(p, c, n, x, y, sl, true);  // rare: the walker finishes it
  {
    uint16_t* dst = s_win + mine * KP;
    if (KP % 8 == 0) {
#pragma unroll
      for (int i = 0; i < KP; i += 8) {
        uint32_t w[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const uint32_t lo = (sl.s[i + 2 * h] < 0) ? kNone : (uint32_t)sl.s[i + 2 * h];
          const uint32_t hi = (sl.s[i + 2 * h + 1] < 0) ? kNone : (uint32_t)sl.s[i + 2 * h + 1];
          w[h] = lo | (hi << 16);
        }
        if (i == 0 && amb) w[0] = (w[0] & 0xFFFF0000u) | kDone;
        *reinterpret_cast<uint4*>(dst + i) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < KP; ++i) dst[i] = (sl.s[i] < 0) ? kNone : (uint16_t)sl.s[i];
      if (amb) dst[0] = kDone;
    }
  }
  __syncthreads();
  x = x0 + threadIdx.x;
  y = y0 + threadIdx.y;
  if (x >= p.W || y >= p.H) return;
  {
    const uint16_t* src = s_win + tid * KP;
    if (KP % 8 == 0) {
#pragma unroll
      for (int i = 0; i < KP; i += 8) {
        const uint4 v = *reinterpret_cast<const uint4*>(src + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const uint32_t lo = w[h] & 0xFFFFu, hi = w[h] >> 16;
          sl.s[i + 2 * h] = (lo >= kDone) ? -1 - (int)(kNone - lo) : (int)lo;  // kNone -> -1, kDone -> -2
          sl.s[i + 2 * h + 1] = (hi == kNone) ? -1 : (int)hi;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < KP; ++i) sl.s[i] = (src[i] >= kDone) ? -1 - (int)(kNone - src[i]) : (int)src[i];
    }
  }
  if (sl.s[0] == -2) return;  // finished by its walker
  c.xf = s_xf[threadIdx.x];
  c.yf = s_yf[threadIdx.y];
#else
  if (!inside) return;
  if (amb) {
    finish_global<KP, false>(p, c, n, x, y, sl, true);
    return;
  }
#endif
  if (KP <= 32 && p.K == KP) {
#ifndef PGDVS_RASTER_NO_SPEC
    if (KP >= 8 && KP <= 16 && p.compositor == PGDVS_COMPOSITE_NORM_WEIGHTED && p.C == 3)
      pixel_epilogue<KP, true, StagedRecords, true>(p, sl, c, n, x, y, staged_rec);
    else
#endif
      pixel_epilogue<KP, true>(p, sl, c, n, x, y, staged_rec);
  } else {
    pixel_epilogue<KP, false>(p, sl, c, n, x, y, staged_rec);
  }
}

// ---------------------------------------------------------------------------------------
// In-place z-sort of the small cells (generic path only).  One thread per cell: the cell's
// records are read into registers, ranked by their z bit pattern (z >= 0, so the pattern orders
// like the value; NaN patterns sort last; equal z keep their order) and written back to their
// ranks.  Cells are disjoint, so no thread touches another's records.  Any order of the records
// of a cell gives the same fragments (ties are resolved by packed index, not by position), so a
// sorted workspace is as good as an unsorted one for every other consumer.
// ---------------------------------------------------------------------------------------
int maybe_sort_cells(RasterParams& p, int KP, cudaStream_t stream);

#if !defined(PGDVS_RASTER_PART) || PGDVS_RASTER_PART == 0
__global__ void __launch_bounds__(128) k_sort_cells(const int* __restrict__ cell_end, float4* rec,
                                                    int64_t n_cells) {
  const int64_t cell = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (cell >= n_cells) return;
  const int s = __ldg(cell_end + cell - 1), e = __ldg(cell_end + cell);
  const int n = e - s;
  if (n < 2 || n > kSortCap) return;
  float4 a[kSortCap], b[kSortCap];
  uint32_t key[kSortCap];
#pragma unroll
  for (int i = 0; i < kSortCap; ++i) {
    key[i] = 0xFFFFFFFFu;
    if (i < n) {
      a[i] = rec[rec_a(s + i)];
      b[i] = rec[rec_b(s + i)];
      key[i] = __float_as_uint(__fadd_rn(a[i].z, 0.0f));
    }
  }
#pragma unroll
  for (int i = 0; i < kSortCap; ++i) {
    if (i < n) {
      int r = 0;  // records strictly before record i in (z pattern, position) order
#pragma unroll
      for (int j = 0; j < kSortCap; ++j) {
        if (j < i) r += (key[j] <= key[i]) ? 1 : 0;
        if (j > i) r += (j < n && key[j] < key[i]) ? 1 : 0;
      }
      if (r != i) {
        rec[rec_a(s + r)] = a[i];
        rec[rec_b(s + r)] = b[i];
      }
    }
  }
}

int maybe_sort_cells(RasterParams& p, int KP, cudaStream_t stream) {
  const double span = 2.0 * p.halo + 1.0;
  const double candidates = span * span * p.density;  // mean candidates per pixel
  bool sort = candidates >= 64.0 && candidates >= 4.0 * KP;
  if (const char* env = getenv("PGDVS_SORT_CELLS")) sort = env[0] == '1';  // developer switch
  p.cells_sorted = 0;
  if (!sort) return 0;
  const int64_t n_cells = (int64_t)p.N * p.GH * p.GW;
  const int64_t blocks = (n_cells + 127) / 128;
  k_sort_cells<<<(unsigned)blocks, 128, 0, stream>>>(p.cell_end, const_cast<float4*>(p.recA), n_cells);
  if (int rc = check_launch()) return rc;
  p.cells_sorted = 1;
  return 0;
}
#endif

#ifndef PGDVS_RASTER_SMEM_BYTES
#define PGDVS_RASTER_SMEM_BYTES (24 * 1024)
#endif
#ifndef PGDVS_RASTER_SMEM_BYTES_WIDE
#define PGDVS_RASTER_SMEM_BYTES_WIDE (100 * 1024)
#endif

template <int KP, int HALO>
static bool launch_tile(RasterParams& p, dim3 grid, dim3 block, double density, cudaStream_t stream) {
  // size the staging buffer from the mean point density (points per pixel of the batch) with
  // 1.5x head-room; tiles that still overflow fall back to global reads inside the kernel.
  // If even the mean tile would not fit in PGDVS_RASTER_SMEM_BYTES_WIDE, the generic kernel
  // (more resident CTAs, no staging) is the better choice.
  const double tile_cells = (double)(kTileW + 2 * HALO) * (kTileH + 2 * HALO);
#ifndef PGDVS_RASTER_HEADROOM
#define PGDVS_RASTER_HEADROOM 1.5
#endif
#ifndef PGDVS_RASTER_HEADROOM_DENSE
#define PGDVS_RASTER_HEADROOM_DENSE 1.15
#endif
  double need = PGDVS_RASTER_HEADROOM * density * tile_cells * 32.0;
  if (need > (double)PGDVS_RASTER_SMEM_BYTES_WIDE) {
    // dense clouds: trade head-room for staging at all (tiles that still overflow walk the
    // global records inside the kernel)
    need = PGDVS_RASTER_HEADROOM_DENSE * density * tile_cells * 32.0;
    if (need > (double)PGDVS_RASTER_SMEM_BYTES_WIDE) return false;
  }
  int smem = (int)need;
  if (smem < 24 * 1024) smem = 24 * 1024;
  smem = (smem + 1023) & ~1023;
  if (HALO == 1 && smem < PGDVS_RASTER_SMEM_BYTES) smem = PGDVS_RASTER_SMEM_BYTES;
  p.smem_records = smem / 32;
  static int attr_smem = 0;  // per instantiation: raise the opt-in limit when needed
  if (smem > attr_smem) {
    cudaFuncSetAttribute(k_raster_tile<KP, HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr_smem = smem;
  }
  k_raster_tile<KP, HALO><<<grid, block, smem, stream>>>(p);
  return true;
}

template <int KP>
int launch_raster(RasterParams& p, cudaStream_t stream) {
  dim3 block(32, 8);
  dim3 grid((p.W + 31) / 32, (p.H + 7) / 8, p.N);
  bool done = false;
#ifndef PGDVS_RASTER_NO_TMA
  if constexpr (KP <= 32) {
    const double density = p.density;
    if (p.r2 < 0.0f || p.force_generic)
      done = false;  // per-point radii: generic kernel
    else if (p.halo == 1)
      done = launch_tile<KP, 1>(p, grid, block, density, stream);
    else if (p.halo == 2)
      done = launch_tile<KP, 2>(p, grid, block, density, stream);
    else if (p.halo == 3)
      done = launch_tile<KP, 3>(p, grid, block, density, stream);
  }
#endif
  if (!done) {
    // many more candidates per pixel than kept hits: z-sort the cells first (see kSortCap)
    if (int rc = maybe_sort_cells(p, KP, stream)) return rc;
    if (p.r2 < 0.0f)
      k_raster_cells<KP, true><<<grid, block, 0, stream>>>(p);
    else
      k_raster_cells<KP, false><<<grid, block, 0, stream>>>(p);
  }
  return check_launch();
}

// The file can be compiled as one translation unit (default) or, to build the K
// instantiations in parallel, several times with -DPGDVS_RASTER_PART=n (see _build.py):
// part 0 holds the C entry point, parts 1..6 one group of launch_raster<KP> each.
#define PGDVS_RASTER_FOR_PART(n, X) PGDVS_RASTER_FOR_PART_I(n, X)
#define PGDVS_RASTER_FOR_PART_I(n, X) PGDVS_RASTER_FOR_PART_##n(X)
#define PGDVS_RASTER_FOR_PART_1(X) X(1) X(2) X(3) X(4)
#define PGDVS_RASTER_FOR_PART_2(X) X(8)
#define PGDVS_RASTER_FOR_PART_3(X) X(16)
#define PGDVS_RASTER_FOR_PART_4(X) X(32)
#define PGDVS_RASTER_FOR_PART_5(X) X(64)
#define PGDVS_RASTER_FOR_PART_6(X) X(PGDVS_MAX_POINTS_PER_PIXEL)
#define PGDVS_RASTER_DECLARE(KP) extern template int launch_raster<KP>(RasterParams&, cudaStream_t);
#define PGDVS_RASTER_DEFINE(KP) template int launch_raster<KP>(RasterParams&, cudaStream_t);
#if defined(PGDVS_RASTER_PART) && PGDVS_RASTER_PART == 0
PGDVS_RASTER_FOR_PART(1, PGDVS_RASTER_DECLARE)
PGDVS_RASTER_FOR_PART(2, PGDVS_RASTER_DECLARE)
PGDVS_RASTER_FOR_PART(3, PGDVS_RASTER_DECLARE)
PGDVS_RASTER_FOR_PART(4, PGDVS_RASTER_DECLARE)
PGDVS_RASTER_FOR_PART(5, PGDVS_RASTER_DECLARE)
PGDVS_RASTER_FOR_PART(6, PGDVS_RASTER_DECLARE)
#elif defined(PGDVS_RASTER_PART)
PGDVS_RASTER_FOR_PART(PGDVS_RASTER_PART, PGDVS_RASTER_DEFINE)
#endif

}  // namespace pgdvs

#if !defined(PGDVS_RASTER_PART) || PGDVS_RASTER_PART == 0
using namespace pgdvs;

extern "C" int pgdvs_rasterize_composite(void* workspace, size_t workspace_bytes, int N,
                                         int64_t P, int H, int W, int K, float radius_max,
                                         int per_point_radius, int C, int compositor,
                                         float rr_weight, const float* background,
                                         const float* static_rgb, int32_t* idx, float* zbuf,
                                         float* dists, float* image, float* mask, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (workspace == nullptr || N < 0 || H <= 0 || W <= 0 || P < 0 || K < 1 ||
      !(radius_max >= 0.0f))
    return PGDVS_E_BADARG;
  if (K > PGDVS_MAX_POINTS_PER_PIXEL) return PGDVS_E_K_TOO_LARGE;
  if (P >= kMaxRecords) return PGDVS_E_BADARG;
  if (compositor < PGDVS_COMPOSITE_NONE || compositor > PGDVS_COMPOSITE_WEIGHTED_SUM)
    return PGDVS_E_BADARG;
  if (compositor != PGDVS_COMPOSITE_NONE) {
    if (C < 1 || C > PGDVS_MAX_FUSED_CHANNELS) return PGDVS_E_CHANNELS;
    if (!(rr_weight > 0.0f)) return PGDVS_E_BADARG;
    if (static_rgb != nullptr && (image == nullptr)) return PGDVS_E_BADARG;
  }
  BinLayout L = make_bin_layout(N, H, W, P, radius_max);
  if (workspace_bytes < L.total) return PGDVS_E_WORKSPACE;
  if (N == 0) return PGDVS_OK;

  const char* ws = static_cast<const char*>(workspace);
  RasterParams p;
  p.cell_end = reinterpret_cast<const int*>(ws + L.off_cells);
  p.recA = reinterpret_cast<const float4*>(ws + L.off_recA);
  p.N = N;
  p.H = H;
  p.W = W;
  p.K = K;
  p.C = (compositor == PGDVS_COMPOSITE_NONE) ? 0 : C;
  p.halo = L.halo;
  p.GW = L.GW;
  p.GH = L.GH;
  p.ax = make_ndc_axis(W, H);
  p.ay = make_ndc_axis(H, W);
  // fp32 r*r, as `radius2 = radius * radius` upstream; negative selects the per-point path
  p.r2 = per_point_radius ? -1.0f : radius_max * radius_max;
  p.rr_weight = rr_weight;
  p.inv_rr_weight = 1.0f / rr_weight;
  p.compositor = compositor;
  for (int c = 0; c < 4; ++c) p.bg[c] = (background != nullptr && c < C) ? background[c] : 0.0f;
  p.static_rgb = static_rgb;
  p.idx = idx;
  p.zbuf = zbuf;
  p.dists = dists;
  p.image = image;
  p.mask = mask;
  p.smem_records = 0;
  p.cells_sorted = 0;
  {
    const char* env = getenv("PGDVS_RASTER_FORCE_GENERIC");  // developer switch
    p.force_generic = (env != nullptr && env[0] == '1') ? 1 : 0;
  }
  // mean points per pixel of the batch (P is the capacity of the packed cloud: an upper bound)
  p.density = (double)P / ((double)N * (double)H * (double)W);

#ifdef PGDVS_RASTER_EXP_K8  // developer experiments (tools/exp_variants.py): one instantiation
  if (K > 4 && K <= 8) return launch_raster<8>(p, stream);
  return PGDVS_E_BADARG;
#else
  if (K <= 1) return launch_raster<1>(p, stream);
  if (K <= 2) return launch_raster<2>(p, stream);
  if (K <= 3) return launch_raster<3>(p, stream);
  if (K <= 4) return launch_raster<4>(p, stream);
  if (K <= 8) return launch_raster<8>(p, stream);
  if (K <= 16) return launch_raster<16>(p, stream);
  if (K <= 32) return launch_raster<32>(p, stream);
  if (K <= 64) return launch_raster<64>(p, stream);
  return launch_raster<PGDVS_MAX_POINTS_PER_PIXEL>(p, stream);
#endif
}
#endif  // part 0 / single translation unit
