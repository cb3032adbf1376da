"""CPU models of the round-2 (session 3) selection algorithms, checked against plain sorting:

* sort4 / merge4 / assign4 of ml-pgdvs_b200/csrc/raster.cu (the batched walks of k_raster_pair and
  k_raster_tile): after any sequence of four-key batches the list holds the KP smallest keys seen,
  in order, and `rej` is the smallest key that ever fell off — the two facts the ambiguity test
  (KeyList::ambiguous_full) relies on;
* the K-nearest statistic of k_knn_query_warp (csrc/knn_grid.cu): K-th smallest by bisection on the
  bit pattern, mean = (sum below the K-th + ties * K-th - the `skip` smallest) / (K - skip), against
  the mean over the sorted prefix that the thread kernel and the reference (torch.mean over
  knn_points' dists) form;
* the block -> (job, source tile) map of k_fill_pre_tiles (csrc/bin.cu): a bijection that visits the
  same tile of all jobs of a view back to back.

No GPU, no product code: these restate the device logic in Python integers."""
import numpy as np
import pytest

EMPTY = 0xFFFFFFFF


def cex(k, i, j):
    lo, hi = min(k[i], k[j]), max(k[i], k[j])
    k[i], k[j] = lo, hi


def sort4(b):
    cex(b, 0, 1); cex(b, 2, 3); cex(b, 0, 2); cex(b, 1, 3); cex(b, 1, 2)
    return b


class BatchList:
    def __init__(self, KP):
        self.KP, self.k, self.rej = KP, [EMPTY] * KP, EMPTY

    def assign4(self, b):
        self.k[:4] = b

    def merge4(self, b):  # b ascending
        k, KP = self.k, self.KP
        r = [max(k[KP - 4 + i], b[3 - i]) for i in range(4)]
        for i in range(4):
            k[KP - 4 + i] = min(k[KP - 4 + i], b[3 - i])
        self.rej = min([self.rej] + r)
        stride = KP // 2
        while stride >= 1:
            for i in range(KP):
                if (i & stride) == 0:
                    cex(k, i, i + stride)
            stride //= 2


@pytest.mark.parametrize("KP", [4, 8, 16])
@pytest.mark.parametrize("seed", range(6))
def test_sort4_merge4_keep_the_smallest_in_order(KP, seed):
    rng = np.random.default_rng(100 * KP + seed)
    for trial in range(300):
        n_batches = int(rng.integers(1, 12))
        # few distinct values -> many duplicates; a share of misses (EMPTY), as in the walk
        universe = int(rng.choice([6, 40, 1 << 20]))
        q, seen = BatchList(KP), []
        for bi in range(n_batches):
            b = [EMPTY if rng.random() < 0.3 else int(rng.integers(0, universe)) for _ in range(4)]
            seen += b
            sort4(b)
            assert b == sorted(b)
            if bi == 0 and rng.random() < 0.5:
                q.assign4(b)  # the first batch of a walk becomes the list
            else:
                q.merge4(b)
            want = sorted(seen)
            want_k = (want + [EMPTY] * KP)[:KP]
            assert q.k == want_k, (trial, bi)
            dropped = want[KP:]
            assert q.rej == (dropped[0] if dropped else EMPTY), (trial, bi)


def _bits(x):
    return np.float32(x).view(np.uint32)


def knn_stat_bisect(d2, K, skip):
    """k_knn_query_warp: kk-th smallest by bisection over the patterns, then the closed-form mean."""
    d2 = np.asarray(d2, np.float32)
    kk = min(K, d2.size)
    if kk <= skip:
        return np.float32(0.0)
    pat = d2.view(np.uint32)
    lo, hi = 0, 0x7F800000
    while lo < hi:
        mid = lo + ((hi - lo) >> 1)
        if int((pat <= mid).sum()) >= kk:
            hi = mid
        else:
            lo = mid + 1
    kth = np.uint32(lo).view(np.float32)
    below = d2[d2 < kth]
    s = np.float32(below.astype(np.float64).sum()) + np.float32(kk - below.size) * kth
    if skip == 1:
        s = s - d2.min()
    return np.float32(s / np.float32(kk - skip))


@pytest.mark.parametrize("seed", range(8))
def test_knn_statistic_by_bisection_equals_sorted_prefix_mean(seed):
    rng = np.random.default_rng(seed)
    for trial in range(200):
        n = int(rng.integers(1, 400))
        d2 = (rng.random(n) ** 2 * 10.0).astype(np.float32)
        if rng.random() < 0.5:  # exact duplicates around the K-th, and a zero self-distance
            d2[rng.integers(0, n, n // 3)] = d2[0]
            d2[rng.integers(0, n)] = 0.0
        for K, skip in ((51, 1), (51, 0), (7, 0), (1, 0)):
            got = knn_stat_bisect(d2, K, skip)
            kk = min(K, n)
            srt = np.sort(d2)
            want = srt[skip:kk].astype(np.float64).mean() if kk > skip else 0.0
            assert abs(float(got) - want) <= 1e-5 * max(abs(want), 1e-12) + 1e-12, (trial, K, skip)


def test_tile_major_fill_order_is_a_bijection_grouped_by_view():
    rng = np.random.default_rng(3)
    for trial in range(50):
        n_views = int(rng.integers(1, 6))
        per_view = rng.integers(1, 9, n_views)
        views = np.repeat(np.arange(n_views), per_view)  # jobs sorted by view
        n_jobs, T = views.size, int(rng.integers(1, 7))
        seen, order = set(), []
        for o in range(n_jobs * T):
            jb = o // T
            v = views[jb]
            f = jb
            while f > 0 and views[f - 1] == v:
                f -= 1
            l = jb
            while l + 1 < n_jobs and views[l + 1] == v:
                l += 1
            c, local = l - f + 1, o - f * T
            jt, j = local // c, f + local % c
            assert 0 <= jt < T and f <= j <= l
            seen.add((j, jt))
            order.append((int(views[j]), jt, j))
        assert len(seen) == n_jobs * T  # every (job, tile) exactly once
        assert order == sorted(order)   # view-major, then tile, then job: a tile of all jobs of a view back to back
