"""CPU model of the tile rasterizer's K-list (ml-pgdvs_b200/csrc/raster.cu: KeyCode, KeyList,
ambiguous / ambiguous_full) checked against the exact (z, idx) order of the CPU rasterizer's
priority queue (oracle/raster_cpu.cpp, after pytorch3d RasterizePointsNaiveCpu).

The kernel's exactness argument is algorithmic: one 32-bit key per slot, (z pattern - base) >> sh
in the high bits and the candidate's ordinal in the low bits; min/max insertion without payload;
every case in which key order could differ from (z, idx) order must raise `ambiguous` (those
pixels are rescanned exactly).  This file restates that logic in Python integers and checks the
claim on random candidate streams: exact ties, truncated keys (sh > 0), -0.0, lists that never
fill, K < KP.  It runs without a GPU and does not touch the product path."""
import numpy as np
import pytest

EMPTY = 0xFFFFFFFF
M32 = 0xFFFFFFFF


def clz32(x):
    return 32 - int(x).bit_length() if x else 32


class KeyCode:
    """KeyCode::init / encode / same_z."""

    def __init__(self, n_candidates, zlo, zhi):
        self.bits = 32 - clz32(max(n_candidates - 1, 0))
        self.mask = (1 << self.bits) - 1
        self.base = zlo
        rng = (zhi - zlo) & M32
        self.sh = max(0, self.bits - clz32(rng))
        if ((((rng >> self.sh) << self.bits) & M32) | self.mask) == EMPTY:
            self.sh += 1

    def encode(self, hit, z, t):
        zb = int(np.float32(np.float32(z) + np.float32(0.0)).view(np.uint32))  # -0.0 -> +0.0
        return (((((zb - self.base) & M32) >> self.sh) << self.bits) & M32) | t if hit else EMPTY

    def same_z(self, a, b):
        return ((a ^ b) & ~self.mask & M32) == 0


class KeyList:
    def __init__(self, KP):
        self.KP = KP
        self.k = [EMPTY] * KP
        self.rej = EMPTY

    def insert(self, c):
        k, KP = self.k, self.KP
        prev = k[0]
        k[0] = min(c, prev)
        for i in range(1, KP):
            cur = k[i]
            k[i] = min(max(c, prev), cur)
            prev = cur
        self.rej = min(self.rej, max(c, prev))

    def insert2(self, c1, c2):
        k, KP = self.k, self.KP
        lo, hi = min(c1, c2), max(c1, c2)
        p2, p1 = k[0], k[1]
        k[0] = min(p2, lo)
        k[1] = min(min(p1, max(p2, lo)), hi)
        for i in range(2, KP):
            cur = k[i]
            k[i] = min(min(cur, max(p1, lo)), max(p2, hi))
            p2, p1 = p1, cur
        self.rej = min(min(self.rej, max(p1, lo)), max(p2, hi))

    def push(self, c, branchfree):
        if branchfree or c < self.k[self.KP - 1]:
            self.insert(c)
        else:
            self.rej = min(self.rej, c)

    def ambiguous_full(self, kc):
        amb = self.rej != EMPTY and kc.same_z(self.rej, self.k[self.KP - 1])
        for i in range(1, self.KP):
            amb = amb or (self.k[i] != EMPTY and kc.same_z(self.k[i], self.k[i - 1]))
        return amb

    def ambiguous(self, K, kc):
        amb = False
        for i in range(1, self.KP):
            if i <= K:
                amb = amb or (self.k[i] != EMPTY and kc.same_z(self.k[i], self.k[i - 1]))
        if K >= self.KP:
            amb = amb or (self.rej != EMPTY and kc.same_z(self.rej, self.k[self.KP - 1]))
        return amb


def _zbits(z):
    return int(np.float32(np.float32(z) + np.float32(0.0)).view(np.uint32))


def _walk(z, hit, KP, K, kc, pairs):
    """The tile kernel's walk over one pixel's candidates (ordinals 0..n-1): insert2 on pairs for
    the branch-free sizes, push for the tail / the larger lists."""
    q = KeyList(KP)
    n = len(z)
    t = 0
    if pairs and KP >= 2:
        while t + 1 < n:
            q.insert2(kc.encode(hit[t], z[t], t), kc.encode(hit[t + 1], z[t + 1], t + 1))
            t += 2
    while t < n:
        q.push(kc.encode(hit[t], z[t], t), branchfree=pairs)
        t += 1
    amb = q.ambiguous_full(kc) if K == KP else q.ambiguous(K, kc)
    return [(k & kc.mask) if k != EMPTY else -1 for k in q.k[:K]], amb


def _exact(z, hit, idx, K):
    """K nearest hits in the (z, idx) order of the reference's priority queue."""
    order = sorted((t for t in range(len(z)) if hit[t]), key=lambda t: (float(np.float32(z[t]) + np.float32(0.0)), idx[t]))
    return (order + [-1] * K)[:K]


@pytest.mark.parametrize("KP,K,pairs", [(8, 8, True), (8, 5, True), (4, 4, True), (2, 2, True), (1, 1, True),
                                        (16, 16, False), (32, 32, False), (16, 11, False)])
@pytest.mark.parametrize("regime", ["continuous", "ties", "wide_range", "sparse"])
def test_keys_order_like_the_reference_unless_flagged(KP, K, pairs, regime):
    rng = np.random.default_rng(KP * 100 + K + len(regime))
    flagged = clean = 0
    for trial in range(150):
        n = int(rng.integers(0, 70))
        if regime == "continuous":
            z = rng.uniform(1.0, 10.0, n).astype(np.float32)
        elif regime == "ties":
            z = (np.round(rng.uniform(0.0, 4.0, n) * 4) / 4).astype(np.float32)
            z[rng.random(n) < 0.1] = np.float32(-0.0)
        elif regime == "wide_range":  # more z patterns than fit beside the ordinal: sh > 0
            z = (10.0 ** rng.uniform(-6, 6, n)).astype(np.float32)
            if n > 3:
                z[1] = np.nextafter(z[0], np.float32(np.inf))  # one ulp apart: equal after truncation
                z[2], z[3] = np.float32(1e-6), np.float32(1e6)  # ~2^28.3 patterns between them
        else:
            z = rng.uniform(1.0, 2.0, n).astype(np.float32)
        hit = rng.random(n) < (0.15 if regime == "sparse" else 0.7)
        idx = rng.permutation(100000)[:n]  # packed indices are unrelated to the walk order
        # the tile's staged records are a superset of the pixel's candidates
        extra = rng.uniform(0.5, 12.0, 5).astype(np.float32) if regime != "wide_range" else z[:0]
        zb = [_zbits(v) for v in np.concatenate([z, extra])] or [0]
        n_max = max(n, int(rng.integers(n, 80)) if n else 1)  # s_max: most candidates of any pixel in the tile
        kc = KeyCode(n_max, min(zb), max(zb))
        if regime == "wide_range" and n > 3 and kc.bits >= 5:
            assert kc.sh > 0
        got, amb = _walk(z, hit, KP, K, kc, pairs)
        if amb:
            flagged += 1
            continue  # the kernel redoes these pixels with rescan_exact
        clean += 1
        want = _exact(z, hit, idx, K)
        assert got == want, f"trial {trial}: unflagged pixel differs: {got} vs {want} (sh={kc.sh}, bits={kc.bits})"
    assert clean > 0
    if regime == "ties":
        assert flagged > 0  # exact ties among the kept hits must be seen


def test_exact_keys_never_flag_distinct_depths():
    """sh == 0 and pairwise distinct z: the key order IS the z order and nothing is ambiguous."""
    rng = np.random.default_rng(3)
    for _ in range(200):
        n = int(rng.integers(1, 60))
        z = np.unique(rng.uniform(1.0, 8.0, n).astype(np.float32))
        rng.shuffle(z)
        hit = np.ones(len(z), bool)
        zb = [_zbits(v) for v in z]
        kc = KeyCode(len(z), min(zb), max(zb))
        assert kc.sh == 0
        got, amb = _walk(z, hit, 8, 8, kc, True)
        assert not amb
        assert got == _exact(z, hit, np.arange(len(z)), 8)


def test_empty_key_is_reserved():
    """No (z, ordinal) pair may encode to the all-ones key that stands for a miss."""
    for n, lo, hi in [(64, 0, 0x03FFFFFF), (33, 0, 0x7F800000), (2, 5, 0x7FFFFFFF), (1, 7, 7), (64, 0, 0xFFFFFFFF >> 6)]:
        kc = KeyCode(n, lo, hi)
        top = ((((hi - kc.base) & M32) >> kc.sh) << kc.bits) & M32 | kc.mask
        assert top != EMPTY, (n, lo, hi)


# ---------------------------------------------------------------------------------------------
# Generic kernel: (z, slot) pair list with tie flag (PairList) and the walk over z-sorted cells
# (k_sort_cells + the per-lane cursor loop of k_raster_cells).
# ---------------------------------------------------------------------------------------------
SORT_CAP = 16  # kSortCap
INF = np.float32(np.inf)


class PairList:
    def __init__(self, KP):
        self.KP = KP
        self.z = [INF] * KP
        self.s = [-1] * KP
        self.tie = False

    def push(self, hit, cz, cs):
        if not hit:
            return
        z, s, KP = self.z, self.s, self.KP
        if cz < z[KP - 1]:
            for i in range(KP):
                if cz < z[i]:
                    z[i], cz = cz, z[i]
                    s[i], cs = cs, s[i]
            self.tie = self.tie or (cs >= 0 and cz == z[KP - 1])  # (cz, cs) is now the evicted element
        else:
            self.tie = self.tie or (cz == z[KP - 1])

    def ambiguous(self, K):
        amb = K >= self.KP and self.tie
        for i in range(1, self.KP):
            if i <= K:
                amb = amb or (self.s[i] >= 0 and self.z[i] == self.z[i - 1])
        return amb


def _sort_cell(zs):
    """k_sort_cells: stable rank by the z bit pattern (after +0.0), cells above the cap untouched."""
    if len(zs) < 2 or len(zs) > SORT_CAP:
        return list(range(len(zs)))
    keys = [_zbits(v) for v in zs]
    return sorted(range(len(zs)), key=lambda i: (keys[i], i))


@pytest.mark.parametrize("KP,K", [(8, 8), (8, 6), (16, 16), (3, 3), (1, 1)])
@pytest.mark.parametrize("ties", [False, True])
def test_sorted_cell_walk_matches_the_reference_unless_flagged(KP, K, ties):
    rng = np.random.default_rng(KP * 10 + K + int(ties))
    flagged = clean = 0
    for trial in range(150):
        n_cells = int(rng.integers(1, 30))
        cells = []  # per cell: list of (z, hit, packed idx)
        next_idx = rng.permutation(5000)
        pos = 0
        for _ in range(n_cells):
            m = int(rng.choice([0, 1, 2, 5, 9, 16, 17, 24]))
            z = rng.uniform(0.0, 6.0, m).astype(np.float32)
            if ties:
                z = (np.round(z * 3) / 3).astype(np.float32)
                z[rng.random(m) < 0.05] = np.float32(-0.0)
            cells.append([(z[i], bool(rng.random() < 0.7), int(next_idx[pos + i])) for i in range(m)])
            pos += m
        # reference order over ALL records of the window
        flat = [rec for cell in cells for rec in cell]
        order = sorted((r for r in flat if r[1]), key=lambda r: (float(r[0] + np.float32(0.0)), r[2]))
        want = ([r[2] for r in order] + [-1] * K)[:K]
        # the kernel: cells sorted in place, cursor walk with the early exit
        q = PairList(KP)
        visited = 0
        for cell in cells:
            perm = _sort_cell([r[0] for r in cell])
            is_sorted = len(cell) <= SORT_CAP
            for i in perm:
                zc, hit, pidx = cell[i]
                visited += 1
                if zc <= q.z[KP - 1]:
                    q.push(hit, zc, pidx)  # the slot stands for the record; its packed idx identifies it
                elif is_sorted:
                    break
        if q.ambiguous(K):
            flagged += 1
            continue  # rescan_exact redoes the pixel
        clean += 1
        got = [q.s[i] if i < KP else -1 for i in range(K)]
        assert got == want, f"trial {trial}: {got} vs {want}"
        assert visited <= len(flat)
    assert clean > 0
    if ties and KP > 1:
        assert flagged > 0


def test_sort_cells_rank_formula_is_a_stable_permutation():
    """k_sort_cells ranks record i by  #{j < i: key_j <= key_i} + #{j > i: key_j < key_i}  — for ANY
    keys (duplicates, NaN patterns, all equal) that is the stable sort permutation, so no record
    is lost or duplicated when the records are written back to their ranks."""
    rng = np.random.default_rng(5)
    for _ in range(500):
        n = int(rng.integers(2, SORT_CAP + 1))
        keys = rng.choice([0, 1, 7, 0x3F800000, 0x7F800000, 0x7FC00000, 0xFFFFFFFF], n).tolist() \
            if rng.random() < 0.5 else rng.integers(0, 2 ** 32, n).tolist()
        rank = [sum(keys[j] <= keys[i] for j in range(i)) + sum(keys[j] < keys[i] for j in range(i + 1, n))
                for i in range(n)]
        assert sorted(rank) == list(range(n))
        out = [None] * n
        for i, r in enumerate(rank):
            out[r] = (keys[i], i)
        assert out == sorted((k, i) for i, k in enumerate(keys))
