#!/usr/bin/env python
"""Developer experiment: the same prepared views rendered through several builds of the library
(in-tree first, then exp_libs/lib_<name>.so for every name in LIBS=a,b) — CUDA-event time of the
rasterize-and-composite call and a bit-for-bit comparison against the in-tree build.  Not part of
the product.

    LIBS=occA,occB CASES="C5 K=8,C5 K=32" python tools/exp_libs_ab.py     (on the GPU box)

A candidate build is made here with e.g.
    PGDVS_NVCC_EXTRA=-DPGDVS_TILE_FIXED_BINS python -m pgdvs_b200._build --force
    cp ml-pgdvs_b200/lib/libpgdvs_b200.so exp_libs/lib_fixedbins.so
followed by a forced rebuild without the flag (exp_libs/ travels to the GPU box, is git-ignored).
"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from pgdvs_b200 import _cabi, synthetic  # noqa: E402
from pgdvs_b200.dyn_renderer import prepare_views, render_prepared  # noqa: E402

CASES = [
    ("C2 K=8", "c2_nvidia_seq", dict()),
    ("C4 K=8", "c4_davis", dict(n_views=16)),
    ("C5 K=8 r=0.01", "c5_stress", dict(K=8, radius=0.01, n_views=4)),
    ("C5 K=16 r=0.005", "c5_stress", dict(K=16, radius=0.005, n_views=4)),
    ("C5 K=32 r=0.02", "c5_stress", dict(K=32, radius=0.02, n_views=2)),
    ("C3 K=16", "c3_iphone", dict(n_views=8)),
]


def main():
    dev = torch.device("cuda:0")
    in_tree = _cabi.LIB_PATH
    libs = [("in-tree", in_tree)] + [(n, ROOT / "exp_libs" / f"lib_{n}.so") for n in os.environ.get("LIBS", "").split(",") if n]
    only = os.environ.get("CASES")
    print("| case | build | raster ms | same bits as in-tree |")
    print("|---|---|---|---|")
    for label, name, kw in CASES:
        if only and not any(o in label for o in only.split(",")):
            continue
        wl = synthetic.make_workload(name, dev, **kw)
        pairs, cams = wl.jobs(range(wl.n_views))
        prep = prepare_views(pairs, cams, wl.H, wl.W, dev)
        ref = None
        for lib_name, path in libs:
            _cabi.LIB_PATH, _cabi._lib = path, None  # next _cabi.lib() loads this build

            def step(ev=None):
                return render_prepared(prep, radius=wl.radius, points_per_pixel=wl.K, compositor="norm",
                                       static_rgb=wl.static_rgb, return_fragments=True, raster_events=ev)
            out = step()
            torch.cuda.synchronize()
            n = 3
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
            for i in range(n):
                out = step(evs[i])
            torch.cuda.synchronize()
            r_ms = sum(a.elapsed_time(b) for a, b in evs) / n
            keys = ("idx", "zbuf", "dists", "image", "mask")
            if ref is None:
                ref = {k: out[k].clone() for k in keys}
                same = "-"
            else:
                same = "yes" if all(torch.equal(ref[k], out[k]) for k in keys) else \
                    "NO: " + ",".join(k for k in keys if not torch.equal(ref[k], out[k]))
            print(f"| {label} ({wl.n_views} views) | {lib_name} | {r_ms:.3f} | {same} |", flush=True)
            del out
        del ref, prep, wl
        torch.cuda.empty_cache()
    _cabi.LIB_PATH, _cabi._lib = in_tree, None


if __name__ == "__main__":
    main()
