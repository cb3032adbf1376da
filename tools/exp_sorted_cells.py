#!/usr/bin/env python
"""Developer experiment: the generic rasterizer with and without z-sorted cells (and, where the
tile kernel applies, against it), on the dense / wide-window BASELINE configs.  Every mode must
produce the same bits; the rasterize-and-composite launch (sort included) is timed with CUDA
events.  Not part of the product.

    python tools/exp_sorted_cells.py > gpurun_out/sorted_cells.md      (on the GPU box)
"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from pgdvs_b200 import synthetic  # noqa: E402
from pgdvs_b200.dyn_renderer import prepare_views, render_prepared  # noqa: E402

CASES = [
    ("C5 K=8 r=0.01", "c5_stress", dict(K=8, radius=0.01, n_views=4)),
    ("C5 K=16 r=0.005", "c5_stress", dict(K=16, radius=0.005, n_views=4)),
    ("C5 K=32 r=0.02", "c5_stress", dict(K=32, radius=0.02, n_views=2)),
    ("C3 K=16", "c3_iphone", dict(n_views=8)),
    ("C4 K=8", "c4_davis", dict(n_views=16)),
]
MODES = [  # name, PGDVS_SORT_CELLS, PGDVS_RASTER_FORCE_GENERIC
    ("default", None, None),
    ("generic unsorted", "0", "1"),
    ("generic sorted", "1", "1"),
]


def set_env(k, v):
    if v is None:
        os.environ.pop(k, None)
    else:
        os.environ[k] = v


def main():
    dev = torch.device("cuda:0")
    only = os.environ.get("CASES")
    print("| case | mode | raster ms | step ms | same bits as first mode |")
    print("|---|---|---|---|---|")
    for label, name, kw in CASES:
        if only and not any(o in label for o in only.split(",")):
            continue
        wl = synthetic.make_workload(name, dev, **kw)
        pairs, cams = wl.jobs(range(wl.n_views))
        prep = prepare_views(pairs, cams, wl.H, wl.W, dev)
        ref = None
        for mode, srt, gen in MODES:
            set_env("PGDVS_SORT_CELLS", srt)
            set_env("PGDVS_RASTER_FORCE_GENERIC", gen)

            def step(ev=None):
                return render_prepared(prep, radius=wl.radius, points_per_pixel=wl.K, compositor="norm",
                                       static_rgb=wl.static_rgb, return_fragments=True, raster_events=ev)
            out = step()
            torch.cuda.synchronize()
            n = 3
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for i in range(n):
                out = step(evs[i])
            t1.record()
            torch.cuda.synchronize()
            r_ms = sum(a.elapsed_time(b) for a, b in evs) / n
            s_ms = t0.elapsed_time(t1) / n
            keys = ("idx", "zbuf", "dists", "image", "mask")
            if ref is None:
                ref = {k: out[k].clone() for k in keys}
                same = "-"
            else:
                same = "yes" if all(torch.equal(ref[k], out[k]) for k in keys) else \
                    "NO: " + ",".join(k for k in keys if not torch.equal(ref[k], out[k]))
            print(f"| {label} ({wl.n_views} views) | {mode} | {r_ms:.3f} | {s_ms:.3f} | {same} |", flush=True)
            if os.environ.get("PROFILE") and mode == "default":
                # per-kernel device times of one step (CUPTI through torch.profiler; not a bench number)
                from torch.profiler import ProfilerActivity, profile
                with profile(activities=[ProfilerActivity.CUDA]) as prof:
                    step()
                    torch.cuda.synchronize()
                for evt in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]:
                    print(f"|  | kernel {evt.key[:60]} | {evt.device_time_total / 1e3:.3f} | x{evt.count} | |", flush=True)
            del out
        del ref, prep, wl
        torch.cuda.empty_cache()
    set_env("PGDVS_SORT_CELLS", None)
    set_env("PGDVS_RASTER_FORCE_GENERIC", None)


if __name__ == "__main__":
    main()
