#!/usr/bin/env python
"""Developer experiment (not part of the product): the same prepared views rendered
  * through several builds of the library (in-tree first, then exp_libs/lib_<name>.so for every
    name in LIBS=a,b — made with tools/build_variant.py), and/or
  * under several settings of the rasterizer's kernel-selection switches
    (MODES="default;no_pair=1;force_generic=1,sort_cells=1", see _cabi.debug_switch),
with the CUDA-event time of the rasterize-and-composite call, of the whole step, and a bit-for-bit
comparison of every output against the first configuration.

    LIBS=pair4 MODES="default;no_pair=1" CASES="C2,C4" python tools/exp_libs_ab.py     (on the GPU box)
    PROFILE=1 ...   additionally prints per-kernel device times of one step (CUPTI, not a bench number)
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
    ("C1 K=8", "c1_nvidia_1view", dict()),
    ("C2 K=8", "c2_nvidia_seq", dict()),
    ("C4 K=8", "c4_davis", dict(n_views=16)),
    ("C5 K=8 r=0.01", "c5_stress", dict(K=8, radius=0.01, n_views=4)),
    ("C5 K=16 r=0.005", "c5_stress", dict(K=16, radius=0.005, n_views=4)),
    ("C5 K=32 r=0.02", "c5_stress", dict(K=32, radius=0.02, n_views=2)),
    ("C3 K=16", "c3_iphone", dict(n_views=8)),
]


def parse_modes():
    modes = []
    for m in os.environ.get("MODES", "default").split(";"):
        m = m.strip()
        if not m:
            continue
        sw = {}
        if m != "default":
            for kv in m.split(","):
                k, v = kv.split("=")
                sw[k.strip()] = int(v)
        modes.append((m, sw))
    return modes


def main():
    dev = torch.device("cuda:0")
    in_tree = _cabi.LIB_PATH
    libs = [("in-tree", in_tree)] + [(n, ROOT / "exp_libs" / f"lib_{n}.so") for n in os.environ.get("LIBS", "").split(",") if n]
    modes = parse_modes()
    only = os.environ.get("CASES")
    n = int(os.environ.get("REPS", "5"))
    print("| case | build | switches | raster ms | step ms | same bits as first |")
    print("|---|---|---|---|---|---|")
    for label, name, kw in CASES:
        if only and not any(o in label for o in only.split(",")):
            continue
        wl = synthetic.make_workload(name, dev, **kw)
        pairs, cams = wl.jobs(range(wl.n_views))
        prep = prepare_views(pairs, cams, wl.H, wl.W, dev)
        ref = None
        for lib_name, path in libs:
            _cabi.LIB_PATH, _cabi._lib = path, None  # next _cabi.lib() loads this build
            for mode, sw in modes:
                for k in _cabi.DEBUG_SWITCHES:
                    _cabi.debug_switch(k, sw.get(k, -1))

                def step(ev=None):
                    return render_prepared(prep, radius=wl.radius, points_per_pixel=wl.K, compositor="norm",
                                           static_rgb=wl.static_rgb, return_fragments=True, raster_events=ev)
                out = step()
                out = step()
                torch.cuda.synchronize()
                evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
                for i in range(n):
                    out = step(evs[i])
                t1.record()
                torch.cuda.synchronize()
                r_ms = sorted(a.elapsed_time(b) for a, b in evs)[n // 2]
                s_ms = t0.elapsed_time(t1) / n
                keys = ("idx", "zbuf", "dists", "image", "mask")
                if ref is None:
                    ref = {k: out[k].clone() for k in keys}
                    same = "-"
                else:
                    same = "yes" if all(torch.equal(ref[k], out[k]) for k in keys) else \
                        "NO: " + ",".join(k for k in keys if not torch.equal(ref[k], out[k]))
                print(f"| {label} ({wl.n_views} views) | {lib_name} | {mode} | {r_ms:.3f} | {s_ms:.3f} | {same} |", flush=True)
                if os.environ.get("PROFILE"):
                    from torch.profiler import ProfilerActivity, profile
                    with profile(activities=[ProfilerActivity.CUDA]) as prof:
                        step()
                        torch.cuda.synchronize()
                    for evt in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]:
                        print(f"|  |  | kernel {evt.key[:60]} | {evt.device_time_total / 1e3:.3f} | x{evt.count} | |", flush=True)
                del out
            for k in _cabi.DEBUG_SWITCHES:
                _cabi.debug_switch(k, -1)
        del ref, prep, wl
        torch.cuda.empty_cache()
    _cabi.LIB_PATH, _cabi._lib = in_tree, None


if __name__ == "__main__":
    main()
