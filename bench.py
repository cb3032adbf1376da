#!/usr/bin/env python
"""Benchmark of the PGDVS dynamic-content point-splat hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path (oracle port)

Workload (BASELINE.json configs[1]): the full NVIDIA-Dynamic-Scenes-shaped sequence —
12 time steps x 12 target cameras = 144 views per GPU, 288x544, 2 source frames per view
(P = 313 344 points), K = 8 splats/pixel, radius 0.01, NormWeighted compositing + mask +
static blend.  One "step" renders all 144 views:  fused unproject->warp->project  ->  binning
->  rasterize-and-composite.  Multi-GPU: views are sharded by rank (weak scaling, 144 views per
rank), NCCL is used only to gather the rendered frames on rank 0.

Prints ONE JSON line (see the contract in the task statement / DESIGN.md §measurement).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "rendered_novel_views_per_s"
UNIT = "views/s"
L2_FLUSH_BYTES = 256 << 20


def algorithmic_bytes_raster(n_views, P_total, H, W, K, C=3):
    """B_rc of BASELINE.md §4: read each point's xyz + C features once; write idx/zbuf/dists
    (4 B x K each), the image (4C) and the mask (4) per pixel."""
    return (12 + 4 * C) * P_total + n_views * (12 * K + 4 * C + 4) * H * W


def measured_peak_hbm():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # "under load" = upper half of the samples (idle samples before/after the loop drag the median)
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": statistics.median(load) if load else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ reference arm
def cpu_reference_view(wl_cpu, view, n_threads, rows_budget_s=None, rows=None):
    """The reference's CPU path for ONE target view, restated by the oracle: compute_dyn_pcl
    geometry for every source pair (torch CPU), camera conversion + transform, naive
    rasterization (bin_size=0) TWICE (rgb, then all-ones for the mask) + NormWeighted
    compositing, static blend.  The naive rasterizer is timed on a band of `rows` image rows and
    scaled by H/rows (its cost is exactly linear in rows: every pixel visits every point)."""
    import numpy as np
    import torch
    from oracle import pgdvs_ref as ref
    from oracle import raster as oracle
    sc = wl_cpu.scene
    H, W, K, r = wl_cpu.H, wl_cpu.W, wl_cpu.K, wl_cpu.radius
    t0 = time.perf_counter()
    pcl, rgb = [], []
    Kc, c2w_t = wl_cpu.view_cams[view]
    flat_tgt = torch.cat([torch.tensor([float(H), float(W)]), torch.from_numpy(Kc).reshape(-1),
                          torch.from_numpy(c2w_t).reshape(-1)])
    for p in wl_cpu.view_pairs[view]:
        a = p._src_frames
        o = ref.compute_dyn_pcl(
            dyn_mask_1=sc.mask[a[0]], rgb_1=sc.rgb[a[0]], depth_1=sc.depth[a[0]], flow_12=p.flow_12.reshape(H, W, 2),
            flow_12_occ_mask=torch.zeros(H, W, 1), rgb_2=sc.rgb[a[1]], depth_2=sc.depth[a[1]],
            K_1=torch.from_numpy(sc.K), c2w_1=torch.from_numpy(sc.c2w[a[0]]), K_2=torch.from_numpy(sc.K),
            c2w_2=torch.from_numpy(sc.c2w[a[1]]), time_1=torch.tensor(sc.times[a[0]]),
            time_2=torch.tensor(sc.times[a[1]]), time_tgt=torch.tensor(p._t_tgt))
        pcl.append(o["pcl"])
        rgb.append(o["rgb"])
    pcl, rgb = torch.cat(pcl), torch.cat(rgb)
    ndc = ref.world_to_ndc(pcl, ref.camera_from_flat_cam(flat_tgt)).numpy()
    t_geom = time.perf_counter() - t0
    P = ndc.shape[0]
    fi, npc = np.zeros(1, np.int64), np.full(1, P, np.int64)
    if rows is None:
        tc = time.perf_counter()
        oracle.rasterize_points_rows(ndc, fi, npc, (H, W), r, K, H // 2, H // 2 + 2, n_threads=n_threads)
        per_row = (time.perf_counter() - tc) / 2
        rows = int(max(2, min(H, (rows_budget_s or 4.0) / max(per_row, 1e-6))))
    y0 = max(0, (H - rows) // 2)
    t1 = time.perf_counter()
    idx, zbuf, dists = oracle.rasterize_points_rows(ndc, fi, npc, (H, W), r, K, y0, y0 + rows, n_threads=n_threads)
    t_band = time.perf_counter() - t1
    t2 = time.perf_counter()
    w = (np.float32(1.0) - np.transpose(dists, (0, 3, 1, 2)) / np.float32(r * r)).astype(np.float32)
    idx_l = np.transpose(idx, (0, 3, 1, 2)).astype(np.int64)
    img = oracle.composite(idx_l, w, np.ascontiguousarray(rgb.numpy().T), "norm")
    ones = oracle.composite(idx_l, w, np.ones((3, P), np.float32), "norm")
    mask = (ones[:, :1] > 0).astype(np.float32)
    st = wl_cpu.static_rgb[view, y0:y0 + rows].numpy().transpose(2, 0, 1)[None]
    _ = (1 - mask) * st + mask * img
    t_comp_band = time.perf_counter() - t2
    scale = H / rows
    t_view = t_geom + 2 * t_band * scale + t_comp_band * scale
    return {"t_view_s": t_view, "t_geom_s": t_geom, "t_raster_band_s": t_band, "rows": rows, "P": P,
            "t_wall_s": time.perf_counter() - t0}


def make_cpu_workload(name, n_views):
    import torch
    from pgdvs_b200 import synthetic
    wl = synthetic.make_workload(name, torch.device("cpu"), n_views=n_views)
    return wl


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the
    reference has no compilable sources) on all host threads; each step = one
    target view of the same workload with the naive rasterizer timed on a bounded row band."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_threads = os.cpu_count() or 1
    wl = make_cpu_workload(args.workload, n_views=max(1, min(args.steps + args.warmup, 16)))
    # bounded sample: the timed row band of every step is sized so that the whole run (warm-up
    # included) stays around two minutes whatever --steps is
    budget = min(args.ref_step_seconds, 120.0 / max(1, args.steps + args.warmup))
    rows = None
    times = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_view(wl, i % wl.n_views, n_threads, rows_budget_s=budget, rows=rows)
        rows = r["rows"]
        if i >= args.warmup:
            times.append(r["t_view_s"])
    t_view = statistics.mean(times)
    value = 1.0 / t_view
    sample = (f"1 target view per step ({wl.H}x{wl.W}, P={r['P']}, K={wl.K}); geometry + composite + blend in full, "
              f"naive rasterizer timed on {rows} of {wl.H} rows and scaled by H/rows, counted twice (rgb + mask pass)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_view, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "image": [wl.H, wl.W], "points_per_view": r["P"],
                   "points_per_pixel": wl.K, "radius": wl.radius, "step": "one target view (extrapolated from a row band)"},
        "mpoints_per_s": value * r["P"] / 1e6,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ our arm
def _events(n):
    import torch
    return [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]


def _spread(ms_list):
    s = sorted(ms_list)
    med = s[len(s) // 2]
    out = {"min": s[0], "median": med, "max": s[-1]}
    slow = [(i, round(t, 3)) for i, t in enumerate(ms_list) if t > 1.5 * med]
    if slow:
        out["steps_over_1.5x_median"] = slow[:8]
    return out


def time_batched(wl, dev, steps, warmup, flush, *, K=None, radius=None, fragments=True, views=None):
    """Device-resident timing of the batched hot path on one workload: `steps` timed steps (L2 flushed
    before each, inside the bracket), CUDA events around every step and around the rasterize call."""
    import torch
    from pgdvs_b200.dyn_renderer import prepare_views, render_prepared
    K = K if K is not None else wl.K
    radius = radius if radius is not None else wl.radius
    views = list(views) if views is not None else list(range(wl.n_views))
    pairs, cams = wl.jobs(views)
    prep = prepare_views(pairs, cams, wl.H, wl.W, dev)
    static = wl.static_rgb[views] if len(views) != wl.n_views else wl.static_rgb

    def step(ev=None):
        return render_prepared(prep, radius=radius, points_per_pixel=K, compositor="norm", static_rgb=static,
                               raster_events=ev, return_fragments=fragments)
    for _ in range(warmup):
        out = step()
    torch.cuda.synchronize()
    total_points = int(out["cloud"]["total"].item())
    rev, sev = _events(steps), _events(steps)
    for i in range(steps):
        sev[i][0].record()
        flush.zero_()
        out = step(rev[i])
        sev[i][1].record()
    torch.cuda.synchronize()
    step_ms = [a.elapsed_time(b) for a, b in sev]
    raster_ms = statistics.mean(a.elapsed_time(b) for a, b in rev)
    del out
    return {"step_ms": statistics.mean(step_ms), "spread": _spread(step_ms), "raster_ms": raster_ms,
            "total_points": total_points, "n_views": len(views)}


def matrix_entry(label, wl, t, K, radius, peak, fragments=True):
    V, H, W = t["n_views"], wl.H, wl.W
    b_rc = algorithmic_bytes_raster(V, t["total_points"], H, W, K if fragments else 0)
    return {"config": label, "views_per_step": V, "image": [H, W], "points_per_view": t["total_points"] // max(V, 1),
            "points_per_pixel": K, "radius": radius, "step_ms": t["step_ms"], "views_per_s": V / (t["step_ms"] / 1e3),
            "mpoints_per_s": t["total_points"] / (t["step_ms"] / 1e3) / 1e6, "raster_ms": t["raster_ms"],
            "frac": b_rc / (t["raster_ms"] / 1e3) / 1e9 / peak}


def run_matrix(dev, flush, peak, args):
    """The other BASELINE.json configs on one GPU (inputs resident in HBM, same timing rules as the
    main workload): C1, C2 with the statistical outlier filter on (as in every published run of the
    reference: scripts/benchmark.sh:81,100), C3 without and with the track branch, C4, C5 at
    K in {8, 16, 32}."""
    import torch
    from types import SimpleNamespace
    from pgdvs_b200 import synthetic, track
    from pgdvs_b200.dyn_renderer import CapturedRender, prepare_views, render_views_filtered
    rows = []
    ms = max(3, min(args.matrix_steps, args.steps))

    def batched(label, name, **kw):
        n_views = kw.pop("n_views", None)
        wl = synthetic.make_workload(name, dev, n_views=n_views, K=kw.get("K"), radius=kw.get("radius"))
        t = time_batched(wl, dev, ms, 2, flush)
        rows.append(matrix_entry(label, wl, t, wl.K, wl.radius, peak))
        del wl
        torch.cuda.empty_cache()

    batched("c1_nvidia_1view", "c1_nvidia_1view")
    # the same single view replayed from a CUDA graph (dyn_renderer.CapturedRender): one launch per step
    wl = synthetic.make_workload("c1_nvidia_1view", dev)
    pairs, cams = wl.jobs(range(wl.n_views))
    cap = CapturedRender(prepare_views(pairs, cams, wl.H, wl.W, dev), radius=wl.radius, points_per_pixel=wl.K,
                         compositor="norm", static_rgb=wl.static_rgb, return_fragments=True)
    for _ in range(2):
        cap.replay()
    torch.cuda.synchronize()
    ev = _events(ms)
    for a, b in ev:
        a.record()
        flush.zero_()
        cap.replay()
        b.record()
    torch.cuda.synchronize()
    t_ms = statistics.mean(a.elapsed_time(b) for a, b in ev)
    ev = _events(ms)
    for a, b in ev:  # back to back, no flush: the latency of a launch-bound call
        a.record()
        cap.replay()
        b.record()
    torch.cuda.synchronize()
    t_hot = statistics.mean(a.elapsed_time(b) for a, b in ev)
    rows.append({"config": "c1_nvidia_1view (CUDA graph replay)", "views_per_step": 1, "image": [wl.H, wl.W],
                 "points_per_view": wl.points_per_view(), "points_per_pixel": wl.K, "radius": wl.radius, "step_ms": t_ms,
                 "views_per_s": 1.0 / (t_ms / 1e3), "step_ms_no_flush": t_hot,
                 "note": "render_prepared captured once (dyn_renderer.CapturedRender), one graph launch per step; "
                         "step_ms includes the 256 MiB L2 flush like every other row, step_ms_no_flush does not"})
    del wl, pairs, cams, cap
    torch.cuda.empty_cache()
    # C2 with dyn_pcl_remove_outlier: KNN (K = 50) statistics per source pair on the device, no host sync
    wl = synthetic.make_workload("c2_nvidia_seq", dev)
    pairs, cams = wl.jobs(range(wl.n_views))
    cfg = SimpleNamespace(dyn_pcl_outlier_knn=50, dyn_pcl_outlier_std_thres=0.1)

    def step_f():
        return render_views_filtered(pairs, cams, wl.H, wl.W, radius=wl.radius, points_per_pixel=wl.K, compositor="norm",
                                     static_rgb=wl.static_rgb, render_cfg=cfg, return_fragments=True)
    for _ in range(2):
        out = step_f()
    torch.cuda.synchronize()
    ev = _events(ms)
    for a, b in ev:
        a.record()
        flush.zero_()
        out = step_f()
        b.record()
    torch.cuda.synchronize()
    t_ms = statistics.mean(a.elapsed_time(b) for a, b in ev)
    kept = int(out["cloud"]["total"].item())
    rows.append({"config": "c2_nvidia_seq+outlier_filter", "views_per_step": wl.n_views, "image": [wl.H, wl.W],
                 "points_per_view": kept // wl.n_views, "points_per_pixel": wl.K, "radius": wl.radius, "step_ms": t_ms,
                 "views_per_s": wl.n_views / (t_ms / 1e3), "mpoints_per_s": kept / (t_ms / 1e3) / 1e6,
                 "note": "uniform-grid KNN (K=50) + median/std threshold per source pair, all on the device"})
    del wl, pairs, cams, out
    torch.cuda.empty_cache()
    batched("c3_iphone", "c3_iphone")
    # C3 through the L2 renderer with the track branch (1 closest pair + tracks over +-3 frames, F = 8)
    data = synthetic.make_data_dict("c3_iphone", dev, n_views=16, n_track_one_side=3, closest_mask_mode="ellipse")
    c3 = synthetic.CONFIGS["c3_iphone"]
    rcfg = SimpleNamespace(dyn_render_type="pcl", dyn_render_pcl_pt_radius=c3["radius"], dyn_render_pcl_pts_per_pixel=c3["K"],
                           dyn_render_use_flow_consistency=False, dyn_pcl_remove_outlier=False, dyn_pcl_outlier_knn=50,
                           dyn_pcl_outlier_std_thres=0.1, dyn_pcl_track_track2base_thres_mult=50)
    # (visibles ~ Bernoulli(0.5) instead of SURVEY 8d's 0.8: a track only counts when BOTH closest frames
    #  miss it, 4 % of the queries at 0.8 — too sparse a cloud to survive the reference's own KNN filters)
    rend = track.PGDVSDynamicTrackRenderer(tracker=synthetic.SyntheticTracker(seed=1234, p_visible=0.5))
    for _ in range(2):
        rgb, mask, info = rend(data, None, rcfg)
    torch.cuda.synchronize()
    ev = _events(ms)
    for a, b in ev:
        a.record()
        flush.zero_()
        rgb, mask, info = rend(data, None, rcfg)
        b.record()
    torch.cuda.synchronize()
    t_ms = statistics.mean(a.elapsed_time(b) for a, b in ev)
    rows.append({"config": "c3_iphone+tracks (L2 forward)", "views_per_step": 16, "image": [c3["H"], c3["W"]],
                 "points_per_pixel": c3["K"], "radius": c3["radius"], "step_ms": t_ms, "views_per_s": 16 / (t_ms / 1e3),
                 "track_pixels_per_view": float(((info["temporal_closest_mask"] == 0) & (info["temporal_track_mask"] > 0)).sum()) / 16,
                 "note": "PGDVSDynamicTrackRenderer.forward, realistic masks (centred ellipse, 15 % of the pixels, in every "
                         "frame): closest-pair cloud + KNN statistics, synthetic tracks "
                         "(F = 8, visibles ~ Bernoulli(0.5)), track cloud + 2 KNN filters, batched splat, merge; "
                         "includes the synthetic tracker and the reference's per-view host syncs"})
    del data, rend, rgb, mask, info
    torch.cuda.empty_cache()
    batched("c4_davis", "c4_davis")
    batched("c5_stress K=8 r=0.01", "c5_stress", K=8, radius=0.01, n_views=8)
    batched("c5_stress K=16 r=0.005", "c5_stress", K=16, radius=0.005, n_views=8)
    batched("c5_stress K=32 r=0.02", "c5_stress", K=32, radius=0.02, n_views=4)
    return rows


def run_strong(dev, rank, world, flush, args):
    """Strong scaling: a FIXED job sharded over the ranks exactly as the reference shards target
    views (DistributedSampler(shuffle=False): view v -> rank v mod world, trainer_pgdvs.py:290-306),
    8-bit frames gathered on rank 0 — dist.shard_views / dist.gather_frames, the code the gloo tests
    cover.  Time = barrier-to-barrier device time of the slowest rank, median of 3 repetitions."""
    import torch
    import torch.distributed as dist
    from pgdvs_b200 import dist as pdist
    from pgdvs_b200 import synthetic
    from pgdvs_b200.dyn_renderer import prepare_views, render_prepared
    out = {}
    for label, name, kw in (("c3_iphone_16_views", "c3_iphone", dict()),
                            ("c4_davis_80_frames", "c4_davis", dict()),
                            ("c5_stress_k8_8_views", "c5_stress", dict(K=8, radius=0.01, n_views=8))):
        wl = synthetic.make_workload(name, dev, seed=1234, **kw)  # the same job on every rank
        V = wl.n_views
        mine = pdist.shard_views(V, rank, world, pad=True)
        pairs, cams = wl.jobs(mine)
        prep = prepare_views(pairs, cams, wl.H, wl.W, dev)
        static = wl.static_rgb[mine]

        def job():
            o = render_prepared(prep, radius=wl.radius, points_per_pixel=wl.K, compositor="norm", static_rgb=static,
                                return_fragments=True, return_u8=True)
            if world > 1:
                return pdist.gather_frames(o["image_u8"], V, dst=0)
            return o["image_u8"]
        job()
        times = []
        for _ in range(3):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            flush.zero_()
            frames = job()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times.append(float(t.item()))
        t_ms = sorted(times)[1]
        if rank == 0:
            assert frames.shape[0] == V
        out[label] = {"views": V, "ms": t_ms, "views_per_s": V / (t_ms / 1e3), "views_per_rank": len(mine)}
        del wl, prep, pairs, cams, static, frames
        torch.cuda.empty_cache()
    return out


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import pgdvs_b200
    from pgdvs_b200 import dist as pdist
    from pgdvs_b200 import ops, synthetic
    from pgdvs_b200.dyn_renderer import prepare_views, render_prepared

    # every rank on its own share of the host cores (its GPU's NUMA node when sysfs tells), before
    # any pinned buffer is allocated
    binding = ({"bound": False, "note": "--no-bind"} if not args.bind else
               pdist.bind_rank_to_cores(local_rank, int(os.environ.get("LOCAL_WORLD_SIZE", world))))

    wl = synthetic.make_workload(args.workload, dev, n_views=args.views, seed=1234 + rank, flow_mode=args.flow,
                                 K=args.K, radius=args.radius)
    V, H, W, K, radius = wl.n_views, wl.H, wl.W, wl.K, wl.radius
    pairs, cams = wl.jobs(range(V))
    prep = prepare_views(pairs, cams, H, W, dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    # N > 1: finished 8-bit frames (quantised in the rasterizer's epilogue, evaluator_pgdvs.py:51-77)
    # go to rank 0 as copy-engine peer writes (dist.PeerFrameSink); --gather nccl / f32 selects the
    # NCCL gather instead (also the fallback when CUDA IPC is unavailable)
    sink, gather_mode = None, None
    comm_stream, gather_list = None, None
    if world > 1:
        gather_mode = args.gather
        if gather_mode == "peer":
            ok = torch.ones(1, device=dev)
            try:
                sink = pdist.PeerFrameSink((V, H, W, 3), torch.uint8, dev, dst=0)
            except Exception as e:  # noqa: BLE001
                ok.zero_()
                print(f"[bench] rank {rank}: peer sink unavailable ({e!r}), falling back to the NCCL gather", file=sys.stderr)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if float(ok.item()) == 0.0:
                sink, gather_mode = None, "nccl"
        if sink is None:
            comm_stream = torch.cuda.Stream(device=dev)
            gdtype = torch.float32 if gather_mode == "f32" else torch.uint8
            if rank == 0:
                gather_list = [torch.empty((V, H, W, 3), dtype=gdtype, device=dev) for _ in range(world)]
    state_g = {"i": 0}
    # 8-bit frames for the gather come from the rasterizer's epilogue: with the raster-order epilogue
    # their stores are coalesced (+0.02 ms per 144 frames; the stand-alone quantiser is 0.075 ms and,
    # with the work-sorted epilogue of session 2, the fused output cost +0.14 ms)
    want_u8 = world > 1 and gather_mode != "f32"
    # caller-owned double buffer for them: the transfer of step i reads buffer i & 1 while step i + 1 renders
    # into the other one (freshly allocated tensors held across steps made the caching allocator grow
    # inside the timed region: 13 - 49 ms cudaMalloc stalls in the first ten steps)
    frames_u8 = ([(torch.empty((V, H, W, 3), dtype=torch.uint8, device=dev), torch.empty((V, H, W, 1), dtype=torch.uint8, device=dev))
                  for _ in range(2)] if want_u8 else None)

    def step(ev=None):
        i = state_g["i"]
        if want_u8 and state_g.get(("ev", i & 1)) is not None:  # the transfer that last read this buffer (step i - 2) is done
            torch.cuda.current_stream().wait_event(state_g[("ev", i & 1)])
        out = render_prepared(prep, radius=radius, points_per_pixel=K, compositor="norm",
                              static_rgb=wl.static_rgb, raster_events=ev, return_fragments=args.fragments,
                              return_u8=want_u8, u8_out=frames_u8[i & 1] if want_u8 else None)
        if world > 1:
            state_g["i"] += 1
            done = torch.cuda.Event()
            done.record()
            if sink is not None:
                sink.push(out["image_u8"], i, after=done)
                state_g[("ev", i & 1)] = sink.stream.record_event()
                sink.commit()
            else:
                payload = out["image"] if gather_mode == "f32" else out["image_u8"]
                comm_stream.wait_event(done)
                with torch.cuda.stream(comm_stream):
                    payload.record_stream(comm_stream)
                    dist.gather(payload, gather_list, dst=0)
                state_g[("ev", i & 1)] = comm_stream.record_event()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out = step()
    barrier()
    total_points = int(out["cloud"]["total"].item())

    # -------------------------------------------------- timed region 1: inputs resident in HBM
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    raster_ev, step_ev = _events(args.steps), _events(args.steps)
    l0 = ops.LAUNCHES["count"]
    barrier()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(args.steps):
        step_ev[i][0].record()
        flush.zero_()  # L2 flush between timed iterations (inside the bracket, i.e. counted)
        out = step(raster_ev[i])
        step_ev[i][1].record()
    if world > 1:
        torch.cuda.current_stream().wait_stream(sink.stream if sink is not None else comm_stream)
    t_end.record()
    barrier()
    launches = ops.LAUNCHES["count"] - l0
    ms = t_start.elapsed_time(t_end)
    raster_ms = statistics.mean(a.elapsed_time(b) for a, b in raster_ev)
    spread = _spread([a.elapsed_time(b) for a, b in step_ev])
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    views_total = V * world
    value = views_total * args.steps / (ms / 1e3)
    gather_check = None
    if sink is not None and rank == 0:
        # the last step's frames of every rank have landed (the barrier above): rank 0's own slot
        # must equal what it rendered
        last = state_g["i"] - 1
        gather_check = bool(torch.equal(sink.frames(last)[0], out["image_u8"]))

    # -------------------------------------------------- timed region 2: end to end, host buffers
    sc = wl.scene
    host_in = {k: getattr(sc, k).cpu().pin_memory() for k in ("rgb", "depth", "mask", "flow_next", "flow_prev")}
    # two sets of device input buffers (the SourcePairs point into them): the H2D copy of step
    # s+1 fills one set while the kernels of step s still read the other
    wl_b = synthetic.make_workload(args.workload, dev, n_views=args.views, seed=1234 + rank, flow_mode=args.flow,
                                   K=args.K, radius=args.radius)
    in_sets = [{k: getattr(w.scene, k) for k in host_in} for w in (wl, wl_b)]
    host_img = torch.empty((V, H, W, 3), dtype=torch.float32).pin_memory()
    host_mask = torch.empty((V, H, W, 1), dtype=torch.float32).pin_memory()
    host_img8 = torch.empty((V, H, W, 3), dtype=torch.uint8).pin_memory()
    host_mask8 = torch.empty((V, H, W, 1), dtype=torch.uint8).pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in host_in.values())

    # The step is pipelined over chunks of views on three streams: H2D of the inputs, the
    # kernels, and the D2H of each finished chunk (which overlaps the next chunk's kernels and,
    # PCIe being full duplex, the next step's H2D).  Job descriptors are built once per (input
    # set, chunk) — they only hold pointers into the device input buffers and source-camera
    # constants; the target cameras, which change from step to step in a real run, are uploaded
    # from pinned memory every step.
    # Chunks per step: 4 for fp32 frames (361 MB of D2H per step: small copies keep the link busy from the
    # first chunk on), 2 for 8-bit frames (90 MB: the copy hides behind one chunk, larger launches run closer
    # to the device-resident rate; measured 2 / 3 / 4 / 6 / 8 chunks: 49.7 / 48.0 / 45.3 / 42.1 / 39.5 k views/s).
    def chunks_for(u8):
        want = args.e2e_chunks if args.e2e_chunks > 0 else (2 if u8 else 4)
        return want if V % want == 0 and (want != 4 or V % 48 == 0) else 1

    pipes = {}
    for nc in sorted({chunks_for(False), chunks_for(True)}):
        pp = [[prepare_views(*w.jobs(range(c * (V // nc), (c + 1) * (V // nc))), H, W, dev) for c in range(nc)] for w in (wl, wl_b)]
        pipes[nc] = (pp, [[q.cams_dev.cpu().pin_memory() for q in ps] for ps in pp])
    cam_bytes = sum(t.numel() for t in pipes[chunks_for(False)][1][0])
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    state = {"inputs_free": [None, None], "step": 0}

    def e2e_step(u8=False):
        n_chunks = chunks_for(u8)
        per = V // n_chunks
        preps, host_cams = pipes[n_chunks]
        cur = torch.cuda.current_stream(dev)
        b = state["step"] & 1
        state["step"] += 1
        dev_in = in_sets[b]
        with torch.cuda.stream(s_in):
            if state["inputs_free"][b] is not None:
                s_in.wait_event(state["inputs_free"][b])  # the kernels that read this set are done
            for k, t in host_in.items():
                dev_in[k].copy_(t, non_blocking=True)
            for c in range(n_chunks):
                preps[b][c].cams_dev.copy_(host_cams[b][c], non_blocking=True)
            ev_in = s_in.record_event()
        cur.wait_event(ev_in)
        for c in range(n_chunks):
            o = render_prepared(preps[b][c], radius=radius, points_per_pixel=K, compositor="norm",
                                static_rgb=wl.static_rgb[c * per:(c + 1) * per], return_u8=u8, return_f32=not u8)
            if u8:  # 8-bit frames as the reference's evaluator / video writer consume them, from the epilogue
                img, msk, dst_i, dst_m = o["image_u8"], o["mask_u8"], host_img8, host_mask8
            else:
                img, msk, dst_i, dst_m = o["image"], o["mask"], host_img, host_mask
            ev = cur.record_event()
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev)
                dst_i[c * per:(c + 1) * per].copy_(img, non_blocking=True)
                dst_m[c * per:(c + 1) * per].copy_(msk, non_blocking=True)
                img.record_stream(s_out)
                msk.record_stream(s_out)
        state["inputs_free"][b] = cur.record_event()

    def e2e_run(u8):
        # wall clock of n steps, max over ranks; the MEDIAN of three such runs (the leg is bound by PCIe
        # and host memory, which other tenants of the box share: single runs scatter by +-20 %)
        n = max(4, min(args.steps, 20))
        e2e_step(u8)
        e2e_step(u8)
        vals = []
        for _ in range(3):
            barrier()
            w0 = time.perf_counter()
            for _ in range(n):
                e2e_step(u8)
            torch.cuda.synchronize()
            dt = time.perf_counter() - w0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            vals.append(views_total * n / dt)
        return sorted(vals)[1], n

    e2e_value, e2e_steps = e2e_run(False)
    e2e8_value, _ = e2e_run(True)

    # -------------------------------------------------- roofline of the dominant kernel
    peak, peak_src = measured_peak_hbm()
    # B_rc of SURVEY.md 8(d): read (x, y, z) + C features per point, write idx / zbuf / dists, image and
    # mask per pixel (the 12 B / pixel the fused blend reads from the static frame are NOT counted);
    # without fragments the 12*K*H*W bytes of idx/zbuf/dists are neither written nor counted
    b_rc = algorithmic_bytes_raster(V, total_points, H, W, K if args.fragments else 0)
    achieved = b_rc / (raster_ms / 1e3) / 1e9
    traffic = None
    tp = ROOT / "profiles" / "raster_traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    # -------------------------------------------------- CPU baseline (rank 0, N = 1 only)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import raster as oracle
        oc = render_prepared(prep, radius=radius, points_per_pixel=K, compositor="norm", return_cloud=True)
        fi = oc["first_idx"].cpu().numpy()
        npc = oc["num_points"].cpu().numpy()
        ndc = oc["cloud"]["xyz_ndc"][fi[0]:fi[0] + npc[0]].cpu().numpy()
        n_threads = len(os.sched_getaffinity(0)) or 1
        z, n1 = np.zeros(1, np.int64), np.full(1, ndc.shape[0], np.int64)
        tc = time.perf_counter()
        oracle.rasterize_points_rows(ndc, z, n1, (H, W), radius, K, H // 2, H // 2 + 2, n_threads=n_threads)
        per_row = (time.perf_counter() - tc) / 2
        rows = int(max(2, min(H, 6.0 / max(per_row, 1e-6))))
        y0 = (H - rows) // 2
        tc = time.perf_counter()
        oracle.rasterize_points_rows(ndc, z, n1, (H, W), radius, K, y0, y0 + rows, n_threads=n_threads)
        t_band = time.perf_counter() - tc
        t_view = 2 * t_band * H / rows  # two rasterization passes per view (rgb + mask)
        cpu_baseline = {
            "value": 1.0 / t_view, "unit": UNIT, "cores": n_threads, "kind": "port",
            "sample": (f"oracle naive rasterizer (pytorch3d RasterizePointsNaiveCpu restatement, row-parallel on "
                       f"{n_threads} threads) on view 0's NDC cloud (P={ndc.shape[0]}), {rows} of {H} rows timed "
                       f"({t_band:.2f} s) and scaled by H/rows, x2 passes (rgb + mask)")}
        del oc

    # -------------------------------------------------- the other configs / the strong-scaling leg
    del wl_b, pipes, in_sets, host_img, host_mask, out
    torch.cuda.empty_cache()
    matrix = run_matrix(dev, flush, peak, args) if (world == 1 and rank == 0 and args.matrix) else None
    strong = run_strong(dev, rank, world, flush, args) if args.strong else None

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "views_per_gpu": V, "image": [H, W],
                       "source_frames_per_view": wl.meta["S"], "points_per_view": total_points // V,
                       "points_per_pixel": K, "radius": radius, "compositor": "norm_weighted+mask+static_blend",
                       "fragments_written": bool(args.fragments),
                       "synthetic_flow": ("9x9-box-smoothed N(0,3px) + 0.1px jitter (piecewise-smooth motion)"
                                          if args.flow == "smooth" else "i.i.d. N(0,3px) per pixel (incoherent stress case)"),
                       "parallelism": (f"views sharded over {world} GPU(s), {V} per rank; 8-bit frames delivered to rank 0 "
                                       + ("as copy-engine peer writes over NVLink (CUDA IPC buffer, no kernels)" if sink is not None
                                          else f"with an NCCL gather ({gather_mode})") + ", overlapped with the next step")
                       if world > 1 else "1 GPU",
                       "cache": f"L2 flushed with a {L2_FLUSH_BYTES >> 20} MiB memset before every step (inside the timed bracket); "
                                f"per-step working set ~{(b_rc + 72 * total_points) / 1e9:.1f} GB >> 126 MB L2",
                       "host_binding": binding},
            "ms_per_step_spread": spread,
            "mpoints_per_s": value * (total_points / V) / 1e6,
            "roofline": {"bound": "hbm", "kernel": "k_raster (rasterize-and-composite)", "achieved": achieved,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "algorithmic_bytes_per_launch": b_rc, "avg_launch_ms": raster_ms},
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d + cam_bytes),
                    "d2h_bytes_per_step": int(V * H * W * 16), "steps": e2e_steps,
                    "note": ("host pinned inputs + target cameras -> H2D -> uwp/bin/raster -> D2H of fp32 frames+masks, wall "
                             f"clock, median of 3 runs of {e2e_steps} steps; {chunks_for(False)} view chunks pipelined on 3 streams, "
                             "device inputs double-buffered")},
            "e2e_u8_frames": {"value": e2e8_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d + cam_bytes),
                              "d2h_bytes_per_step": int(V * H * W * 4),
                              "note": f"same loop in {chunks_for(True)} chunks, 8-bit frames and masks written by the rasterizer's "
                                      "epilogue (evaluator_pgdvs.py:51-77) instead of fp32"},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if gather_check is not None:
            line["gather_check"] = gather_check
        if matrix is not None:
            line["matrix"] = matrix
        if strong is not None:
            line["strong_scaling"] = strong
        print(json.dumps(line))
    if sink is not None:
        sink.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2_nvidia_seq")
    ap.add_argument("--views", type=int, default=None, help="views per GPU (default: the config's)")
    ap.add_argument("--e2e-chunks", type=int, default=0, help="view chunks of the pipelined end-to-end leg (default: 4 when they divide the views)")
    ap.add_argument("--K", type=int, default=None, help="points per pixel (default: the config's)")
    ap.add_argument("--radius", type=float, default=None, help="splat radius in NDC (default: the config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fragments", dest="fragments", action="store_false",
                    help="do not materialise idx/zbuf/dists (fused-only mode; B_rc drops the 12*K*H*W term)")
    ap.add_argument("--ref-step-seconds", type=float, default=4.0)
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl", "f32"],
                    help="N>1: 8-bit frames to rank 0 as copy-engine peer writes (default), as an NCCL gather, "
                         "or raw fp32 frames over NCCL")
    ap.add_argument("--no-bind", dest="bind", action="store_false", help="do not pin the rank to its GPU's host cores")
    ap.add_argument("--no-matrix", dest="matrix", action="store_false",
                    help="skip the per-config matrix (C1, C2+outlier filter, C3, C3+tracks, C4, C5) at N=1")
    ap.add_argument("--matrix-steps", type=int, default=5)
    ap.add_argument("--no-strong", dest="strong", action="store_false",
                    help="skip the strong-scaling leg (C4's 80 frames / C5's 8 views sharded over the ranks)")
    ap.add_argument("--flow", default="smooth", choices=["smooth", "iid"],
                    help="synthetic optical flow: piecewise-smooth (default) or i.i.d. per pixel (stress)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
