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
    from pgdvs_b200 import ops, synthetic
    from pgdvs_b200.dyn_renderer import prepare_views, render_prepared

    wl = synthetic.make_workload(args.workload, dev, n_views=args.views, seed=1234 + rank, flow_mode=args.flow,
                                 K=args.K, radius=args.radius)
    V, H, W, K, radius = wl.n_views, wl.H, wl.W, wl.K, wl.radius
    pairs, cams = wl.jobs(range(V))
    prep = prepare_views(pairs, cams, H, W, dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    comm_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    gather_list = None
    gdtype = torch.float32 if args.gather == "f32" else torch.uint8
    if world > 1 and rank == 0:
        gather_list = [torch.empty((V, H, W, 3), dtype=gdtype, device=dev) for _ in range(world)]
    frames_u8 = [torch.empty((V, H, W, 3), dtype=torch.uint8, device=dev) for _ in range(2)] if world > 1 else None
    state_g = {"i": 0}

    def step(ev=None):
        out = render_prepared(prep, radius=radius, points_per_pixel=K, compositor="norm",
                              static_rgb=wl.static_rgb, raster_events=ev, return_fragments=args.fragments)
        if world > 1:
            # finished frames are gathered on rank 0 over NCCL/NVLink on a side stream, overlapped
            # with the next step; by default as the 8-bit frames the reference's evaluator / video
            # writer consume (engines/evaluator_pgdvs.py:75-77), --gather f32 sends raw floats
            payload = out["image"]
            if args.gather == "u8":
                payload = ops.quantize_u8(out["image"], out=frames_u8[state_g["i"] & 1])
                state_g["i"] += 1
            done = torch.cuda.Event()
            done.record()
            comm_stream.wait_event(done)
            with torch.cuda.stream(comm_stream):
                payload.record_stream(comm_stream)
                dist.gather(payload, gather_list, dst=0)
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
    raster_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = ops.LAUNCHES["count"]
    barrier()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (inside the bracket, i.e. counted)
        out = step(raster_ev[i])
    if world > 1:
        torch.cuda.current_stream().wait_stream(comm_stream)
    t_end.record()
    barrier()
    launches = ops.LAUNCHES["count"] - l0
    ms = t_start.elapsed_time(t_end)
    raster_ms = statistics.mean(a.elapsed_time(b) for a, b in raster_ev)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    views_total = V * world
    value = views_total * args.steps / (ms / 1e3)

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
    h2d = sum(t.numel() * t.element_size() for t in host_in.values())
    d2h = host_img.numel() * 4 + host_mask.numel() * 4

    # The step is pipelined over chunks of views on three streams: H2D of the inputs, the
    # kernels, and the D2H of each finished chunk (which overlaps the next chunk's kernels and,
    # PCIe being full duplex, the next step's H2D).
    n_chunks = 4 if (V % 48 == 0) else 1
    per = V // n_chunks
    job_sets = [[w.jobs(range(c * per, (c + 1) * per)) for c in range(n_chunks)] for w in (wl, wl_b)]
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    state = {"inputs_free": [None, None], "step": 0}

    host_img8 = torch.empty((V, H, W, 3), dtype=torch.uint8).pin_memory()
    host_mask8 = torch.empty((V, H, W, 1), dtype=torch.uint8).pin_memory()

    def e2e_step(u8=False):
        cur = torch.cuda.current_stream(dev)
        b = state["step"] & 1
        state["step"] += 1
        dev_in, chunk_jobs = in_sets[b], job_sets[b]
        with torch.cuda.stream(s_in):
            if state["inputs_free"][b] is not None:
                s_in.wait_event(state["inputs_free"][b])  # the kernels that read this set are done
            for k, t in host_in.items():
                dev_in[k].copy_(t, non_blocking=True)
            ev_in = s_in.record_event()
        cur.wait_event(ev_in)
        extra_bytes = 0
        for c in range(n_chunks):
            cp, cc = chunk_jobs[c]
            p = prepare_views(cp, cc, H, W, dev)  # job/camera descriptors: host algebra + small H2D
            extra_bytes += p.h2d_bytes
            o = render_prepared(p, radius=radius, points_per_pixel=K, compositor="norm",
                                static_rgb=wl.static_rgb[c * per:(c + 1) * per])
            img, msk = o["image"], o["mask"]
            dst_i, dst_m = host_img, host_mask
            if u8:  # 8-bit frames as the reference's evaluator / video writer consume them
                img, msk = ops.quantize_u8(img), ops.quantize_u8(msk)
                dst_i, dst_m = host_img8, host_mask8
            ev = cur.record_event()
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev)
                dst_i[c * per:(c + 1) * per].copy_(img, non_blocking=True)
                dst_m[c * per:(c + 1) * per].copy_(msk, non_blocking=True)
                img.record_stream(s_out)
                msk.record_stream(s_out)
        state["inputs_free"][b] = cur.record_event()
        return extra_bytes

    e2e_steps = max(2, min(args.steps, 5))
    extra = e2e_step()
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - w0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = views_total * e2e_steps / e2e_s
    # secondary figure: the same loop delivering 8-bit frames + masks (4x fewer D2H bytes)
    e2e_step(u8=True)
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step(u8=True)
    torch.cuda.synchronize()
    e2e8_s = time.perf_counter() - w0
    if world > 1:
        t = torch.tensor([e2e8_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e8_s = float(t.item())
    e2e8_value = views_total * e2e_steps / e2e8_s

    # -------------------------------------------------- roofline of the dominant kernel
    peak, peak_src = measured_peak_hbm()
    # B_rc of SURVEY.md 8(d) (+ the static frame read by the fused blend); without fragments the
    # 12*K*H*W bytes of idx/zbuf/dists are neither written nor counted
    b_rc = algorithmic_bytes_raster(V, total_points, H, W, K if args.fragments else 0) + 12 * V * H * W
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
        n_threads = os.cpu_count() or 1
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
                       "parallelism": (f"views sharded over {world} GPU(s); NCCL gather of {args.gather} frames to rank 0, "
                                       "overlapped with the next step") if world > 1 else "1 GPU",
                       "cache": f"L2 flushed with a {L2_FLUSH_BYTES >> 20} MiB memset before every step (inside the timed bracket); "
                                f"per-step working set ~{(b_rc + 72 * total_points) / 1e9:.1f} GB >> 126 MB L2"},
            "mpoints_per_s": value * (total_points / V) / 1e6,
            "roofline": {"bound": "hbm", "kernel": "k_raster (rasterize-and-composite)", "achieved": achieved,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "algorithmic_bytes_per_launch": b_rc, "avg_launch_ms": raster_ms},
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d + extra),
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                    "note": ("host pinned inputs -> H2D -> prepare descriptors -> uwp/bin/raster -> D2H of fp32 "
                             f"frames+masks, wall clock; {n_chunks} view chunks pipelined on 3 streams, device inputs double-buffered")},
            "e2e_u8_frames": {"value": e2e8_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d + extra),
                              "d2h_bytes_per_step": int(host_img8.numel() + host_mask8.numel()),
                              "note": "same loop, frames and masks quantised to 8 bit on the GPU "
                                      "(evaluator_pgdvs.py:51-77) before the D2H copy"},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2_nvidia_seq")
    ap.add_argument("--views", type=int, default=None, help="views per GPU (default: the config's)")
    ap.add_argument("--K", type=int, default=None, help="points per pixel (default: the config's)")
    ap.add_argument("--radius", type=float, default=None, help="splat radius in NDC (default: the config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fragments", dest="fragments", action="store_false",
                    help="do not materialise idx/zbuf/dists (fused-only mode; B_rc drops the 12*K*H*W term)")
    ap.add_argument("--ref-step-seconds", type=float, default=4.0)
    ap.add_argument("--gather", default="u8", choices=["u8", "f32"],
                    help="N>1: gather 8-bit frames (what the reference writes / scores) or raw fp32 on rank 0")
    ap.add_argument("--flow", default="smooth", choices=["smooth", "iid"],
                    help="synthetic optical flow: piecewise-smooth (default) or i.i.d. per pixel (stress)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
