#!/usr/bin/env python
"""Developer experiment harness: build variants of the CUDA library with -D switches and time the
C-ABI stages per variant with CUDA events on the C2 workload.  Not part of the product.

    python tools/exp_variants.py build   (here, needs nvcc)
    python tools/exp_variants.py run     (on the GPU box)
"""
import ctypes, subprocess, sys, os
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
OUT = ROOT / "exp_libs"
K8 = ["-DPGDVS_RASTER_EXP_K8"]  # only the K=8 instantiation: seconds instead of minutes per variant
VARIANTS = {
    "base": K8,
    "fill8": K8 + ["-DPGDVS_FILL_UNROLL=8"],
    "fill2": K8 + ["-DPGDVS_FILL_UNROLL=2"],
    "uwp_mb2": K8 + ["-DPGDVS_UWP_MINBLOCKS=2"],
    "uwp_mb4": K8 + ["-DPGDVS_UWP_MINBLOCKS=4"],
}
if os.environ.get("PREBUILT"):  # time prebuilt exp_libs/lib_<name>.so files
    VARIANTS = {k: [] for k in os.environ["PREBUILT"].split(",")}
if os.environ.get("VARIANTS"):
    VARIANTS = {k: v for k, v in VARIANTS.items() if k in os.environ["VARIANTS"].split(",")}
SRCS = ["bin.cu", "raster.cu", "composite.cu", "uwp.cu", "knn.cu", "knn_grid.cu"]


def build():
    OUT.mkdir(exist_ok=True)
    procs = []
    for name, flags in VARIANTS.items():
        cmd = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
               "-Xcompiler", "-fPIC", "-shared", "-o", str(OUT / f"lib_{name}.so")] + flags + \
              [str(ROOT / "ml-pgdvs_b200" / "csrc" / s) for s in SRCS]
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)))
    for name, p in procs:
        assert p.wait() == 0, name


def run():
    import torch
    import pgdvs_b200
    from pgdvs_b200 import synthetic
    from pgdvs_b200.dyn_renderer import prepare_views
    dev = torch.device("cuda:0")
    flow = os.environ.get("FLOW", "smooth")
    wl = synthetic.make_workload("c2_nvidia_seq", dev, flow_mode=flow)
    pairs, cams = wl.jobs(range(wl.n_views))
    prep = prepare_views(pairs, cams, wl.H, wl.W, dev)
    prep.pack_frames()
    n_jobs, n_views, H, W, K = prep.n_jobs, prep.n_views, prep.H, prep.W, wl.K
    cap = n_jobs * H * W
    first = torch.empty(n_views, dtype=torch.int64, device=dev)
    num = torch.empty(n_views, dtype=torch.int64, device=dev)
    total = torch.empty(1, dtype=torch.int64, device=dev)
    idx = torch.empty((n_views, H, W, K), dtype=torch.int32, device=dev)
    zbuf = torch.empty((n_views, H, W, K), dtype=torch.float32, device=dev)
    dists = torch.empty((n_views, H, W, K), dtype=torch.float32, device=dev)
    image = torch.empty((n_views, H, W, 3), dtype=torch.float32, device=dev)
    mask = torch.empty((n_views, H, W, 1), dtype=torch.float32, device=dev)
    V = ctypes.c_void_p
    for name in VARIANTS:
        L = ctypes.CDLL(str(OUT / f"lib_{name}.so"))
        L.pgdvs_uwp_bin_workspace_bytes.argtypes = [ctypes.c_int] * 4 + [ctypes.c_float, ctypes.POINTER(ctypes.c_size_t)]
        L.pgdvs_uwp_bin.argtypes = [V, ctypes.c_int, V, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_float] + [V] * 5 + [V, V, ctypes.c_int] + [V, ctypes.c_size_t, V]
        L.pgdvs_rasterize_composite.argtypes = [V, ctypes.c_size_t, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_float, V, V, V, V, V, V, V, V]
        nb = ctypes.c_size_t(0)
        assert L.pgdvs_uwp_bin_workspace_bytes(n_jobs, n_views, H, W, wl.radius, ctypes.byref(nb)) == 0
        ws = torch.empty(nb.value + 256, dtype=torch.uint8, device=dev)
        wp = (ws.data_ptr() + 255) & ~255
        st = torch.cuda.current_stream().cuda_stream

        def uwp():
            rc = L.pgdvs_uwp_bin(prep.jobs_dev.data_ptr(), n_jobs, prep.cams_dev.data_ptr(), n_views, H, W, wl.radius,
                                 None, None, first.data_ptr(), num.data_ptr(), total.data_ptr(), *prep.group_args(),
                                 wp, nb.value, st)
            assert rc == 0, rc

        def ras(frag=True):
            rc = L.pgdvs_rasterize_composite(wp, nb.value, n_views, cap, H, W, K, wl.radius, 0, 3, 2,
                                             wl.radius * wl.radius, None, wl.static_rgb.data_ptr(),
                                             idx.data_ptr() if frag else None, zbuf.data_ptr() if frag else None,
                                             dists.data_ptr() if frag else None, image.data_ptr(), mask.data_ptr(), st)
            assert rc == 0, rc

        def timeit(fn, n=5):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        t_u = timeit(uwp)
        t_r = timeit(ras)
        t_r0 = timeit(lambda: ras(False))
        print(f"{name:10s} flow={flow} uwp_bin {t_u:.3f} ms   raster {t_r:.3f} ms   raster(no frags) {t_r0:.3f} ms", flush=True)


if __name__ == "__main__":
    {"build": build, "run": run}[sys.argv[1]]()
