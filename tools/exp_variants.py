#!/usr/bin/env python
"""Developer experiment harness: build variants of the CUDA library with -D switches and time one
C-ABI stage per variant with CUDA events on the C2 workload.  Not part of the product.

    python tools/exp_variants.py build   (here, needs nvcc)
    python tools/exp_variants.py run     (on the GPU box)
"""
import ctypes, subprocess, sys, os
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
OUT = ROOT / "exp_libs"
VARIANTS = {
    "base": [],
    "no_atomic": ["-DPGDVS_EXP_NO_ATOMIC"],
    "no_taps": ["-DPGDVS_EXP_NO_TAPS"],
    "no_store": ["-DPGDVS_EXP_NO_STORE"],
    "no_geom_div": ["-DPGDVS_EXP_FAST_DIV"],
}
SRCS = ["bin.cu", "raster.cu", "composite.cu", "uwp.cu", "knn.cu"]


def build():
    OUT.mkdir(exist_ok=True)
    procs = []
    for name, flags in VARIANTS.items():
        cmd = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
               "-Xcompiler", "-fPIC", "-shared", "-o", str(OUT / f"lib_{name}.so")] + flags + \
              [str(ROOT / "ml-pgdvs_b200" / "csrc" / s) for s in SRCS]
        procs.append((name, subprocess.Popen(cmd)))
    for name, p in procs:
        assert p.wait() == 0, name


def run():
    import torch
    import pgdvs_b200
    from pgdvs_b200 import synthetic, ops, _cabi
    from pgdvs_b200.dyn_renderer import prepare_views
    dev = torch.device("cuda:0")
    wl = synthetic.make_workload("c2_nvidia_seq", dev)
    pairs, cams = wl.jobs(range(wl.n_views))
    prep = prepare_views(pairs, cams, wl.H, wl.W, dev)
    prep.pack_frames()
    n_jobs, n_views, H, W = prep.n_jobs, prep.n_views, prep.H, prep.W
    first = torch.empty(n_views, dtype=torch.int64, device=dev)
    num = torch.empty(n_views, dtype=torch.int64, device=dev)
    total = torch.empty(1, dtype=torch.int64, device=dev)
    for name in VARIANTS:
        L = ctypes.CDLL(str(OUT / f"lib_{name}.so"))
        L.pgdvs_uwp_bin_workspace_bytes.argtypes = [ctypes.c_int] * 4 + [ctypes.c_float, ctypes.POINTER(ctypes.c_size_t)]
        L.pgdvs_uwp_bin.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_float] + [ctypes.c_void_p] * 6 + [ctypes.c_size_t, ctypes.c_void_p]
        nb = ctypes.c_size_t(0)
        assert L.pgdvs_uwp_bin_workspace_bytes(n_jobs, n_views, H, W, wl.radius, ctypes.byref(nb)) == 0
        ws = torch.empty(nb.value + 256, dtype=torch.uint8, device=dev)
        wp = (ws.data_ptr() + 255) & ~255
        st = torch.cuda.current_stream().cuda_stream

        def call():
            rc = L.pgdvs_uwp_bin(prep.jobs_dev.data_ptr(), n_jobs, prep.cams_dev.data_ptr(), n_views, H, W, wl.radius,
                                 None, None, first.data_ptr(), num.data_ptr(), total.data_ptr(), wp, nb.value, st)
            assert rc == 0, rc
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            call()
        e1.record()
        torch.cuda.synchronize()
        print(f"{name:14s} uwp_bin total {e0.elapsed_time(e1) / 5:.3f} ms  (points {int(total)})", flush=True)


if __name__ == "__main__":
    {"build": build, "run": run}[sys.argv[1]]()
