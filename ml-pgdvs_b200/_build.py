"""Build the sm_100a C-ABI library in-tree (nvcc cross-compiles without a GPU).

    python -m pgdvs_b200._build          # or __graft_entry__.build()

Produces ml-pgdvs_b200/lib/libpgdvs_b200.so.  The .so is git-ignored but travels to the GPU
box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_DIR = PKG_DIR / "lib"
LIB_PATH = LIB_DIR / "libpgdvs_b200.so"
OBJ_DIR = LIB_DIR / "obj"
SOURCES = ["bin.cu", "raster.cu", "composite.cu", "uwp.cu", "knn.cu", "knn_grid.cu", "track.cu", "softsplat.cu", "mesh.cu", "ipc.cu", "outlier.cu"]
# raster.cu is compiled seven times in parallel: part 0 = C entry point, parts 1..6 = one group of
# points_per_pixel instantiations each (the file also builds as a single translation unit)
PARTS = {"raster.cu": [("raster_p%d" % i, ["-DPGDVS_RASTER_PART=%d" % i]) for i in range(7)]}
HEADERS = [CSRC / "common.cuh", PKG_DIR.parent / "include" / "pgdvs_b200.h"]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
]


# extra -D flags for experiments (tools/exp_variants.py), e.g. PGDVS_NVCC_EXTRA="-DPGDVS_RASTER_NO_SORT"
EXTRA_FLAGS = os.environ.get("PGDVS_NVCC_EXTRA", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the product path is CUDA-only and cannot be built without it")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    srcs = [CSRC / s for s in SOURCES if (CSRC / s).exists()]

    jobs = []
    for src in srcs:
        for stem, defs in PARTS.get(src.name, [(src.stem, [])]):
            jobs.append((src, stem, defs))

    def compile_one(job):
        src, stem, defs = job
        obj = OBJ_DIR / (stem + ".o")
        if force or _stale(obj, [src] + HEADERS):
            cmd = ([nvcc] + NVCC_FLAGS + EXTRA_FLAGS + defs + (["-Xptxas", "-v"] if verbose else [])
                   + ["-c", str(src), "-o", str(obj)])
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode != 0:
                print(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src.name}")
        return obj

    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, jobs))
    if force or _stale(LIB_PATH, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH)] + [str(o) for o in objs]
        subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
