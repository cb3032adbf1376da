#!/usr/bin/env python
"""Developer helper: build a variant of the library with extra -D flags into exp_libs/lib_<name>.so
(objects under /tmp, the in-tree build is untouched).  Used with tools/exp_libs_ab.py.

    python tools/build_variant.py nopair -DPGDVS_RASTER_NO_PAIR
"""
import importlib.util
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
spec = importlib.util.spec_from_file_location("_b", ROOT / "ml-pgdvs_b200" / "_build.py")
b = importlib.util.module_from_spec(spec)
spec.loader.exec_module(b)


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    objdir = Path("/tmp/pgdvs_variants") / name
    objdir.mkdir(parents=True, exist_ok=True)
    jobs = []
    for src in [b.CSRC / s for s in b.SOURCES]:
        for stem, defs in b.PARTS.get(src.name, [(src.stem, [])]):
            jobs.append((src, stem, defs))

    def one(job):
        src, stem, defs = job
        obj = objdir / (stem + ".o")
        cmd = [b._nvcc()] + b.NVCC_FLAGS + flags + defs + ["-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stdout + r.stderr)
            raise SystemExit(1)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(one, jobs))
    out = ROOT / "exp_libs" / f"lib_{name}.so"
    out.parent.mkdir(exist_ok=True)
    subprocess.run([b._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(out)] + [str(o) for o in objs], check=True)
    print(out)


if __name__ == "__main__":
    main()
