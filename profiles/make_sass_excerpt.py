#!/usr/bin/env python
"""Write profiles/r02_sass_k_raster_pair8.txt: excerpts of `cuobjdump -sass` of the shipped library for
k_raster_pair<8> — the mbarrier / bulk-copy set-up, the batched walk loop (LDS.128, IMAD keys, the
VIMNMX sort4 + merge4 networks) and the fragment stores.

    python profiles/make_sass_excerpt.py            (needs the in-tree build: python -m pgdvs_b200._build)
"""
import collections
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "ml-pgdvs_b200" / "lib" / "libpgdvs_b200.so"
FUN = "_ZN5pgdvs13k_raster_pairILi8EEEvNS_12RasterParamsE"


def main():
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", FUN, str(LIB)], capture_output=True, text=True, check=True).stdout
    ins = []
    for l in txt.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    ops = collections.Counter()
    for _, t in ins:
        op = t.split()[1] if t.startswith("@") else t.split()[0]
        ops[op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STG", "SYNCS", "UBLKCP")) and "." in op else "")] += 1
    loops = []
    for a, t in ins:
        m = re.search(r"BRA\S*\s+.*?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            loops.append((int(m.group(1), 16), a))
    # the walk loop: the (smallest) loop with the most min / max instructions — the exact-key instantiation
    def minmax(l):
        return sum(1 for a, t in ins if l[0] <= a <= l[1] and "VIMNMX" in t)
    walk = max((l for l in loops if 100 <= (l[1] - l[0]) // 16 <= 400), key=lambda l: (minmax(l), -(l[1] - l[0])), default=None)

    def dump(lo, hi, out, limit=None):
        n = 0
        for a, t in ins:
            if lo <= a <= hi:
                out.append(f"{a:04x}  {t} ;")
                n += 1
                if limit and n >= limit:
                    out.append("...")
                    break

    out = ["# SASS excerpt — k_raster_pair<8> (sm_100a), final round-2 build (batched walk + raster-order epilogue)",
           f"# cuobjdump -sass -fun {FUN} ml-pgdvs_b200/lib/libpgdvs_b200.so   (regenerate: python profiles/make_sass_excerpt.py)",
           "# whole kernel: %d instructions; " % len(ins) + ", ".join(
               f"{k} x{ops[k]}" for k in ("UBLKCP.S", "SYNCS.EXCH", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "VIMNMX3", "VIMNMX", "LDS.128", "STG.E", "IMAD", "BAR") if ops.get(k)),
           ""]
    first_sync = next(a for a, t in ins if "SYNCS.EXCH" in t)
    out.append("## (1) mbarrier init (SYNCS.EXCH), expect-tx arm (SYNCS.ARRIVE.TRANS64) and the per-row 1-D bulk copies (cp.async.bulk -> UBLKCP)")
    dump(first_sync - 0x40, first_sync + 0x20, out)
    out.append("...")
    blk = next(a for a, t in ins if "UBLKCP" in t)
    arr = max(a for a, t in ins if "SYNCS.ARRIVE" in t and a < blk)
    dump(arr - 0x20, blk + 0x10, out)
    out.append("")
    wait = next(a for a, t in ins if "SYNCS.PHASECHK" in t)
    out.append("## (2) wait for the staged rows (mbarrier.try_wait -> SYNCS.PHASECHK.TRANS64.TRYWAIT)")
    dump(wait - 0x10, wait + 0x20, out)
    out.append("")
    if walk:
        out.append(f"## (3) the batched walk loop ({(walk[1] - walk[0]) // 16 + 1} instructions per trip of FOUR records x two pixels: 4 LDS.128, IMAD keys, "
                   "sort4 + merge4 as VIMNMX / VIMNMX3 networks)")
        dump(walk[0], walk[1], out)
        out.append("")
    stg = [a for a, t in ins if t.split()[0].startswith("STG.E.128") or (t.startswith("@") and "STG.E.128" in t)]
    if stg:
        out.append("## (4) fragment stores of the epilogue (idx / zbuf / dists as 128-bit stores)")
        dump(stg[0] - 0x30, stg[0] + 0x90, out)
    (ROOT / "profiles" / "r02_sass_k_raster_pair8.txt").write_text("\n".join(out) + "\n")
    print("\n".join(out[:4]))


if __name__ == "__main__":
    main()
