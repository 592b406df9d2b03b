#!/usr/bin/env python
"""Builds bench_maps/*.bsp.gz: deterministic synthetic maps (qengine_b200.mapgen) compiled with the
REFERENCE qbsp3 + qvis3 (oracle/_ref, built by `make -C oracle ref` from /root/reference), so that the
reference qrad3 and libqrad_cuda light identical BSPs.  qvis3 on the large grids takes many minutes of
CPU, which is why the compiled maps are committed instead of being rebuilt on the GPU box.

usage: python tools/build_bench_maps.py [name ...]
"""
from __future__ import annotations

import gzip
import shutil
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from qengine_b200 import mapgen  # noqa: E402

# name: (grid n, seed, qbsp3 -chop[, room size])
SPECS = {
    "grid13r512": (13, 4, None, 512),  # BASELINE configs[2]: ~8k faces / ~150k patches at qrad3 -chop 32 (lump < 2 MB)
    "grid6": (6, 1, None),
    "grid10": (10, 1, None),
    "grid26": (26, 1, None),        # BASELINE configs[2]: ~20k faces / ~250k patches at qrad3 -chop 32
    "grid16c128": (16, 3, 128),     # BASELINE configs[4]: ~1M patches at qrad3 -chop 8 (SURVEY.md finding F8)
}


HALLS = {
    "hall2000": (2000, 2),          # BASELINE configs[3]: 2 000 light / light_spot entities in one pillared hall
}


def main():
    tools = mapgen.find_map_tools()
    if tools is None:
        raise SystemExit("oracle/_ref/qbsp3 missing: run `make -C oracle ref`")
    out = ROOT / "bench_maps"
    out.mkdir(exist_ok=True)
    for name in (sys.argv[1:] or list(SPECS) + list(HALLS)):
        t0 = time.time()
        work = Path("/tmp/qrad_bench_maps") / name
        shutil.rmtree(work, ignore_errors=True)
        if name in HALLS:
            nl, seed = HALLS[name]
            bsp = mapgen.build_case(work, name, mapgen.light_hall_map(nl, seed=seed), tools)
        else:
            n, seed, chop = SPECS[name][:3]
            room = SPECS[name][3] if len(SPECS[name]) > 3 else 256
            bsp = mapgen.build_case(work, name, mapgen.room_grid_map(n, n, seed=seed, room=room, clutter=0.5, lamps=0.3),
                                    tools, chop=chop)
        with open(bsp, "rb") as f, gzip.GzipFile(out / (name + ".bsp.gz"), "wb", compresslevel=9, mtime=0) as g:
            shutil.copyfileobj(f, g)
        print(f"{name}: {bsp.stat().st_size} bytes -> {(out / (name + '.bsp.gz')).stat().st_size} gz, {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
