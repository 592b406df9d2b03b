#!/usr/bin/env python
"""Regenerates the qvis3 fixtures of tests/golden/ from the REFERENCE run in this container (needs `make -C oracle ref`).

Every case: a synthetic .map --reference qbsp3--> <case>.bsp + <case>.prt --oracle/_ref/qvis_harness (the reference's qvis3
translation units, oracle/qvis_harness.c)--> per-portal dump.  Committed:

  vis_<case>.prt.gz     the portal file (input of LoadPortals)
  vis_<case>.npz        from the reference's dump: planes [P, 4], leaf [P], nummightsee [P], rank in SortPortals' order
                        [P], and the bit vectors portalfront / portalflood / portalvis as the reference holds them
                        ([P, portalbytes] uint8), plus the md5 of the visibility lump the run wrote, and the md5 of the
                        lump and of the whole file the unmodified `qvis3 -fast` writes for the same input
  vis_<case>.bsp.gz     qbsp3's output (the input of qvis3)
"""
from __future__ import annotations

import gzip
import hashlib
import shutil
import struct
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "oracle"))
sys.path.insert(0, str(HERE))

from qengine_b200 import mapgen  # noqa: E402
import make_golden as MG  # noqa: E402
import qvis_stage1 as V  # noqa: E402

REF = REPO / "oracle" / "_ref"
CASES = {
    "grid3": lambda: mapgen.room_grid_map(3, 3, seed=1, clutter=0.5, lamps=0.3),
    "grid5": lambda: mapgen.room_grid_map(5, 5, seed=2, clutter=1.0, lamps=0.3),
    "angled": MG.angled_map_text,
}


def main():
    for name, text in CASES.items():
        with tempfile.TemporaryDirectory() as tmp:
            assets = mapgen.write_assets(Path(tmp))
            mp = assets / "maps" / (name + ".map")
            mp.write_text(text())
            r = subprocess.run([str(REF / "qbsp3"), str(mp)], capture_output=True, text=True)
            assert r.returncode == 0 and "LEAKED" not in r.stdout, r.stdout[-500:]
            bsp, prt = mp.with_suffix(".bsp"), mp.with_suffix(".prt")
            dump = Path(tmp) / "dump.bin"
            pre = bsp.read_bytes()     # qbsp3's output: no visibility yet
            r = subprocess.run([str(REF / "qvis_harness"), str(bsp), str(prt), str(dump)], capture_output=True, text=True)
            assert r.returncode == 0 and "QVISDUMP" in r.stdout, r.stdout[-500:]
            d = V.read_dump(dump)
            raw = bsp.read_bytes()
            ofs, ln = struct.unpack_from("<ii", raw, 8 + 3 * 8)
            # the same map through `qvis3 -fast` (the unmodified tool): lump and whole file
            bsp.write_bytes(pre)
            r = subprocess.run([str(REF / "qvis3"), "-fast", str(bsp)], capture_output=True, text=True)
            assert r.returncode == 0, r.stdout[-500:]
            fast = bsp.read_bytes()
            fofs, fln = struct.unpack_from("<ii", fast, 8 + 3 * 8)
            with gzip.GzipFile(HERE / f"vis_{name}.bsp.gz", "wb", compresslevel=9, mtime=0) as g:
                g.write(pre)
            with open(prt, "rb") as f, gzip.GzipFile(HERE / f"vis_{name}.prt.gz", "wb", compresslevel=9, mtime=0) as g:
                shutil.copyfileobj(f, g)
            np.savez_compressed(HERE / f"vis_{name}.npz", planes=d["plane"], leaf=d["leaf"], might=d["might"], rank=d["rank"],
                                front=V.pack_bits(d["front"]), flood=V.pack_bits(d["flood"]), vis=V.pack_bits(d["vis"]),
                                lump_md5=np.array(hashlib.md5(raw[ofs:ofs + ln]).hexdigest()),
                                fast_lump_md5=np.array(hashlib.md5(fast[fofs:fofs + fln]).hexdigest()),
                                fast_file_md5=np.array(hashlib.md5(fast).hexdigest()))
            print(name, "portals", d["P"], "clusters", d["nc"], "mean mightsee", d["might"].mean())


if __name__ == "__main__":
    main()
