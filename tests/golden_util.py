"""Helpers shared by the tests: golden cases, staging a .bsp under an assets/ tree."""
from __future__ import annotations

import json
import shutil
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
CASES = ["samplec", "samplec_extra", "samplec_direct", "grid2", "grid3_chop16", "hall", "styles", "manystyles", "nolights", "angled"]

RAY_DTYPE = np.dtype([("start", "<f4", 3), ("stop", "<f4", 3), ("result", "<i4")])


def load_case(name):
    meta = json.loads((GOLDEN / (name + ".json")).read_text())
    return meta, list(meta["flags"])


def stage_case(name, tmp: Path) -> Path:
    """Copy the case's unlit .bsp + the synthetic pics/textures into tmp/assets/...; returns the .bsp path."""
    meta, _ = load_case(name)
    src = GOLDEN / "assets"
    dst = Path(tmp) / "assets"
    (dst / "maps").mkdir(parents=True, exist_ok=True)
    for sub in ("pics", "textures"):
        if not (dst / sub).exists():
            shutil.copytree(src / sub, dst / sub)
    out = dst / "maps" / meta["bsp"]
    shutil.copyfile(src / "maps" / meta["bsp"], out)
    return out


def params_from_flags(flags):
    """The reference's flag parser (qrad3.c:555-601) as the library implements it."""
    from qengine_b200 import qrad as Q
    return Q.params_from_flags(flags)


def golden_rays(name):
    return np.fromfile(GOLDEN / (name + ".rays"), dtype=RAY_DTYPE)


def golden_rayleafs(name):
    """(points, PointInLeafnum of the reference) -- the point is the ray's end, snapped to the 16-grid in x for odd rays
    of the harness's own numbering (oracle/ref_harness.c dump_rays); the committed sample keeps every k-th ray."""
    rays = golden_rays(name)
    leafs = np.fromfile(GOLDEN / (name + ".rayleafs"), dtype="<i4")
    meta = json.loads((GOLDEN / (name + ".json")).read_text())
    stride = meta["rays_stride"]
    pts = rays["stop"].copy()
    odd = (np.arange(len(rays)) * stride) & 1 == 1
    pts[odd, 0] = (16 * (pts[odd, 0] / np.float32(16)).astype(np.int32)).astype(np.float32)
    return pts, leafs


def golden_numtransfers(name):
    return np.fromfile(GOLDEN / (name + ".numtransfers"), dtype="<i4")
