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
    """The reference's flag parser (qrad3.c:555-601) applied to a HostParams."""
    from qengine_b200 import qrad as Q
    p = Q.default_params()
    it = iter(flags)
    for f in it:
        if f == "-bounce":
            p.numbounce = int(next(it))
        elif f == "-extra":
            p.extrasamples = 1
        elif f == "-chop":
            p.subdiv = float(int(next(it)))
        elif f == "-scale":
            p.lightscale = float(next(it))
        elif f == "-direct":
            p.direct_scale = p.direct_scale * float(next(it))
        elif f == "-entity":
            p.entity_scale = p.entity_scale * float(next(it))
        elif f == "-nopvs":
            p.nopvs = 1
        elif f == "-ambient":
            p.ambient = float(next(it)) * 128
        elif f == "-maxlight":
            p.maxlight = min(float(next(it)) * 128, 255.0)
        elif f == "-threads":
            next(it)
        elif f == "-v":
            p.verbose = 1
        else:
            raise ValueError(f)
    return p


def golden_rays(name):
    return np.fromfile(GOLDEN / (name + ".rays"), dtype=RAY_DTYPE)


def golden_numtransfers(name):
    return np.fromfile(GOLDEN / (name + ".numtransfers"), dtype="<i4")
