"""Pins the CPU oracle (oracle/qrad_oracle.c, a plain-C restatement) to the reference: it must reproduce the
golden vectors in tests/golden/ -- which are outputs of the reference qrad3 itself -- byte for byte."""
from __future__ import annotations

import sys

import numpy as np
import pytest

import golden_util as G

sys.path.insert(0, str(G.ROOT / "oracle"))
import oracle as O  # noqa: E402


def make(case, tmp_path):
    meta, flags = G.load_case(case)
    bsp = G.stage_case(case, tmp_path)
    kw = dict(numbounce=8, extrasamples=0, subdiv=64.0, ambient=0.0, maxlight=196.0, lightscale=1.0)
    it = iter(flags)
    for f in it:
        if f == "-bounce":
            kw["numbounce"] = int(next(it))
        elif f == "-extra":
            kw["extrasamples"] = 1
        elif f == "-chop":
            kw["subdiv"] = float(int(next(it)))
        elif f == "-ambient":
            kw["ambient"] = float(np.float32(float(next(it)) * 128))
        elif f == "-scale":
            kw["lightscale"] = float(np.float32(float(next(it))))
        elif f == "-maxlight":
            kw["maxlight"] = float(np.float32(min(float(next(it)) * 128, 255)))
    return meta, O.Oracle(bsp, **kw)


@pytest.mark.parametrize("case", ["samplec", "samplec_extra", "samplec_direct", "grid2", "styles", "manystyles", "nolights", "angled"])
def test_oracle_reproduces_reference_lump(case, tmp_path):
    meta, o = make(case, tmp_path)
    lump, added = o.rad_world()
    want = (G.GOLDEN / (case + ".lump")).read_bytes()
    assert lump.tobytes() == want
    assert np.array_equal(added, np.array(meta["added"], np.float32))
    assert o.L.oq_num_patches() == meta["num_patches"]
    if meta["total_transfer"]:
        assert o.L.oq_total_transfer() == meta["total_transfer"]
        assert np.array_equal(o.patches()["numtransfers"], G.golden_numtransfers(case))


def test_oracle_transfer_rows_and_rays(tmp_path):
    meta, o = make("samplec", tmp_path)
    o.rad_world()
    z = np.load(G.GOLDEN / "samplec.transfers.npz")
    patch, tr = o.transfers()
    assert np.array_equal(patch, z["patch"].astype(np.uint32)) and np.array_equal(tr, z["transfer"])
    assert np.array_equal(o.patches()["totallight"], np.load(G.GOLDEN / "samplec.totallight.npy"))
    assert np.array_equal(o.facelight_style0(), np.load(G.GOLDEN / "samplec.facelight0.npy"))
    rays = G.golden_rays("samplec")
    assert np.array_equal(o.test_lines(rays["start"], rays["stop"]) != 0, rays["result"] != 0)
    # SURVEY.md appendix B ray counts: 8 380 (BuildFacelights) + 38 414 (MakeTransfers)
    assert o.ray_counts() == (8380, 38414)


@pytest.mark.parametrize("case", ["grid3_chop16", "hall"])
def test_oracle_testline_golden(case, tmp_path):
    meta, o = make(case, tmp_path)
    rays = G.golden_rays(case)
    assert np.array_equal(o.test_lines(rays["start"], rays["stop"]) != 0, rays["result"] != 0)
