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


@pytest.mark.parametrize("case", ["samplec", "grid2", "grid3_chop16", "hall", "angled"])
def test_oracle_point_in_leaf_golden(case, tmp_path):
    """PointInLeafnum (qrad3.c:131-150) of the reference for points next to / on node planes."""
    meta, o = make(case, tmp_path)
    pts, leafs = G.golden_rayleafs(case)
    assert np.array_equal(o.point_in_leaf(pts), leafs)


@pytest.mark.parametrize("case", ["samplec", "samplec_extra", "grid2", "styles", "angled"])
def test_oracle_ray_counts_match_reference(case, tmp_path):
    """Unit of the rays/s metric: one root TestLine_r call.  The reference's own count (ld --wrap in oracle/ref_harness.c)
    is in the golden metadata."""
    meta, o = make(case, tmp_path)
    o.rad_world()
    assert o.ray_counts() == (meta["rays_direct"], meta["rays_transfers"])


def test_limit_lifted_harness_is_byte_identical_to_the_unmodified_reference(tmp_path):
    """oracle/_ref/ref_harness_big (MAX_PATCHES / u16 patch index lifted by sed on a scratch copy, oracle/Makefile)
    writes the same file and counts the same rays as the unmodified reference on every golden case."""
    import hashlib
    import json
    import subprocess
    import refdump as R
    if not (R.REF_DIR / "ref_harness_big").exists():
        pytest.skip("oracle/_ref (reference-built harnesses) not present")
    for case in G.CASES:
        meta, flags = G.load_case(case)
        got = {}
        for exe in ("ref_harness", "ref_harness_big"):
            bsp = G.stage_case(case, tmp_path / exe)
            r = subprocess.run([str(R.REF_DIR / exe), *flags, str(bsp)], capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
            info = json.loads([l for l in r.stdout.splitlines() if l.startswith("REFTIMING ")][0][10:])
            got[exe] = (hashlib.md5(bsp.read_bytes()).hexdigest(), info["rays_direct"], info["rays_transfers"],
                        info["total_transfer"], info["added"])
        assert got["ref_harness"] == got["ref_harness_big"], case
        assert got["ref_harness"][0] == meta["file_md5"], case
