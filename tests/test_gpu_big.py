"""GPU parity on the LARGE BASELINE configs (C3, C4, C5) against fixtures the reference itself produced
(tests/golden/make_golden_big.py: unmodified reference harness for C4; the limit-lifted harness, byte-identical to the
unmodified one on every small golden, for C3 and the row / column samples of C5).

C5 (1 006 616 patches, 5.5e9 transfers) cannot be lit by the reference in finite CPU time, so it is pinned through
  * a deterministic 1/256 sample of the source rows: numtransfers, every (patch, weight) record (64-bit digest per
    row) and the root TestLine_r count of each row, bit-exact;
  * a window of 64 receivers for which the reference ran EVERY source that can see them: totallight and radiosity
    after the first bounce, bit-exact -- this exercises the transposed (gather) matrix and k_bounce on the big path.
"""
from __future__ import annotations

import gzip
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

import golden_util as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Q():
    from qengine_b200 import qrad
    qrad.lib()
    return qrad


def _row_digest(patch, transfer) -> int:
    h = hashlib.blake2b(digest_size=8)
    h.update(np.ascontiguousarray(patch, "<u4").tobytes())
    h.update(np.ascontiguousarray(transfer, "<u2").tobytes())
    return int.from_bytes(h.digest(), "little")


def _light_workload(Q, name, tmp_path, flags=None):
    from qengine_b200 import workloads as W
    bsp = W.stage_workload(name, tmp_path)
    p = Q.params_from_flags(flags or W.WORKLOADS[name][1])
    sc = Q.Scene(bsp).prepare(p)
    dev = Q.Device(0)
    added = Q.rad_world(sc, dev, p)
    return sc, dev, added, bsp


def _check_full(Q, golden, tmp_path):
    meta = json.loads((G.GOLDEN / (golden + ".json")).read_text())
    sc, dev, added, bsp = _light_workload(Q, meta["workload"], tmp_path, meta["flags"])
    want = np.frombuffer(gzip.open(G.GOLDEN / (golden + ".lump.gz")).read(), np.uint8)
    lump = np.frombuffer(sc.lighting(), np.uint8)
    assert lump.size == want.size == meta["lightdatasize"]
    d = np.abs(lump.astype(int) - want.astype(int))
    frac_ok = float((d <= 1).mean())
    assert frac_ok >= 0.999, f"only {frac_ok:.5f} within +-1, max deviation {d.max()}"      # the north-star tolerance
    assert int(d.max()) == 0, f"{int((d > 0).sum())} of {d.size} bytes differ, max deviation {int(d.max())}"
    assert np.array_equal(added, np.array(meta["added"], np.float32))
    st = dev.stats()
    assert st["num_patches"] == meta["num_patches"]
    assert (st["rays_direct"], st["rays_transfers"]) == (meta["rays_direct"], meta["rays_transfers"])
    if meta["total_transfer"]:
        nt = dev.numtransfers()
        want_nt = np.frombuffer(gzip.open(G.GOLDEN / (golden + ".numtransfers.gz")).read(), "<i4")
        assert np.array_equal(nt, want_nt)
        assert int(nt.astype(np.int64).sum()) == meta["numtransfers_sum"]
        tl = dev.patch_light()["totallight"]
        assert hashlib.md5(np.ascontiguousarray(tl, "<f4").tobytes()).hexdigest() == meta["totallight_md5"]
    sc.write(bsp)
    assert hashlib.md5(bsp.read_bytes()).hexdigest() == meta["file_md5"]


def test_c4_hall2000_direct_light_identical(Q, tmp_path):
    """BASELINE configs[3]: 2 000 light / light_spot entities, -bounce 0 -extra (305 M shadow rays)."""
    _check_full(Q, "c4_hall2000", tmp_path)


def test_c4_hall2000_two_bounces_identical(Q, tmp_path):
    _check_full(Q, "c4_hall2000_b2", tmp_path)


def test_c3_grid13_identical(Q, tmp_path):
    """BASELINE configs[2]: 149 021 patches (> 65 535: u32 receiver indices), -chop 32, 8 bounces."""
    _check_full(Q, "c3_grid13", tmp_path)


def test_c5_row_sample_and_column_window(Q, tmp_path):
    """BASELINE configs[4], the benchmarked map."""
    from qengine_b200 import workloads as W
    z = np.load(G.GOLDEN / "c5_rows.npz")
    zc = np.load(G.GOLDEN / "c5_cols.npz")
    name = "c5_grid16_chop8"
    bsp = W.stage_workload(name, tmp_path)
    p = Q.params_from_flags(W.WORKLOADS[name][1])
    sc = Q.Scene(bsp).prepare(p)
    assert sc.counts()["num_patches"] == 1006616
    dev = Q.Device(0).upload(sc)
    dev.build_facelights(p.extrasamples, p.numbounce)
    dev.make_transfers(p.nopvs)
    # ---- rows ----
    sel = z["patch"]
    assert sel[0] == z["first"] and sel[1] - sel[0] == z["stride"]
    nt = dev.numtransfers()
    assert np.array_equal(nt[sel], z["numtransfers"]), "numtransfers of the sampled rows"
    row_ptr, patch, transfer, rays = dev.rows(sel)
    assert np.array_equal(np.diff(row_ptr), z["numtransfers"])
    assert np.array_equal(rays, z["rays"]), "root TestLine_r calls per row"
    assert np.array_equal(patch[row_ptr[0]:row_ptr[1]], z["row0_patch"])
    assert np.array_equal(transfer[row_ptr[0]:row_ptr[1]], z["row0_transfer"])
    got = np.array([_row_digest(patch[row_ptr[k]:row_ptr[k + 1]], transfer[row_ptr[k]:row_ptr[k + 1]])
                    for k in range(sel.size)], "<u8")
    bad = np.nonzero(got != z["digest"])[0]
    assert bad.size == 0, f"{bad.size} of {sel.size} sampled rows differ, first: patch {sel[bad[0]]}"
    # ---- columns: one bounce, the 64 receivers whose every source the reference ran ----
    dev.bounce(1)
    pl = dev.patch_light()
    a = int(zc["first"])
    n = zc["totallight"].shape[0]
    assert np.array_equal(pl["totallight"][a:a + n], zc["totallight"]), "totallight after one bounce"
    assert np.array_equal(pl["radiosity"][a:a + n], zc["radiosity"]), "radiosity after one bounce"
    assert float(zc["totallight"].max()) > 0
