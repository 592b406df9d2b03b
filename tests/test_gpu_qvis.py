"""qvis3 first stage on the GPU (k_portal_front, k_simple_flood) against the reference's own per-portal vectors
(tests/golden/vis_*.npz, made by oracle/qvis_harness.c over the reference's translation units)."""
from __future__ import annotations

import gzip
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("case", ["grid3", "grid5", "angled"])
def test_base_portal_vis_is_bit_identical(case, tmp_path):
    from qengine_b200 import qvis
    prt = tmp_path / (case + ".prt")
    prt.write_bytes(gzip.decompress((GOLDEN / f"vis_{case}.prt.gz").read_bytes()))
    want = np.load(GOLDEN / f"vis_{case}.npz")
    pt = qvis.Portals(prt)
    front, flood, might, ms = pt.base_portal_vis(0)
    assert np.array_equal(front, want["front"]), "portalfront (flow.c:623-644)"
    assert np.array_equal(flood, want["flood"]), "portalflood (SimpleFlood, flow.c:576-598)"
    assert np.array_equal(might, want["might"]), "nummightsee"
    # the order SortPortals derives from it (stable sort by nummightsee) is the reference's
    order = np.argsort(might, kind="stable")
    rank = np.empty(len(might), np.int64)
    rank[order] = np.arange(len(might))
    assert np.array_equal(rank, want["rank"])


def test_base_portal_vis_on_the_c5_map(tmp_path):
    """The map of BASELINE configs[4] (5 608 file portals, 2 211 clusters): digests of the reference's vectors."""
    import hashlib
    from qengine_b200 import qvis
    root = Path(__file__).resolve().parent.parent
    prt = tmp_path / "grid16c128.prt"
    prt.write_bytes(gzip.decompress((root / "bench_maps" / "grid16c128.prt.gz").read_bytes()))
    want = np.load(GOLDEN / "vis_c5_grid16.npz")
    pt = qvis.Portals(prt)
    assert hashlib.md5(pt.arrays()["planes"].tobytes()).hexdigest() == str(want["planes_md5"])
    front, flood, might, ms = pt.base_portal_vis(0)
    assert hashlib.md5(front.tobytes()).hexdigest() == str(want["front_md5"])
    assert hashlib.md5(flood.tobytes()).hexdigest() == str(want["flood_md5"])
    assert np.array_equal(might, want["might"])
    print(f"BasePortalVis on the C5 map: k_portal_front {ms[0]:.2f} ms, k_simple_flood {ms[1]:.2f} ms")


@pytest.mark.parametrize("case", ["grid3", "grid5", "angled"])
def test_fast_vis_writes_the_reference_file(case, tmp_path):
    """`qvis3 -fast` end to end: qbsp3's BSP + portal file in, the file the unmodified reference tool writes out."""
    import hashlib
    from qengine_b200 import qvis
    want = np.load(GOLDEN / f"vis_{case}.npz")
    bsp = tmp_path / (case + ".bsp")
    bsp.write_bytes(gzip.decompress((GOLDEN / f"vis_{case}.bsp.gz").read_bytes()))
    prt = tmp_path / (case + ".prt")
    prt.write_bytes(gzip.decompress((GOLDEN / f"vis_{case}.prt.gz").read_bytes()))
    qvis.fast_vis(bsp, prt, bsp, 0)
    assert hashlib.md5(bsp.read_bytes()).hexdigest() == str(want["fast_file_md5"])
