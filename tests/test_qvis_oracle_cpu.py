"""Groundwork for the next stage outward (qvis3): the oracle comes first.

The reference qvis3 itself, compiled in place (`make -C oracle ref`), is the oracle; oracle/qvis_harness.c drives its
translation units and dumps the per-portal sets.  These tests need /root/reference-built tools (oracle/_ref) and are
skipped where they are absent.  They pin
  * the harness: it writes the visibility lump the unmodified reference qvis3 writes (full vis, and -fast);
  * the numpy restatement of the order-independent first stage (oracle/qvis_stage1.py): planes, portalfront, portalflood,
    nummightsee and the sorted order equal the reference's, bit for bit;
  * the fact the design of a GPU PortalFlow must start from: the reference's result depends on the portal order;
  * the C restatement of PortalFlow (oracle/qvis_oracle.c) and the schedule rule it is built on -- "portal j counts as
    finished for portal k iff j comes before k in the sorted order": every portal, computed on its own from the
    reference's finished vectors, and the whole run in sorted order, reproduce the reference's portalvis bit for bit.
"""
from __future__ import annotations

import struct
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT))
REF = ROOT / "oracle" / "_ref"
pytestmark = pytest.mark.skipif(not (REF / "qvis_harness").exists() or not (REF / "qbsp3").exists(),
                                reason="reference tools not built (make -C oracle ref needs /root/reference)")


def vis_lump(bsp: Path) -> bytes:
    d = bsp.read_bytes()
    ofs, ln = struct.unpack_from("<ii", d, 8 + 3 * 8)
    return d[ofs:ofs + ln]


def compile_bsp(tmp: Path, name: str, text: str):
    from qengine_b200 import mapgen
    assets = mapgen.write_assets(tmp)
    mp = assets / "maps" / (name + ".map")
    mp.write_text(text)
    r = subprocess.run([str(REF / "qbsp3"), str(mp)], capture_output=True, text=True)
    assert r.returncode == 0 and "LEAKED" not in r.stdout, r.stdout[-500:]
    return mp.with_suffix(".bsp"), mp.with_suffix(".prt")


def maps():
    from qengine_b200 import mapgen
    sys.path.insert(0, str(ROOT / "tests" / "golden"))
    import make_golden as MG
    return {"grid3": mapgen.room_grid_map(3, 3, seed=1, clutter=0.5, lamps=0.3),
            "grid5": mapgen.room_grid_map(5, 5, seed=2, clutter=1.0, lamps=0.3),
            "angled": MG.angled_map_text()}          # brushes with non-axial faces: general separating planes


@pytest.mark.parametrize("name", ["grid3", "grid5"])
@pytest.mark.parametrize("flags", [[], ["-fast"]])
def test_harness_writes_the_reference_lump(name, flags, tmp_path):
    bsp, prt = compile_bsp(tmp_path, name, maps()[name])
    pre = bsp.read_bytes()
    r = subprocess.run([str(REF / "qvis3")] + flags + [str(bsp)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-500:]
    want = vis_lump(bsp)
    bsp.write_bytes(pre)
    r = subprocess.run([str(REF / "qvis_harness")] + flags + [str(bsp), str(prt), str(tmp_path / "dump.bin")], capture_output=True, text=True)
    assert r.returncode == 0 and "QVISDUMP" in r.stdout, r.stdout[-500:]
    assert vis_lump(bsp) == want


@pytest.mark.parametrize("name", ["grid3", "grid5"])
def test_stage1_restatement_matches_the_reference(name, tmp_path):
    import qvis_stage1 as V
    bsp, prt = compile_bsp(tmp_path, name, maps()[name])
    r = subprocess.run([str(REF / "qvis_harness"), str(bsp), str(prt), str(tmp_path / "dump.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-500:]
    ref = V.read_dump(tmp_path / "dump.bin")
    pt = V.Portals(prt)
    assert pt.P == ref["P"] and pt.nc == ref["nc"]
    assert np.array_equal(pt.normal, ref["plane"][:, :3]) and np.array_equal(pt.dist, ref["plane"][:, 3])
    assert np.array_equal(pt.leaf, ref["leaf"])
    front = V.base_portal_vis(pt)
    assert np.array_equal(front, ref["front"])
    flood = V.simple_flood(pt, front)
    assert np.array_equal(flood, ref["flood"])
    might = flood.sum(1)
    assert np.array_equal(might, ref["might"])
    order, rank = V.sorted_rank(might)
    assert np.array_equal(rank, ref["rank"])
    # the final sets lie inside the flood sets, and a portal never sees itself
    assert not (ref["vis"] & ~ref["flood"]).any()
    assert not ref["vis"][np.arange(pt.P), np.arange(pt.P)].any()
    # every finished portal a portal may consult comes earlier in the order: the levels are a valid schedule
    level = V.dependency_levels(flood, order, rank)
    k, j = np.nonzero(flood & (rank[None, :] < rank[:, None]))
    assert (level[j] < level[k]).all()


def test_portal_order_changes_the_result(tmp_path):
    """-nosort (qvis3.c:127) processes the portals in file order: a different lump.  A parallel PortalFlow has to
    reproduce the sorted order's `finished` relation, not just any order."""
    bsp, prt = compile_bsp(tmp_path, "grid5", maps()["grid5"])
    pre = bsp.read_bytes()
    lumps = []
    for flags in ([], ["-nosort"]):
        bsp.write_bytes(pre)
        r = subprocess.run([str(REF / "qvis3")] + flags + [str(bsp)], capture_output=True, text=True)
        assert r.returncode == 0
        lumps.append(vis_lump(bsp))
    assert lumps[0] != lumps[1]


@pytest.mark.parametrize("name", ["grid3", "grid5", "angled"])
def test_portal_flow_restatement_and_schedule_rule(name, tmp_path):
    import qvis_stage1 as V
    if not (ROOT / "oracle" / "libqvis_oracle.so").exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "libqvis_oracle.so"], check=True, capture_output=True)
    bsp, prt = compile_bsp(tmp_path, name, maps()[name])
    r = subprocess.run([str(REF / "qvis_harness"), str(bsp), str(prt), str(tmp_path / "dump.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-500:]
    ref = V.read_dump(tmp_path / "dump.bin")
    pt = V.Portals(prt)
    flood = V.simple_flood(pt, V.base_portal_vis(pt))
    order, rank = V.sorted_rank(flood.sum(1))
    assert np.array_equal(rank, ref["rank"])
    fo = V.FlowOracle(pt, flood, rank)
    want = V.pack_bits(ref["vis"])
    # every portal on its own, with the reference's vectors for the portals ranked before it
    for k in range(pt.P):
        got, chains = fo.portal_flow(k, want)
        assert np.array_equal(got, want[k]), (name, k)
        assert chains >= 1
    # and the whole run in sorted order, from nothing
    assert np.array_equal(fo.full(order), want)
