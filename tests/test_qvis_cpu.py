"""Host half of the qvis3 first stage (include/qvis_cuda.h) against fixtures made by the reference
(tests/golden/make_golden_vis.py): LoadPortals / PlaneFromWinding, and the boundary's error behaviour without a GPU."""
from __future__ import annotations

import gzip
from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).resolve().parent / "golden"
CASES = ["grid3", "grid5", "angled"]


def stage(case, tmp: Path) -> Path:
    prt = tmp / (case + ".prt")
    prt.write_bytes(gzip.decompress((GOLDEN / f"vis_{case}.prt.gz").read_bytes()))
    return prt


@pytest.mark.parametrize("case", CASES)
def test_load_portals_matches_the_reference(case, tmp_path):
    from qengine_b200 import qvis
    want = np.load(GOLDEN / f"vis_{case}.npz")
    pt = qvis.Portals(stage(case, tmp_path))
    a = pt.arrays()
    assert pt.numportals == len(want["leaf"]) and pt.portalbytes == want["front"].shape[1]
    assert np.array_equal(a["planes"], want["planes"])          # bit-identical floats
    assert np.array_equal(a["leaf"], want["leaf"])
    # two memory portals per file portal: the second is the first one reversed, with the opposite plane
    wf, pts = a["winding_first"], a["points"]
    for k in range(0, pt.numportals, 2):
        f, b = pts[wf[k]:wf[k + 1]], pts[wf[k + 1]:wf[k + 2]]
        assert np.array_equal(f, b[::-1])
        assert np.array_equal(a["planes"][k], -a["planes"][k + 1])
    # every portal leaves exactly one cluster: the one its twin leads into
    owner = np.full(pt.numportals, -1)
    for l in range(pt.numclusters):
        owner[a["leaf_portals"][a["leaf_first"][l]:a["leaf_first"][l + 1]]] = l
    assert np.array_equal(owner[0::2], a["leaf"][1::2]) and np.array_equal(owner[1::2], a["leaf"][0::2])


def test_load_portals_errors(tmp_path):
    from qengine_b200 import qvis
    with pytest.raises(qvis.QvisError, match="couldn't read"):
        qvis.Portals(tmp_path / "missing.prt")
    bad = tmp_path / "bad.prt"
    bad.write_text("PRT2\n1\n0\n")
    with pytest.raises(qvis.QvisError, match="not a portal file"):
        qvis.Portals(bad)
    bad.write_text("PRT1\n2\n1\n4 0 5 (0 0 0 ) (0 1 0 ) (1 1 0 ) (1 0 0 ) \n")
    with pytest.raises(qvis.QvisError, match="reading portal 0"):
        qvis.Portals(bad)


def test_no_cpu_fallback(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from qengine_b200 import qvis
    pt = qvis.Portals(stage("grid3", tmp_path))
    with pytest.raises(qvis.QvisError, match="no CPU fallback"):
        pt.base_portal_vis(0)


@pytest.mark.parametrize("case", CASES)
def test_merge_lump_matches_the_reference(case, tmp_path):
    """ClusterMerge + CalcPHS + CompressVis on the host: from the reference's FINAL vectors the lump the reference wrote
    (full vis), from its flood vectors the lump of `qvis3 -fast` (fastvis takes portalflood for portalvis, qvis3.c:235)."""
    import hashlib
    from qengine_b200 import qvis
    want = np.load(GOLDEN / f"vis_{case}.npz")
    pt = qvis.Portals(stage(case, tmp_path))
    lump, vis, hear, warned = pt.merge_lump(want["vis"])
    assert hashlib.md5(lump).hexdigest() == str(want["lump_md5"])
    assert 0 < vis <= hear <= pt.numclusters
    lump, vis2, hear2, _ = pt.merge_lump(want["flood"])
    assert hashlib.md5(lump).hexdigest() == str(want["fast_lump_md5"])
    assert vis2 >= vis
    with pytest.raises(qvis.QvisError, match="Vismap expansion overflow"):
        pt.merge_lump(want["vis"], cap=4 + 8 * pt.numclusters + 16)


def test_qvis3_cli_errors(tmp_path):
    """The drop-in executable speaks the reference's command line (qvis3.c:493-566): banner, usage, unknown options; it
    stops without -fast (PortalFlow is not built) and, with -fast but no GPU, says there is no CPU fallback."""
    import subprocess
    import torch
    exe = Path(__file__).resolve().parent.parent / "qengine_b200" / "qvis3"
    if not exe.exists():
        pytest.skip("qvis3 executable not built")
    def run(*a):
        return subprocess.run([str(exe), *a], capture_output=True, text=True, cwd=tmp_path)
    r = run("-fast")
    assert r.returncode == 1 and "---- vis ----" in r.stdout and "usage: vis" in r.stdout
    r = run("-bogus", "x.bsp")
    assert r.returncode == 1 and 'Unknown option "-bogus"' in r.stdout
    r = run("x.bsp")
    assert r.returncode == 1 and "PortalFlow is not built" in r.stdout
    r = run("-fast", "missing.bsp")
    assert r.returncode == 1 and "************ ERROR ************" in r.stdout
    if not torch.cuda.is_available():
        import gzip as gz
        (tmp_path / "grid3.bsp").write_bytes(gz.decompress((GOLDEN / "vis_grid3.bsp.gz").read_bytes()))
        (tmp_path / "grid3.prt").write_bytes(gz.decompress((GOLDEN / "vis_grid3.prt.gz").read_bytes()))
        r = run("-fast", "grid3.bsp")
        assert r.returncode == 1 and "no CPU fallback" in r.stdout and "121 numportals" in r.stdout
