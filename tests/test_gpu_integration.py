"""The reference qrad3 itself, with RadWorld() bound to libqrad_cuda by the stub of INTEGRATION.md
(integration/radworld_cuda.c + faceview_cuda.c, built by integration/Makefile from the reference sources in place):
the reference's own main(), flag parser, BSP loader, MakePatches / SubdividePatches / CreateDirectLights and
WriteBSPFile around the GPU library must write the same files as the unmodified reference."""
from __future__ import annotations

import hashlib
import subprocess

import pytest

import golden_util as G

pytestmark = pytest.mark.gpu
EXE = G.ROOT / "integration" / "_build" / "qrad3_cuda"


@pytest.mark.skipif(not EXE.exists(), reason="integration/_build/qrad3_cuda not built (needs /root/reference at build time)")
@pytest.mark.parametrize("case", ["samplec", "samplec_extra", "grid3_chop16", "styles", "manystyles", "angled"])
def test_reference_qrad3_with_cuda_radworld(case, tmp_path):
    meta, flags = G.load_case(case)
    bsp = G.stage_case(case, tmp_path)
    r = subprocess.run([str(EXE), *flags, str(bsp)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]
    assert "----- Radiosity ----" in r.stdout
    assert hashlib.md5(bsp.read_bytes()).hexdigest() == meta["file_md5"]
