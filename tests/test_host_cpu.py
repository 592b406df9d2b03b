"""CPU-side tests: the C-ABI library loads and exports what include/*.h declares, and the host half
(BSP io, entities, reflectivity, patches, dicing, light lists) reproduces the reference's outputs.
No CUDA compute is called here."""
from __future__ import annotations

import ctypes
import hashlib
import re
import shutil
import struct
from pathlib import Path

import numpy as np
import pytest

import golden_util as G
import refdump as R

ROOT = G.ROOT


@pytest.fixture(scope="module")
def Q():
    from qengine_b200 import build
    build.build()
    from qengine_b200 import qrad
    qrad.lib()
    return qrad


def declared_symbols():
    names = set()
    for h in (ROOT / "include").glob("*.h"):
        txt = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        for m in re.finditer(r"^\s*(?:const\s+)?(?:int|void|char|int64_t)\s*\*?\s*(q(?:rad|vis)\w+)\s*\(", txt, flags=re.M):
            names.add(m.group(1))
    return sorted(names)


def test_library_exports_every_declared_symbol(Q):
    lib = Q.lib()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ but not exported by libqrad_cuda.so"
    assert lib.qrad_cuda_abi_version() == 2


def test_no_cpu_fallback_without_gpu(Q):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(Q.QradError, match="no CPU fallback"):
        Q.Device(0)


@pytest.mark.parametrize("case", ["samplec", "grid2", "hall", "styles"])
def test_bsp_roundtrip_is_byte_identical(Q, case, tmp_path):
    """WriteBSPFile's fixed lump order and padding (bspfile.c:466-503): re-writing an unlit file changes nothing."""
    bsp = G.stage_case(case, tmp_path)
    before = bsp.read_bytes()
    sc = Q.Scene(bsp)
    out = tmp_path / "assets" / "maps" / "rewritten.bsp"
    sc.write(out)
    assert out.read_bytes() == before


def test_load_errors_are_reported(Q, tmp_path):
    with pytest.raises(Q.QradError, match="no 'assets'"):
        Q.Scene(tmp_path / "nowhere.bsp")
    maps = tmp_path / "assets" / "maps"
    maps.mkdir(parents=True)
    (maps / "junk.bsp").write_bytes(b"JUNK" + bytes(200))
    with pytest.raises(Q.QradError, match="not a IBSP file"):
        Q.Scene(maps / "junk.bsp")
    bad = bytearray((G.GOLDEN / "assets" / "maps" / "sample.bsp").read_bytes())
    struct.pack_into("<i", bad, 4, 46)
    (maps / "v46.bsp").write_bytes(bytes(bad))
    with pytest.raises(Q.QradError, match="is version 46, not 38"):
        Q.Scene(maps / "v46.bsp")
    (maps / "ok.bsp").write_bytes((G.GOLDEN / "assets" / "maps" / "sample.bsp").read_bytes())
    with pytest.raises(Q.QradError, match="colormap.pcx"):
        Q.Scene(maps / "ok.bsp")   # palette is required (patches.c:51-54)


def test_sample_counts(Q, tmp_path):
    """38 faces, 238 patches after subdivision, 2 direct lights, 7 clusters, 35 nodes."""
    bsp = G.stage_case("samplec", tmp_path)
    sc = Q.Scene(bsp).prepare(G.params_from_flags(["-bounce", "8"]))
    c = sc.counts()
    assert (c["numfaces"], c["num_patches"], c["numdlights"], c["numclusters"], c["numnodes"], c["numleafs"]) == (38, 238, 2, 7, 35, 37)
    p = sc.patches()
    assert (p["area"] >= 1).all()
    # every patch is on exactly one face list
    seen = np.zeros(c["num_patches"], int)
    for h in p["face_head"]:
        while h >= 0:
            seen[h] += 1
            h = p["next"][h]
    assert (seen == 1).all()


def test_chop_controls_patch_count(Q, tmp_path):
    bsp = G.stage_case("samplec", tmp_path)
    n = {}
    for chop in (0, 16, 64, 128):
        sc = Q.Scene(bsp).prepare(G.params_from_flags(["-chop", str(chop)]))
        n[chop] = sc.counts()["num_patches"]
    assert n[0] == 38            # subdiv < 1 disables dicing (patches.c:481-482): one patch per face
    assert n[16] > n[64] >= n[128] >= 38
    assert n[16] == 2338         # sample.map -chop 16


def test_missing_texture_falls_back_to_half_grey(Q, tmp_path):
    bsp = G.stage_case("samplec", tmp_path)
    shutil.rmtree(tmp_path / "assets" / "textures")
    sc = Q.Scene(bsp)
    assert np.array_equal(sc.reflectivity(), np.full((3, 3), 0.5, np.float32))


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref (reference-built harness) not present")
@pytest.mark.parametrize("case,flags", [
    ("samplec", ["-bounce", "8"]), ("grid2", ["-bounce", "8", "-chop", "32"]), ("grid3_chop16", ["-chop", "16"]),
    ("hall", ["-bounce", "2"]), ("styles", ["-bounce", "3", "-direct", "2", "-entity", "0.5"]),
])
def test_host_setup_matches_reference_dumps(Q, case, flags, tmp_path):
    """Patches, windings, clusters, reflectivity and the per-cluster light lists against the reference's globals."""
    bsp = G.stage_case(case, tmp_path)
    unlit = bsp.read_bytes()
    R.run_ref_harness(bsp, flags, dumpdir=tmp_path / "dump", stopafter=0)
    bsp.write_bytes(unlit)
    sc = Q.Scene(bsp).prepare(G.params_from_flags(flags))
    rp, heads = R.read_patches(tmp_path / "dump" / "patches_pre.bin")
    mp = sc.patches()
    assert len(rp) == sc.counts()["num_patches"]
    assert np.array_equal(R.read_reflectivity(tmp_path / "dump" / "reflectivity.bin"), sc.reflectivity())
    for k in ("origin", "normal", "area", "cluster", "reflectivity", "baselight", "face", "next"):
        assert np.array_equal(mp[k], rp[k]), k
    assert np.array_equal(mp["sky"].astype(np.int32), rp["sky"])
    assert np.array_equal(mp["bounds"][:, :3], rp["mins"]) and np.array_equal(mp["bounds"][:, 3:], rp["maxs"])
    assert np.array_equal(mp["face_head"], heads)
    cnt, pts = sc.windings()
    rw = R.read_windings(tmp_path / "dump" / "windings.bin")
    assert np.array_equal(cnt, [len(w) for w in rw]) and np.array_equal(pts, np.concatenate(rw))
    rl = R.read_lights(tmp_path / "dump" / "lights.bin")
    ml = sc.lights()
    assert np.array_equal(np.repeat(np.arange(len(ml["light_first"]) - 1), np.diff(ml["light_first"])), rl["cluster"])
    for k in ("type", "intensity", "style", "origin", "color", "normal", "stopdot"):
        assert np.array_equal(ml[k], rl[k]), k


@pytest.mark.parametrize("parts", [1, 2, 3, 5])
@pytest.mark.parametrize("case", ["samplec", "grid3_chop16", "manystyles", "angled"])
def test_shared_prepare_gives_the_same_scene(Q, case, parts, tmp_path):
    """Several processes that light one map share the dicing (qrad_host_prepare_part / _merge): whatever the number of
    parts, every process must end with the patches, windings and lights of a plain prepare -- itself pinned to the
    reference's dumps above."""
    meta, flags = G.load_case(case)
    bsp = G.stage_case(case, tmp_path)
    want = Q.Scene(bsp).prepare(G.params_from_flags(flags))
    scenes = [Q.Scene(bsp) for _ in range(parts)]
    blobs = [sc.prepare_part(G.params_from_flags(flags), k, parts) for k, sc in enumerate(scenes)]
    assert sum(len(b) for b in blobs) > 0
    for sc in scenes[:2]:   # every process merges all blobs; two of them are enough here
        sc.prepare_merge(blobs)
        assert sc.counts() == want.counts()
        a, b = sc.patches(), want.patches()
        for k in b:
            assert np.array_equal(a[k], b[k]), k
        la, lb = sc.lights(), want.lights()
        for k in lb:
            assert np.array_equal(la[k], lb[k]), k
        (ca, pa), (cb, pb) = sc.windings(), want.windings()
        assert np.array_equal(ca, cb) and np.array_equal(pa, pb)


def test_shared_prepare_rejects_bad_parts(Q, tmp_path):
    meta, flags = G.load_case("samplec")
    bsp = G.stage_case("samplec", tmp_path)
    scenes = [Q.Scene(bsp) for _ in range(2)]
    blobs = [sc.prepare_part(G.params_from_flags(flags), k, 2) for k, sc in enumerate(scenes)]
    with pytest.raises(Q.QradError):
        scenes[0].prepare_merge([blobs[0], blobs[0]])          # a range twice, another missing
    scenes[0].prepare_part(G.params_from_flags(flags), 0, 2)
    with pytest.raises(Q.QradError):
        scenes[0].prepare_merge([blobs[0], blobs[1][:-4]])      # truncated
    with pytest.raises(Q.QradError):
        Q.Scene(bsp).prepare_merge(blobs)                       # prepare_part has not run
