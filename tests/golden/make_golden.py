#!/usr/bin/env python
"""Regenerates tests/golden/ from the REFERENCE run in this container.

Needs /root/reference (through `make -C oracle ref`, which compiles the reference qbsp3 / qvis3 /
qrad3 in place).  Every case is: a .map  --reference qbsp3 + qvis3-->  <case>.bsp (committed, the
input both sides light)  --reference qrad3 (oracle/_ref/ref_harness)-->  expected outputs:

  <case>.lump            the lighting lump bytes
  <case>.json            flags, counts, md5 of the lump / faces lump / whole output file, `added`
  <case>.numtransfers    int32 per patch
  <case>.rays            a sample of TestLine_r (start, stop, result) triples
  <case>.transfers.npz   (sample only) every transfer row: patch/transfer u16 arrays

The side inputs (palette, .wal textures) are the synthetic ones of qengine_b200.mapgen, written to
tests/golden/assets/.  The sample map text itself is not committed: it is read from
/root/reference/assets/maps/sample.map, every light gets an explicit `_color` (SURVEY.md finding F4).
When oracle/_ref/assets (the reference's real palette/textures) is present, the SURVEY appendix-B
md5 values are checked too.
"""
from __future__ import annotations

import hashlib
import json
import shutil
import struct
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))

import refdump as R  # noqa: E402
from qengine_b200 import mapgen  # noqa: E402

CASES = {
    # name: (map builder, qbsp chop, qrad3 flags)
    "samplec": ("sample", None, ["-bounce", "8"]),
    "samplec_extra": ("sample", None, ["-bounce", "8", "-extra"]),
    "samplec_direct": ("sample", None, ["-bounce", "0"]),
    "grid2": ("grid2", None, ["-bounce", "8", "-chop", "32"]),
    "grid3_chop16": ("grid3", None, ["-bounce", "4", "-chop", "16", "-extra"]),
    "hall": ("hall", None, ["-bounce", "2", "-extra"]),
    "styles": ("styles", None, ["-bounce", "3", "-ambient", "0.02", "-scale", "1.5", "-maxlight", "1.6"]),
    "manystyles": ("manystyles", None, ["-bounce", "2", "-chop", "48"]),
    "nolights": ("nolights", None, ["-bounce", "2"]),
    "angled": ("angled", None, ["-bounce", "4", "-chop", "24", "-extra"]),
}


def sample_map_text() -> str:
    src = Path("/root/reference/assets/maps/sample.map").read_text()
    return src.replace('"classname" "light"\n', '"classname" "light"\n"_color" "1 1 1"\n')


def styles_map_text() -> str:
    """2x1 rooms; lights with styles, a targeted spot, a bmodel with an origin brush and _minlight."""
    txt = mapgen.room_grid_map(2, 1, seed=7, clutter=1.0, lamps=1.0)
    extra = (
        '{\n"classname" "light"\n"origin" "-200 20 30"\n"light" "250"\n"_color" "1 0.6 0.3"\n"style" "3"\n}\n'
        '{\n"classname" "light"\n"origin" "120 -40 20"\n"_light" "180"\n"_color" "0.4 0.7 1"\n"_style" "7"\n}\n'
        '{\n"classname" "light"\n"origin" "100 60 40"\n"light" "300"\n"_color" "1 1 1"\n"target" "t1"\n"_cone" "25"\n}\n'
        '{\n"classname" "info_null"\n"targetname" "t1"\n"origin" "160 0 -60"\n}\n'
        # a brush model whose entity asks for a mottled minimum light (rand() path, lightmap.c:1187-1189)
        '{\n"classname" "func_wall"\n"_minlight" "0.35"\n' + mapgen.box_brush(-220, -100, -64, -196, -60, 0, mapgen.TEX_WALL) + '}\n'
        # a brush model with an origin brush: faces are offset into place (patches.c:214-223, 291-305)
        '{\n"classname" "func_rotating"\n' + mapgen.box_brush(40, 70, -64, 72, 102, -16, mapgen.TEX_FLOOR)
        + mapgen.box_brush(48, 78, -48, 64, 94, -32, mapgen.TEX_WALL, contents=0x1000000) + '}\n'
    )
    return txt + extra


def manystyles_map_text() -> str:
    """One room lit by seven lights with seven different styles: faces collect more than MAXLIGHTMAPS (4) styles
    (lightmap.c:1088 reserves space for all of them, :1152-1156 writes four)."""
    txt = mapgen.room_grid_map(1, 1, seed=11, clutter=1.0)
    extra = ""
    for k, sty in enumerate([1, 2, 3, 5, 8, 13, 31]):
        extra += ('{\n"classname" "light"\n"origin" "%d %d %d"\n"light" "%d"\n"_color" "1 %g %g"\n"style" "%d"\n}\n'
                  % (-90 + 30 * k, 60 - 20 * k, 20, 150 + 20 * k, 0.5 + 0.05 * k, 1.0 - 0.1 * k, sty))
    return txt + extra


def nolights_map_text() -> str:
    """A sealed room with no light entity and no emissive surface: everything clamps to the floor value 1."""
    txt = mapgen.room_grid_map(1, 1, seed=12)
    i = txt.index('{\n"classname" "light"')
    return txt[:i]


def angled_map_text() -> str:
    """2x2 rooms with non-axial geometry (45-degree pillars, ramps: PLANE_ANYX/ANYY/ANYZ nodes in the trace tree),
    a sky panel (SURF_SKY: patches that never collect, qrad3.c:393-397) and a warp surface (SURF_WARP: unlit face,
    lightmap.c:981)."""
    txt = mapgen.room_grid_map(2, 2, seed=21, clutter=0.3, lamps=0.5)
    head, tail = txt.split('}\n{\n"classname" "info_player_start"', 1)
    extra = ""
    # room pitch is 272; rooms are centred on (+-136, +-136) around the origin-shifted grid, z from -64 to 64
    for (cx, cy) in [(-200, -200), (120, -176), (-120, 152), (200, 200)]:
        extra += mapgen.diamond_pillar(cx, cy, 32, -64, 32, mapgen.TEX_WALL)
    extra += mapgen.ramp_brush(-232, -120, -64, -168, -56, 8, 40, mapgen.TEX_FLOOR)
    extra += mapgen.ramp_brush(40, 40, -64, 136, 104, 48, 8, mapgen.TEX_FLOOR, top_flags=8)       # SURF_WARP top
    extra += mapgen.box_brush(64, -232, 56, 128, -168, 64, mapgen.TEX_WALL, bottom_tex=mapgen.TEX_CEIL)
    extra = extra.replace("b200/ceil 0 0 0 1 1 0 0 0", "b200/ceil 0 0 0 1 1 0 4 0", 1)            # SURF_SKY panel
    return head + extra + '}\n{\n"classname" "info_player_start"' + tail


def map_text(kind: str) -> str:
    if kind == "angled":
        return angled_map_text()
    if kind == "manystyles":
        return manystyles_map_text()
    if kind == "nolights":
        return nolights_map_text()
    if kind == "sample":
        return sample_map_text()
    if kind == "grid2":
        return mapgen.room_grid_map(2, 2, seed=1, clutter=0.5, lamps=0.3)
    if kind == "grid3":
        return mapgen.room_grid_map(3, 3, seed=5, room=128, height=96, door_w=48, door_h=64, clutter=0.6, lamps=0.5)
    if kind == "hall":
        return mapgen.light_hall_map(200, seed=2, size=1024, pillars=3)
    if kind == "styles":
        return styles_map_text()
    raise KeyError(kind)


def main():
    if not R.have_ref():
        raise SystemExit("oracle/_ref is missing: run `make -C oracle ref` first (needs /root/reference)")
    tools = mapgen.find_map_tools()
    assets = mapgen.write_assets(HERE)          # tests/golden/assets/{pics,textures,maps}
    work = Path("/tmp/qrad_make_golden")
    shutil.rmtree(work, ignore_errors=True)
    compiled = {}
    for name, (kind, chop, flags) in CASES.items():
        if kind not in compiled:
            root = work / kind
            maps = mapgen.link_assets(root, assets)
            mp = maps / (kind + ".map")
            mp.write_text(map_text(kind))
            bsp = mapgen.compile_map(mp, tools, chop=chop)
            dst = assets / "maps" / (kind + ".bsp")
            shutil.copyfile(bsp, dst)
            compiled[kind] = dst
        src = compiled[kind]
        run_root = work / ("run_" + name)
        staged = R.stage_bsp(src, run_root, assets)
        dump = run_root / "dump"
        info = R.run_ref_harness(staged, flags, dumpdir=dump)
        lump = R.lighting_lump(staged)
        (HERE / (name + ".lump")).write_bytes(lump)
        if (dump / "transfers.bin").exists():
            cnt, _, _ = R.read_transfers(dump / "transfers.bin")
        else:
            cnt = np.zeros(int(info["num_patches"]), "<i4")
        cnt.astype("<i4").tofile(HERE / (name + ".numtransfers"))
        rays = R.read_rays(dump / "rays.bin")
        rays[:: max(1, len(rays) // 3000)].tofile(HERE / (name + ".rays"))
        leafs = np.fromfile(dump / "rayleafs.bin", "<i4")
        assert len(leafs) == len(rays)
        leafs[:: max(1, len(rays) // 3000)].tofile(HERE / (name + ".rayleafs"))
        meta = dict(
            bsp=kind + ".bsp", flags=flags, num_patches=int(info["num_patches"]),
            total_transfer=int(info.get("total_transfer", 0)), numdlights=int(info.get("numdlights", 0)),
            lightdatasize=len(lump), added=[float(np.float32(a)) for a in info.get("added", [])],
            rays_stride=max(1, len(rays) // 3000), rays_direct=int(info.get("rays_direct", 0)),
            rays_transfers=int(info.get("rays_transfers", 0)),
            lump_md5=hashlib.md5(lump).hexdigest(), faces_md5=hashlib.md5(R.faces_lump(staged)).hexdigest(),
            file_md5=hashlib.md5(staged.read_bytes()).hexdigest(),
            totallight_md5=hashlib.md5(R.read_vec3(dump / "totallight.bin").tobytes()).hexdigest()
            if (dump / "totallight.bin").exists() else None,
        )
        (HERE / (name + ".json")).write_text(json.dumps(meta, indent=1) + "\n")
        if name == "samplec":
            cnt, patch, tr = R.read_transfers(dump / "transfers.bin")
            np.savez_compressed(HERE / "samplec.transfers.npz", numtransfers=cnt, patch=patch, transfer=tr)
            np.save(HERE / "samplec.totallight.npy", R.read_vec3(dump / "totallight.bin"))
            fl = R.read_facelight(dump / "facelight.bin")
            np.save(HERE / "samplec.facelight0.npy",
                    np.concatenate([f["samples"][0] if f["samples"] else np.zeros((0, 3), "f4") for f in fl]))
        print(name, json.dumps(meta))
    # maps/ only keeps the compiled .bsp inputs
    for p in (assets / "maps").iterdir():
        if p.suffix != ".bsp":
            p.unlink()

    # cross-check with SURVEY.md appendix B when the reference's own palette/textures are at hand
    real = R.REF_DIR / "assets"
    if (real / "pics" / "colormap.pcx").exists():
        known = {("-bounce", "8"): "2a19db27f82f9c271ba87d264c38d7ff",
                 ("-bounce", "8", "-extra"): "8ff35aaad6b9e9e35df60b0649d809ec",
                 ("-bounce", "0"): "becdca71e8f7e7748f18b54b20c72e88"}
        for flags, md5 in known.items():
            staged = R.stage_bsp(compiled["sample"], work / "appb", real)
            R.run_ref_harness(staged, list(flags))
            got = hashlib.md5(R.lighting_lump(staged)).hexdigest()
            print("appendix B", flags, "OK" if got == md5 else f"MISMATCH {got}")
            assert got == md5


if __name__ == "__main__":
    main()
