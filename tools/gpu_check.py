#!/usr/bin/env python
"""Phase-by-phase parity check of libqrad_cuda against the reference (oracle/_ref/ref_harness).

Development aid (the pytest suite under tests/ is the judged gate).  For each case it stages a
.bsp, runs the reference with dumps, runs the CUDA path seam by seam and prints what differs.
usage: python tools/gpu_check.py [case ...]     cases: sample sample_extra grid2 grid4 hall grid10
"""
from __future__ import annotations

import hashlib
import json
import shutil
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))

import refdump as R  # noqa: E402
from qengine_b200 import mapgen  # noqa: E402
from qengine_b200 import qrad as Q  # noqa: E402

WORK = Path("/tmp/qrad_gpu_check")


def make_case(name):
    tools = mapgen.find_map_tools()
    root = WORK / name
    shutil.rmtree(root, ignore_errors=True)
    flags = ["-bounce", "8"]
    if name.startswith("sample"):
        maps = mapgen.link_assets(root, R.REF_DIR / "assets")
        src = (R.REF_DIR / "assets" / "maps" / "sample.map").read_text()
        src = src.replace('"classname" "light"\n', '"classname" "light"\n"_color" "1 1 1"\n')
        (maps / "samplec.map").write_text(src)
        bsp = mapgen.compile_map(maps / "samplec.map", tools)
        if name == "sample_extra":
            flags += ["-extra"]
        if name == "sample_chop16":
            flags += ["-chop", "16", "-extra"]
    elif name.startswith("grid"):
        n = int(name[4:])
        bsp = mapgen.build_case(root, name, mapgen.room_grid_map(n, n, seed=1, clutter=0.5, lamps=0.3), tools)
        flags += ["-chop", "32"]
    elif name == "hall":
        bsp = mapgen.build_case(root, name, mapgen.light_hall_map(200, seed=2, size=1024, pillars=3), tools)
        flags = ["-bounce", "2", "-extra"]
    else:
        raise SystemExit("unknown case " + name)
    return bsp, flags


def params_from_flags(flags):
    kw = {}
    it = iter(flags)
    for f in it:
        if f == "-chop":
            kw["subdiv"] = float(int(next(it)))
        elif f == "-bounce":
            kw["numbounce"] = int(next(it))
        elif f == "-extra":
            kw["extrasamples"] = 1
    return Q.default_params(**kw)


def cmp(name, a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape:
        print(f"   {name}: SHAPE {a.shape} vs {b.shape}")
        return False
    ok = np.array_equal(a, b)
    if ok:
        print(f"   {name}: identical ({a.size} values)")
    else:
        bad = np.argwhere(a != b)
        d = np.abs(a.astype(np.float64) - b.astype(np.float64))
        print(f"   {name}: {len(bad)} of {a.size} differ, max |d| = {d.max():.6g}, first at {bad[0].tolist()}: {a[tuple(bad[0])]} vs {b[tuple(bad[0])]}")
    return ok


def run_case(name):
    print(f"=== {name}")
    bsp, flags = make_case(name)
    unlit = bsp.with_suffix(".unlit")
    shutil.copyfile(bsp, unlit)
    dump = WORK / name / "dump"
    t0 = time.time()
    ref = R.run_ref_harness(bsp, flags, dumpdir=dump)
    print(f"   reference: {time.time() - t0:.2f} s  {json.dumps({k: v for k, v in ref.items() if k != 'added'})}")
    ref_lump = R.lighting_lump(bsp)
    ref_faces = R.faces_lump(bsp)
    ref_file = bsp.read_bytes()
    shutil.copyfile(unlit, bsp)

    p = params_from_flags(flags)
    sc = Q.Scene(bsp).prepare(p)
    dev = Q.Device(0)
    dev.set_count_nodes(True)
    dev.upload(sc)
    ok = True
    # rays
    rays = R.read_rays(dump / "rays.bin")
    got = dev.test_lines(rays["start"], rays["stop"])
    ok &= cmp("TestLine", got.astype(np.int32), (rays["result"] != 0).astype(np.int32))
    # facelights
    dev.build_facelights(p.extrasamples, p.numbounce)
    first, nst, sty, origins, slots = dev.facelight()
    rfl = R.read_facelight(dump / "facelight.bin")
    r_orig = np.concatenate([f["origins"] for f in rfl]) if rfl else np.zeros((0, 3))
    ok &= cmp("facelight.origins", origins, r_orig)
    r_ns = np.array([len(f["stylenums"]) for f in rfl])
    ok &= cmp("facelight.numstyles", nst, r_ns)
    r_s0 = np.concatenate([f["samples"][0] if f["samples"] else np.zeros((f["numsamples"], 3), "f4") for f in rfl])
    ok &= cmp("facelight.samples[style0]", slots[0], r_s0)
    rp, _ = R.read_patches(dump / "patches.bin")
    pl = dev.patch_light()
    ok &= cmp("patch.samplelight", pl["samplelight"], rp["samplelight"])
    ok &= cmp("patch.samples", pl["samples"], rp["samples"])
    if p.numbounce > 0:
        nnz = dev.make_transfers(0)
        cnt, rpatch, rtr = R.read_transfers(dump / "transfers.bin")
        ok &= cmp("numtransfers", dev.numtransfers(), cnt)
        row_ptr, gp, gt = dev.transfers()
        ok &= cmp("transfer.patch", gp, rpatch.astype(np.uint32))
        ok &= cmp("transfer.transfer", gt, rtr)
        added = dev.bounce(p.numbounce)
        ok &= cmp("added", added, np.array(ref["added"], np.float32))
        pl = dev.patch_light()
        ok &= cmp("totallight", pl["totallight"], R.read_vec3(dump / "totallight.bin"))
    data, lightofs, styles = dev.final_light(p.ambient, p.maxlight, p.lightscale, p.subdiv, p.numbounce)
    a = np.frombuffer(ref_lump, np.uint8)
    if a.size == data.size:
        d = np.abs(a.astype(int) - data.astype(int))
        print(f"   lump: {a.size} bytes, {int((d > 0).sum())} differ, {int((d > 1).sum())} by more than 1, max {int(d.max()) if d.size else 0}; "
              f"md5 ref {hashlib.md5(ref_lump).hexdigest()} ours {hashlib.md5(data.tobytes()).hexdigest()}")
        ok &= bool((d == 0).all())
    else:
        print(f"   lump: SIZE {a.size} vs {data.size}")
        ok = False
    sc.store_lighting(data, lightofs, styles)
    sc.write(bsp)
    same_file = bsp.read_bytes() == ref_file
    print(f"   faces lump identical: {R.faces_lump(bsp) == ref_faces}; whole .bsp identical: {same_file}")
    ok &= same_file
    st = dev.stats()
    print("   stats:", json.dumps(st))
    print(f"   gpu ms: facelights {st['ms_facelights']:.2f} transfers {st['ms_transfers']:.2f} (trace {st['ms_transfers_trace']:.2f}) "
          f"bounce {st['ms_bounce']:.2f} final {st['ms_final']:.2f}  | ref s: {ref.get('facelights_s')} {ref.get('transfers_s')} {ref.get('bounce_s')} {ref.get('final_s')}")
    print(f"=== {name}: {'PASS' if ok else 'FAIL'}")
    return ok


if __name__ == "__main__":
    cases = sys.argv[1:] or ["sample", "sample_extra", "grid2", "hall"]
    res = {}
    for cname in cases:
        try:
            res[cname] = run_case(cname)
        except Exception as e:  # keep going: one gpurun call should tell as much as possible
            import traceback
            traceback.print_exc()
            res[cname] = False
    print("SUMMARY", json.dumps(res))
    sys.exit(0 if all(res.values()) else 1)
