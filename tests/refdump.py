"""Readers for the flat binary dumps written by oracle/ref_harness.c (test infrastructure)."""
from __future__ import annotations

import json
import os
import shutil
import struct
import subprocess
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
REF_DIR = REPO / "oracle" / "_ref"

PATCH_DTYPE = np.dtype([
    ("origin", "<f4", 3), ("normal", "<f4", 3), ("dist", "<f4"), ("area", "<f4"),
    ("cluster", "<i4"), ("sky", "<i4"),
    ("reflectivity", "<f4", 3), ("baselight", "<f4", 3), ("totallight", "<f4", 3), ("samplelight", "<f4", 3),
    ("samples", "<i4"), ("face", "<i4"), ("next", "<i4"),
    ("mins", "<f4", 3), ("maxs", "<f4", 3), ("numtransfers", "<i4"), ("numpoints", "<i4"),
])
LIGHT_DTYPE = np.dtype([
    ("cluster", "<i4"), ("type", "<i4"), ("intensity", "<f4"), ("style", "<i4"),
    ("origin", "<f4", 3), ("color", "<f4", 3), ("normal", "<f4", 3), ("stopdot", "<f4"),
])
RAY_DTYPE = np.dtype([("start", "<f4", 3), ("stop", "<f4", 3), ("result", "<i4")])


def have_ref() -> bool:
    return (REF_DIR / "ref_harness").exists()


def read_patches(path):
    buf = Path(path).read_bytes()
    n = struct.unpack_from("<i", buf, 0)[0]
    rec = np.frombuffer(buf, dtype=PATCH_DTYPE, count=n, offset=4)
    off = 4 + n * PATCH_DTYPE.itemsize
    nf = struct.unpack_from("<i", buf, off)[0]
    heads = np.frombuffer(buf, dtype="<i4", count=nf, offset=off + 4)
    return rec, heads


def read_windings(path):
    buf = Path(path).read_bytes()
    n = struct.unpack_from("<i", buf, 0)[0]
    off = 4
    out = []
    for _ in range(n):
        k = struct.unpack_from("<i", buf, off)[0]
        off += 4
        out.append(np.frombuffer(buf, dtype="<f4", count=3 * k, offset=off).reshape(k, 3))
        off += 12 * k
    return out


def read_lights(path):
    buf = Path(path).read_bytes()
    n = struct.unpack_from("<i", buf, 0)[0]
    return np.frombuffer(buf, dtype=LIGHT_DTYPE, count=n, offset=4)


def read_rays(path):
    buf = Path(path).read_bytes()
    n = struct.unpack_from("<i", buf, 0)[0]
    return np.frombuffer(buf, dtype=RAY_DTYPE, count=n, offset=4)


def read_vec3(path):
    buf = Path(path).read_bytes()
    n = struct.unpack_from("<i", buf, 0)[0]
    return np.frombuffer(buf, dtype="<f4", count=3 * n, offset=4).reshape(n, 3)


def read_reflectivity(path):
    return read_vec3(path)


def read_transfers(path):
    """-> (numtransfers[N], patch[nnz] u16, transfer[nnz] u16) in the reference's row order."""
    buf = Path(path).read_bytes()
    n = struct.unpack_from("<i", buf, 0)[0]
    cnt = np.frombuffer(buf, dtype="<i4", count=n, offset=4)
    rec = np.frombuffer(buf, dtype="<u2", offset=4 + 4 * n).reshape(-1, 2)
    assert rec.shape[0] == int(cnt.sum())
    return cnt, rec[:, 0], rec[:, 1]


def read_facelight(path):
    """-> list per face of dict(numsamples, stylenums, origins[n,3], samples[numstyles][n,3])."""
    buf = Path(path).read_bytes()
    nf = struct.unpack_from("<i", buf, 0)[0]
    off = 4
    faces = []
    for _ in range(nf):
        ns, nst = struct.unpack_from("<ii", buf, off)
        off += 8
        styles = list(struct.unpack_from("<%di" % nst, buf, off)) if nst else []
        off += 4 * nst
        origins = np.frombuffer(buf, dtype="<f4", count=3 * ns, offset=off).reshape(ns, 3) if ns else np.zeros((0, 3), "f4")
        off += 12 * ns
        samples = []
        for _s in range(nst):
            samples.append(np.frombuffer(buf, dtype="<f4", count=3 * ns, offset=off).reshape(ns, 3))
            off += 12 * ns
        faces.append(dict(numsamples=ns, stylenums=styles, origins=origins, samples=samples))
    return faces


def lighting_lump(bsp_path) -> bytes:
    d = Path(bsp_path).read_bytes()
    ofs, ln = struct.unpack_from("<ii", d, 8 + 7 * 8)
    return d[ofs:ofs + ln]


def faces_lump(bsp_path) -> bytes:
    d = Path(bsp_path).read_bytes()
    ofs, ln = struct.unpack_from("<ii", d, 8 + 6 * 8)
    return d[ofs:ofs + ln]


def run_ref_harness(bsp_path, flags, dumpdir=None, stopafter=None, timeout=3600):
    """Run the reference (oracle/_ref/ref_harness) IN PLACE on bsp_path (it overwrites it). Returns REFTIMING dict."""
    cmd = [str(REF_DIR / "ref_harness")] + list(flags)
    if dumpdir is not None:
        os.makedirs(dumpdir, exist_ok=True)
        cmd += ["-dumpdir", str(dumpdir)]
    if stopafter is not None:
        cmd += ["-stopafter", str(stopafter)]
    cmd.append(str(bsp_path))
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError("ref_harness failed: " + r.stdout[-1500:] + r.stderr[-500:])
    for line in r.stdout.splitlines():
        if line.startswith("REFTIMING "):
            return json.loads(line[len("REFTIMING "):])
    raise RuntimeError("no REFTIMING line: " + r.stdout[-500:])


def stage_bsp(src_bsp, dst_root, assets_src):
    """Copy src_bsp to dst_root/assets/maps/ with pics/ + textures/ from assets_src next to it."""
    from qengine_b200 import mapgen
    maps = mapgen.link_assets(Path(dst_root), Path(assets_src))
    dst = maps / Path(src_bsp).name
    shutil.copyfile(src_bsp, dst)
    return dst
