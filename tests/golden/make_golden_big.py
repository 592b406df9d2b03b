#!/usr/bin/env python
"""Reference-derived fixtures for the LARGE BASELINE configs (C3, C4, C5), from the reference run in this container.

Needs `make -C oracle ref` (compiles the reference in place; ref_harness_big = the same translation units with the
65 000-patch limit lifted, oracle/Makefile) and bench_maps/*.bsp.gz (compiled by the reference qbsp3 + qvis3).

  c4_hall2000[_b2]   BASELINE configs[3], 2 000 lights: lit completely by the UNMODIFIED reference harness
                     (-bounce 0 -extra, and -bounce 2)                     -> lump (gz), md5s, numtransfers, added
  c3_grid13          BASELINE configs[2], 149 021 patches, -chop 32 -bounce 8: lit completely by ref_harness_big
                     (about 10 minutes of one core)                        -> lump (gz), md5s, numtransfers (gz), added
  c5_rows            BASELINE configs[4], 1 006 616 patches, -chop 8: MakeTransfers of the reference for a
                     deterministic 1/256 sample of the source rows (i = 100 + 256 k): per row numtransfers, root
                     TestLine_r calls and a 64-bit digest of all its (patch, weight) records
  c5_cols            the same map: for a window of 64 receivers, EVERY source that can see one of them is run
                     through the reference's MakeTransfers and BuildFacelights; the first bounce's ShootLight sum
                     (ascending source order, float32, qrad3.c:419-441) and CollectLight give their totallight /
                     radiosity after one bounce

usage: python tests/golden/make_golden_big.py [c4] [c3] [c5rows] [c5cols]      (default: all; c3 and c5 are slow)
"""
from __future__ import annotations

import gzip
import hashlib
import json
import struct
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))

import bench  # noqa: E402
import refdump as R  # noqa: E402

C5_ROWS_FIRST, C5_ROWS_STRIDE = 100, 256
C5_COLS_N = 64
NPROC = 8


def row_digest(patch: np.ndarray, transfer: np.ndarray) -> int:
    """64-bit digest of one transfer row: blake2b over the u32 receiver indices followed by the u16 weights."""
    h = hashlib.blake2b(digest_size=8)
    h.update(np.ascontiguousarray(patch, "<u4").tobytes())
    h.update(np.ascontiguousarray(transfer, "<u2").tobytes())
    return int.from_bytes(h.digest(), "little")


def read_rows(path):
    """-> list of (patch index, rays, receivers u32[], weights u16[]) from a -rowsout file of the harness."""
    buf = Path(path).read_bytes()
    n = struct.unpack_from("<i", buf, 0)[0]
    off, out = 4, []
    for _ in range(n):
        i, nt, rays = struct.unpack_from("<iiq", buf, off)
        off += 16
        pj = np.frombuffer(buf, "<u4", nt, off)
        off += 4 * nt
        w = np.frombuffer(buf, "<u2", nt, off)
        off += 2 * nt
        out.append((i, rays, pj, w))
    return out


def full_case(name, workload, harness, extra_flags=None):
    _, flags, _ = bench.WORKLOADS[workload]
    flags = list(extra_flags or flags)
    t0 = time.time()
    with tempfile.TemporaryDirectory() as tmp:
        bsp = bench.stage_workload(workload, Path(tmp))
        dump = Path(tmp) / "dump"
        dump.mkdir()
        cmd = [str(R.REF_DIR / harness), *flags, "-dumpdir", str(dump), str(bsp)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            raise SystemExit(r.stdout[-2000:] + r.stderr[-2000:])
        info = json.loads([l for l in r.stdout.splitlines() if l.startswith("REFTIMING ")][0][10:])
        lump = R.lighting_lump(bsp)
        with gzip.GzipFile(HERE / (name + ".lump.gz"), "wb", compresslevel=9, mtime=0) as g:
            g.write(lump)
        meta = dict(workload=workload, bsp=bench.WORKLOADS[workload][0], flags=flags, harness=harness,
                    num_patches=int(info["num_patches"]), total_transfer=int(info.get("total_transfer", 0)),
                    lightdatasize=len(lump), added=[float(np.float32(a)) for a in info.get("added", [])],
                    rays_direct=int(info["rays_direct"]), rays_transfers=int(info["rays_transfers"]),
                    lump_md5=hashlib.md5(lump).hexdigest(), faces_md5=hashlib.md5(R.faces_lump(bsp)).hexdigest(),
                    file_md5=hashlib.md5(bsp.read_bytes()).hexdigest(),
                    ref_seconds={k: info[k] for k in ("setup_s", "facelights_s", "transfers_s", "bounce_s", "final_s")})
        if (dump / "transfers.bin").exists():
            buf = (dump / "transfers.bin").read_bytes()
            n = struct.unpack_from("<i", buf, 0)[0]
            cnt = np.frombuffer(buf, "<i4", n, 4)
            with gzip.GzipFile(HERE / (name + ".numtransfers.gz"), "wb", compresslevel=9, mtime=0) as g:
                g.write(cnt.tobytes())
            meta["numtransfers_sum"] = int(cnt.astype(np.int64).sum())
        if (dump / "totallight.bin").exists():
            meta["totallight_md5"] = hashlib.md5(R.read_vec3(dump / "totallight.bin").tobytes()).hexdigest()
        (HERE / (name + ".json")).write_text(json.dumps(meta, indent=1) + "\n")
        print(name, json.dumps(meta), f"{time.time() - t0:.0f} s", flush=True)


def c5_stage(tmp):
    return bench.stage_workload("c5_grid16_chop8", Path(tmp))


def c5_rows():
    _, flags, _ = bench.WORKLOADS["c5_grid16_chop8"]
    t0 = time.time()
    with tempfile.TemporaryDirectory() as tmp:
        bsp = c5_stage(tmp)

        def part(k):
            out = Path(tmp) / f"rows{k}.bin"
            cmd = [str(R.REF_DIR / "ref_harness_big"), *flags, "-rows",
                   f"{C5_ROWS_FIRST + k * C5_ROWS_STRIDE}:{C5_ROWS_STRIDE * NPROC}", "-rowsout", str(out), str(bsp)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode:
                raise SystemExit(r.stdout[-2000:] + r.stderr[-2000:])
            return read_rows(out)

        with ThreadPoolExecutor(NPROC) as ex:
            rows = [x for p in ex.map(part, range(NPROC)) for x in p]
    rows.sort(key=lambda x: x[0])
    idx = np.array([r[0] for r in rows], "<i4")
    assert np.array_equal(idx, np.arange(C5_ROWS_FIRST, idx[-1] + 1, C5_ROWS_STRIDE))
    np.savez_compressed(HERE / "c5_rows.npz", patch=idx, numtransfers=np.array([len(r[2]) for r in rows], "<i4"),
                        rays=np.array([r[1] for r in rows], "<i8"),
                        digest=np.array([row_digest(r[2], r[3]) for r in rows], "<u8"),
                        first=np.int32(C5_ROWS_FIRST), stride=np.int32(C5_ROWS_STRIDE),
                        # the first sampled row in full, as a readable example / debugging aid
                        row0_patch=rows[0][2], row0_transfer=rows[0][3])
    print(f"c5_rows: {len(rows)} rows, {sum(len(r[2]) for r in rows)} transfers, {sum(r[1] for r in rows)} rays, "
          f"{time.time() - t0:.0f} s", flush=True)


def c5_cols(first_receiver=None):
    _, flags, _ = bench.WORKLOADS["c5_grid16_chop8"]
    t0 = time.time()
    with tempfile.TemporaryDirectory() as tmp:
        bsp = c5_stage(tmp)
        a = first_receiver if first_receiver is not None else pick_window(bsp, flags)

        def part(k):
            out = Path(tmp) / f"cols{k}.bin"
            cmd = [str(R.REF_DIR / "ref_harness_big"), *flags, "-cols", f"{a}:{C5_COLS_N}", "-part", f"{k}:{NPROC}",
                   "-colsout", str(out), str(bsp)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode:
                raise SystemExit(r.stdout[-2000:] + r.stderr[-2000:])
            return out.read_bytes()

        with ThreadPoolExecutor(NPROC) as ex:
            bufs = list(ex.map(part, range(NPROC)))
    recv = np.dtype([("area", "<f4"), ("sky", "<i4"), ("totallight", "<f4", 3), ("reflectivity", "<f4", 3)])
    srcs = []
    for buf in bufs:
        a2, n = struct.unpack_from("<ii", buf, 0)
        assert (a2, n) == (a, C5_COLS_N)
        win = np.frombuffer(buf, recv, n, 8)
        off = 8 + n * recv.itemsize
        while off < len(buf):
            i = struct.unpack_from("<i", buf, off)[0]
            rad = np.frombuffer(buf, "<f4", 3, off + 4)
            cnt = struct.unpack_from("<i", buf, off + 16)[0]
            pairs = np.frombuffer(buf, "<i4", 2 * cnt, off + 20).reshape(cnt, 2)
            off += 20 + 8 * cnt
            srcs.append((i, rad, pairs))
    srcs.sort(key=lambda x: x[0])
    # ShootLight of the first bounce for the window, sources ascending, float32 mul then add (qrad3.c:430-439)
    ill = np.zeros((C5_COLS_N, 3), np.float32)
    nsrc = 0
    for i, rad, pairs in srcs:
        if not len(pairs):
            continue
        nsrc += 1
        send = (rad / np.float32(65536)).astype(np.float32)
        for j, w in pairs:
            ill[j - a] = ill[j - a] + send * np.float32(w)
    total = win["totallight"].copy()
    radio = np.zeros_like(ill)
    for k in range(C5_COLS_N):   # CollectLight, qrad3.c:383-409
        if win["sky"][k]:
            continue
        total[k] = total[k] + ill[k] / win["area"][k]
        radio[k] = ill[k] * win["reflectivity"][k]
    np.savez_compressed(HERE / "c5_cols.npz", first=np.int32(a), totallight=total, radiosity=radio,
                        sources_examined=np.int32(len(srcs)), sources_hitting=np.int32(nsrc))
    print(f"c5_cols: receivers {a}..{a + C5_COLS_N - 1}, {len(srcs)} sources examined, {nsrc} with a transfer into the "
          f"window, totallight mean {total.mean():.4f}, {time.time() - t0:.0f} s", flush=True)


def pick_window(bsp, flags):
    """A window of 64 consecutive patches of ONE face that is lit and sees as few clusters as possible (fewest
    sources to run): chosen from the GPU-free host half of the product (patch clusters only; no result of the
    product is used as an expectation)."""
    import golden_util as G
    from qengine_b200 import qrad as Q
    p = G.params_from_flags(flags)
    sc = Q.Scene(bsp).prepare(p)
    pt = sc.patches()
    cl = pt["cluster"]
    # PVS sizes per cluster via the harness-independent decompress in numpy
    d = Path(bsp).read_bytes()
    ofs, ln = struct.unpack_from("<ii", d, 8 + 3 * 8)
    vis = d[ofs:ofs + ln]
    nc = struct.unpack_from("<i", vis, 0)[0]
    counts = np.bincount(cl[cl >= 0], minlength=nc)
    seen_by = np.zeros(nc, np.int64)     # patches whose cluster sees cluster d
    for c in range(nc):
        o = struct.unpack_from("<i", vis, 4 + 8 * c)[0]
        row = bytearray()
        nb = (nc + 7) >> 3
        while len(row) < nb:
            if vis[o]:
                row.append(vis[o]); o += 1
            else:
                row.extend(b"\0" * vis[o + 1]); o += 2
        bits = np.unpackbits(np.frombuffer(bytes(row[:nb]), np.uint8), bitorder="little")[:nc]
        seen_by += bits.astype(np.int64) * counts[c]
    best = None
    n = len(cl)
    for a in range(0, n - C5_COLS_N, C5_COLS_N):
        w = cl[a:a + C5_COLS_N]
        if w.min() < 0 or w.min() != w.max() or pt["sky"][a:a + C5_COLS_N].any():
            continue
        if len(set(pt["face"][a:a + C5_COLS_N])) != 1:
            continue
        s = seen_by[w[0]]
        # prefer small source sets, but not degenerate ones
        if s > 20000 and (best is None or s < best[0]):
            best = (s, a)
    print("window", best, flush=True)
    return best[1]


def main():
    if not (R.REF_DIR / "ref_harness_big").exists():
        raise SystemExit("oracle/_ref/ref_harness_big missing: run `make -C oracle ref`")
    what = sys.argv[1:] or ["c4", "c3", "c5rows", "c5cols"]
    if "c4" in what:
        full_case("c4_hall2000", "c4_hall2000", "ref_harness")
        full_case("c4_hall2000_b2", "c4_hall2000", "ref_harness", ["-bounce", "2"])
    if "c5rows" in what:
        c5_rows()
    if "c5cols" in what:
        c5_cols()
    if "c3" in what:
        full_case("c3_grid13", "c3_grid13", "ref_harness_big")


if __name__ == "__main__":
    main()
