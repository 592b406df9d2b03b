"""One small end-to-end invocation of the hot path on cuda:0, checked against the oracle's golden output.

Used by __graft_entry__.smoke().  The golden lump was produced by the reference qrad3 itself
(tests/golden/make_golden.py); when the C restatement (oracle/liboracle.so) is built it is run on
the same input as a second checker.
"""
from __future__ import annotations

import hashlib
import json
import shutil
import tempfile
from pathlib import Path

import numpy as np

import golden_util as G
from qengine_b200 import qrad as Q


def run(case: str = "samplec") -> None:
    meta, flags = G.load_case(case)
    with tempfile.TemporaryDirectory() as tmp:
        bsp = G.stage_case(case, Path(tmp))
        p = G.params_from_flags(flags)
        sc = Q.Scene(bsp).prepare(p)
        dev = Q.Device(0)
        Q.rad_world(sc, dev, p)
        lump = sc.lighting()
        want = (G.GOLDEN / (case + ".lump")).read_bytes()
        if lump != want:
            a = np.frombuffer(lump, np.uint8).astype(int)
            b = np.frombuffer(want, np.uint8).astype(int)
            raise AssertionError(f"smoke: lighting lump differs from the golden ({(a != b).sum()} of {a.size} bytes)")
        # second checker: the plain-C oracle on the same input
        import sys
        sys.path.insert(0, str(G.ROOT / "oracle"))
        import oracle as O
        o = O.Oracle(bsp, numbounce=p.numbounce, extrasamples=p.extrasamples, subdiv=p.subdiv)
        olump, _ = o.rad_world()
        assert olump.tobytes() == lump, "smoke: CUDA lump differs from the C oracle"
        nt = dev.numtransfers()
        assert int(nt.sum()) == meta["total_transfer"], (int(nt.sum()), meta["total_transfer"])
        st = dev.stats()
        assert st["kernel_launches"] > 0
        print(f"smoke ok: {case} {len(lump)} lump bytes identical to the reference, "
              f"{st['rays_transfers'] + st['rays_direct']} rays, {st['kernel_launches']} kernel launches")
