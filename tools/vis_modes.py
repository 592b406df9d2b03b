#!/usr/bin/env python
"""Times pass 1 (k_visibility) of a workload under different proof modes (QRAD_VIS_MODE bits: 1 slice x word shaft proofs,
2 source x word shaft proofs, 4 per-ray shaft proofs, 8 interval walks, 16 guess walks) and checks that the accepted
transfers do not depend on the mode.  usage: python tools/vis_modes.py [workload] [mode ...]"""
import json
import os
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from qengine_b200 import qrad as Q  # noqa: E402
from qengine_b200 import workloads as W  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c5_grid16_chop8"
modes = [int(x) for x in sys.argv[2:]] or [23, 7, 31, 24, 16, 22, 21]
with tempfile.TemporaryDirectory() as tmp:
    bsp = W.stage_workload(name, Path(tmp))
    p = Q.params_from_flags(W.WORKLOADS[name][1])
    sc = Q.Scene(bsp).prepare(p)
    dev = Q.Device(0).upload(sc)
    ref = None
    out = []
    for m in modes:
        os.environ["QRAD_VIS_MODE"] = str(m)
        for rep in range(2):
            dev.make_transfers(p.nopvs)
        st = dev.stats()
        nt = dev.numtransfers()
        if ref is None:
            ref = nt
        same = bool((nt == ref).all())
        out.append({"mode": m, "trace_ms": st["ms_transfers_trace"], "nnz": st["total_transfers"], "same_as_first": same})
        print(json.dumps(out[-1]), flush=True)
