#!/usr/bin/env python
"""One make_transfers + one bounce + final_light on a workload: the target of the ncu captures kept under profiles/.
usage: python tools/profile_transfers.py [workload]   (default c5_chop16, the C5 map at -chop 16: its kernels fit a replay)"""
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from qengine_b200 import qrad as Q  # noqa: E402
from qengine_b200 import workloads as W  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c5_chop16"
with tempfile.TemporaryDirectory() as tmp:
    bsp = W.stage_workload(name, Path(tmp))
    p = Q.params_from_flags(W.WORKLOADS[name][1])
    sc = Q.Scene(bsp).prepare(p)
    dev = Q.Device(0).upload(sc)
    dev.build_facelights(p.extrasamples, p.numbounce)
    dev.make_transfers(p.nopvs)
    dev.bounce(1, want_added=False)
    dev.final_light(p.ambient, p.maxlight, p.lightscale, p.subdiv, p.numbounce)
    st = dev.stats()
    print({k: st[k] for k in ("num_patches", "total_transfers", "rays_transfers", "ms_transfers_trace", "ms_tr_rows", "ms_tr_columns")})
