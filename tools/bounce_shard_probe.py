#!/usr/bin/env python
"""Times one k_bounce launch over the first 1/n of the slices of a bench workload, on ONE GPU, for both variants of
the kernel: what an n-GPU shard's launch costs without needing n GPUs (results of these runs are not lighting).
  python tools/bounce_shard_probe.py [workload]"""
import json
import os
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402
import golden_util as G  # noqa: E402
from qengine_b200 import qrad as Q  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
_, flags, _ = bench.WORKLOADS[name]
p = G.params_from_flags(flags)
out = []
with tempfile.TemporaryDirectory() as tmp:
    bsp = bench.stage_workload(name, Path(tmp))
    sc = Q.Scene(bsp).prepare(p)
    dev = Q.Device(0)
    dev.upload(sc)
    dev.build_facelights(p.extrasamples, p.numbounce)
    os.environ["QRAD_DEBUG_SLICES"] = "1"
    dev.make_transfers(p.nopvs)
    full = dev.stats()["transfer_bytes"]
    fracs = [float(x) for x in os.environ.get("PROBE_SHARDS", "1,0.5,0.25,0.125,0.01").split(",")]
    kernels = os.environ.get("PROBE_KERNELS", "regs,ring").split(",")
    for frac in fracs:
        for kernel in kernels:
            os.environ["QRAD_DEBUG_BOUNCE_SHARD"] = str(frac)
            os.environ["QRAD_BOUNCE_KERNEL"] = kernel
            dev.bounce(4, want_added=False)
            dev.bounce(8, want_added=False)
            ms = dev.stats()["ms_bounce_kernel"]
            out.append({"shard": frac, "kernel": kernel, "ms_per_launch": ms})
            print(out[-1], file=sys.stderr)
print(json.dumps({"workload": name, "stored_bytes_full": full, "runs": out}, indent=1))
