#!/usr/bin/env python
"""Times the bounce launches over the first 1/n of the slices of a bench workload, on ONE GPU: what an n-GPU shard's
launch costs without needing n GPUs (results of these runs are not lighting).  For every shard size: the warp-per-slice
kernels alone (QRAD_BOUNCE_COOP huge) and with the block-cooperative kernel taking the longest lists (default rule).
  python tools/bounce_shard_probe.py [workload]"""
import json
import os
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from qengine_b200 import qrad as Q  # noqa: E402
from qengine_b200 import workloads as W  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else W.DEFAULT_WORKLOAD
p = Q.params_from_flags(W.WORKLOADS[name][1])
out = []
with tempfile.TemporaryDirectory() as tmp:
    bsp = W.stage_workload(name, Path(tmp))
    sc = Q.Scene(bsp).prepare(p)
    dev = Q.Device(0)
    dev.upload(sc)
    dev.build_facelights(p.extrasamples, p.numbounce)
    fracs = [float(x) for x in os.environ.get("PROBE_SHARDS", "1,0.5,0.25,0.125").split(",")]
    for frac in fracs:
        variants = ["warp", "warp+coop"] + ["warp+coop:" + f for f in os.environ.get("PROBE_FACTORS", "").split(",") if f]
        for variant in variants:
            os.environ["QRAD_DEBUG_BOUNCE_SHARD"] = str(frac)
            os.environ.pop("QRAD_BOUNCE_COOP_FACTOR", None)
            if variant == "warp":
                os.environ["QRAD_BOUNCE_COOP"] = "1000000000"
            else:
                os.environ.pop("QRAD_BOUNCE_COOP", None)
                if ":" in variant:
                    os.environ["QRAD_BOUNCE_COOP_FACTOR"] = variant.split(":")[1]
            dev.make_transfers(p.nopvs)      # the work lists are chosen here
            full = dev.stats()["transfer_bytes"]
            dev.bounce(4, want_added=False)
            dev.bounce(8, want_added=False)
            ms = dev.stats()["ms_bounce_kernel"]
            out.append({"shard": frac, "kernels": variant, "ms_per_launch": ms,
                        "stored_gbs": full * frac / (ms * 1e-3) / 1e9 if ms > 0 else None})
            print(out[-1], file=sys.stderr, flush=True)
print(json.dumps({"workload": name, "stored_bytes_full": full, "runs": out}, indent=1))
