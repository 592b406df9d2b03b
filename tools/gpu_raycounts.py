#!/usr/bin/env python
"""Counts the root TestLine_r calls of a bench workload on the GPU (bit-exact parity makes the count equal to
the reference's; tests compare it with the C oracle's own counter on the small cases).  Writes JSON to stdout."""
import json
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402
import golden_util as G  # noqa: E402
from qengine_b200 import qrad as Q  # noqa: E402

out = {}
for name in sys.argv[1:]:
    _, flags, _ = bench.WORKLOADS[name]
    p = G.params_from_flags(flags)
    with tempfile.TemporaryDirectory() as tmp:
        bsp = bench.stage_workload(name, Path(tmp))
        sc = Q.Scene(bsp).prepare(p)
        dev = Q.Device(0)
        dev.set_count_nodes(True)
        Q.rad_world(sc, dev, p)
        st = dev.stats()
        out[name] = {k: st[k] for k in ("rays_transfers", "rays_direct", "node_visits", "candidate_pairs",
                                        "total_transfers", "num_patches", "num_luxels", "padded_entries", "full_walks", "box_pruned_rays")}
        print(name, json.dumps(st), file=sys.stderr)
print(json.dumps(out, indent=1))
