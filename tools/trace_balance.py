#!/usr/bin/env python
"""How evenly does pass 1 (k_visibility) of a workload divide over `world` ranks?  Runs every rank's share alone on ONE
GPU (qrad_cuda_trace_shard) for several settings of QRAD_OWNER_CHUNKS (contiguous chunks of row groups per rank).
  python tools/trace_balance.py [workload] [world] [chunks ...]"""
import ctypes as C
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
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
settings = [int(x) for x in sys.argv[3:]] or [8, 32, 128]
with tempfile.TemporaryDirectory() as tmp:
    bsp = W.stage_workload(name, Path(tmp))
    p = Q.params_from_flags(W.WORKLOADS[name][1])
    sc = Q.Scene(bsp).prepare(p)
    dev = Q.Device(0).upload(sc)
    L = Q.lib()
    for chunks in settings:
        os.environ["QRAD_OWNER_CHUNKS"] = str(chunks)
        ms, rays = [], []
        for r in range(world):
            t, n = C.c_float(), C.c_int64()
            for rep in range(2):
                Q._dck(L.qrad_cuda_trace_shard(dev._h, C.c_int(0), C.c_int(r), C.c_int(world), C.byref(t), C.byref(n)))
            ms.append(t.value)
            rays.append(n.value)
        out = {"workload": name, "world": world, "owner_chunks": chunks, "trace_ms": [round(x, 1) for x in ms],
               "max_ms": max(ms), "mean_ms": sum(ms) / world, "imbalance": max(ms) / (sum(ms) / world) - 1,
               "rays_max_over_mean": max(rays) / (sum(rays) / world) - 1}
        print(json.dumps(out), flush=True)
