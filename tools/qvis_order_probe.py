#!/usr/bin/env python
"""How much parallelism does the reference order of qvis3 PortalFlow leave?  Reads a .prt file (reference qbsp3 output),
runs the numpy restatement of BasePortalVis + SimpleFlood (oracle/qvis_stage1.py), sorts the portals by nummightsee as
SortPortals does and levels the dependency graph "portal k consults the finished portals j < k of its flood set"
(flow.c:391-396).  Evidence for DESIGN.md section 9; a development tool, not part of the product.
  usage: python tools/qvis_order_probe.py map.prt"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "oracle"))
import qvis_stage1 as V  # noqa: E402

t0 = time.time()
pt = V.Portals(sys.argv[1])
front = V.base_portal_vis(pt)
print(f"front {time.time() - t0:.1f} s, P {pt.P} nc {pt.nc} density {front.mean():.3f}")
flood = V.simple_flood(pt, front)
might = flood.sum(1)
print(f"flood done {time.time() - t0:.1f} s; mean mightsee {might.mean():.1f}")
order, rank = V.sorted_rank(might)
level = V.dependency_levels(flood, order, rank)
print("levels", level.max(), "of", pt.P, "portals; widest level", np.bincount(level).max(), "; mean width", round(pt.P / level.max(), 1))
