#!/usr/bin/env python
"""How much parallelism does the reference order of qvis3 PortalFlow leave?  Reads a .prt file (reference qbsp3 output),
restates BasePortalVis + SimpleFlood (flow.c:576-652) in numpy, sorts the portals by nummightsee as SortPortals does
(qvis3.c:118-129) and levels the dependency graph "portal k consults the finished portals j < k of its flood set"
(flow.c:391-396).  Evidence for DESIGN.md section 9; not part of the product.
  usage: python tools/qvis_order_probe.py map.prt"""
import sys
import time
from pathlib import Path

import numpy as np


def load_prt(path):
    lines = Path(path).read_text().split("\n")
    assert lines[0].strip() == "PRT1"
    nc, npz = int(lines[1]), int(lines[2])
    wind = []; leafs = []
    for ln in lines[3:3 + npz]:
        head, rest = ln.split("(", 1)
        n, l0, l1 = (int(x) for x in head.split())
        pts = np.array([[float(v) for v in seg.split(")")[0].split()] for seg in ("(" + rest).split("(")[1:]], np.float32)
        assert len(pts) == n
        wind.append(pts); leafs.append((l0, l1))
    return nc, wind, leafs
def plane_from_winding(w):
    v1 = w[2] - w[1]; v2 = w[0] - w[1]
    n = np.cross(v2, v1).astype(np.float32)
    ln = np.float32(np.sqrt(np.float64(n[0]*n[0] + n[1]*n[1] + n[2]*n[2])))
    n = (n * np.float32(1.0 / ln)).astype(np.float32)
    return n, np.float32(np.dot(w[0], n))
path = sys.argv[1]
nc, wind, leafs = load_prt(path)
P = 2 * len(wind)
normal = np.zeros((P, 3), np.float32); dist = np.zeros(P, np.float32); leaf = np.zeros(P, np.int32); W = []
leaf_portals = [[] for _ in range(nc)]
for i, (w, (l0, l1)) in enumerate(zip(wind, leafs)):
    n, d = plane_from_winding(w)
    normal[2*i] = -n; dist[2*i] = -d; leaf[2*i] = l1; W.append(w); leaf_portals[l0].append(2*i)
    normal[2*i+1] = n; dist[2*i+1] = d; leaf[2*i+1] = l0; W.append(w[::-1]); leaf_portals[l1].append(2*i+1)
maxp = max(len(w) for w in W)
pts = np.zeros((P, maxp, 3), np.float32); cnt = np.array([len(w) for w in W])
for i, w in enumerate(W): pts[i, :len(w)] = w; pts[i, len(w):] = w[0]
t0 = time.time()
front = np.zeros((P, P), bool)
for a in range(0, P, 256):
    b = min(P, a + 256)
    # d[p, j, k] = pts[j, k] . normal[p] - dist[p]
    d1 = np.einsum('jkc,pc->pjk', pts, normal[a:b]) - dist[a:b, None, None]
    anyfront = (d1 > 0.1).any(2)                       # some point of j in front of p
    d2 = np.einsum('pkc,jc->pjk', pts[a:b], normal) - dist[None, :, None]
    anyback = (d2 < -0.1).any(2)                       # some point of p behind j
    front[a:b] = anyfront & anyback
front[np.arange(P), np.arange(P)] = False
print("front", round(time.time() - t0, 1), "s, P", P, "nc", nc, "density", front.mean())
# flood
flood = np.zeros((P, P), bool)
for p in range(P):
    seen_leaf = np.zeros(nc, bool); stack = [leaf[p]]; f = flood[p]; fr = front[p]
    while stack:
        l = stack.pop()
        if seen_leaf[l]: continue
        seen_leaf[l] = True
        for q in leaf_portals[l]:
            if fr[q] and not f[q]:
                f[q] = True; stack.append(leaf[q])
might = flood.sum(1)
print("flood done", round(time.time() - t0, 1), "s; mean mightsee", might.mean())
order = np.argsort(might, kind='stable'); rank = np.empty(P, np.int64); rank[order] = np.arange(P)
# level of k = 1 + max level of j in flood[k] with rank[j] < rank[k]
level = np.zeros(P, np.int64)
for k in order:
    js = np.flatnonzero(flood[k] & (rank < rank[k]))
    level[k] = 1 + (level[js].max() if len(js) else 0)
print("levels", level.max(), "of", P, "portals; widest level", np.bincount(level).max(), "; mean width", P / level.max())
