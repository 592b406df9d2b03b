"""TEST INFRASTRUCTURE ONLY (oracle/).  numpy restatement of the first stage of the reference's qvis3 -- the part that does
not depend on the order portals are processed in -- for the next stage outward of this project:

  load_prt            LoadPortals            src/tools/qvis3/qvis3.c:304-412 (two memory portals per file portal, the
                                             forward one with the flipped plane, the backward one with the reversed winding)
  plane_from_winding  PlaneFromWinding       qvis3.c:58-69 (float cross product, VectorNormalize with a double square root
                                             and reciprocal, mathlib.c:106-122)
  base_portal_vis     BasePortalVis          flow.c:605-645  (portalfront: some point of j in front of p by more than
                                             ON_EPSILON and some point of p behind j by more than ON_EPSILON; the float
                                             distance is compared with the DOUBLE 0.1, vis.h:31)
  simple_flood        SimpleFlood            flow.c:576-598  (portalflood: closure over leaves through portalfront)
  sorted_rank         SortPortals            qvis3.c:118-129 (qsort by nummightsee; glibc's qsort is a merge sort, i.e.
                                             stable, for arrays of this size)
  dependency_levels   which finished portals a portal consults in that order (flow.c:391-396), levelled

Pinned by tests/test_qvis_oracle_cpu.py against dumps of the reference itself (oracle/qvis_harness.c).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np


def load_prt(path):
    lines = Path(path).read_text().split("\n")
    if lines[0].strip() != "PRT1":
        raise ValueError("LoadPortals: not a portal file")
    nc, nfile = int(lines[1]), int(lines[2])
    windings, pairs = [], []
    for ln in lines[3:3 + nfile]:
        head, rest = ln.split("(", 1)
        n, l0, l1 = (int(x) for x in head.split())
        # "%lf" into a double, then assigned to the float vec_t
        pts = np.array([[float(v) for v in seg.split(")")[0].split()] for seg in ("(" + rest).split("(")[1:]],
                       np.float64).astype(np.float32)
        if len(pts) != n:
            raise ValueError("LoadPortals: reading portal")
        windings.append(pts)
        pairs.append((l0, l1))
    return nc, windings, pairs


def plane_from_winding(w):
    v1 = (w[2] - w[1]).astype(np.float32)
    v2 = (w[0] - w[1]).astype(np.float32)
    f = np.float32
    n = np.array([f(f(v2[1] * v1[2]) - f(v2[2] * v1[1])), f(f(v2[2] * v1[0]) - f(v2[0] * v1[2])),
                  f(f(v2[0] * v1[1]) - f(v2[1] * v1[0]))], np.float32)
    s = f(f(f(n[0] * n[0]) + f(n[1] * n[1])) + f(n[2] * n[2]))
    length = f(np.sqrt(np.float64(s)))
    if length != 0:
        il = f(np.float64(1.0) / np.float64(length))
        n = np.array([f(n[0] * il), f(n[1] * il), f(n[2] * il)], np.float32)
    d = f(f(f(w[0][0] * n[0]) + f(w[0][1] * n[1])) + f(w[0][2] * n[2]))
    return n, d


class Portals:
    def __init__(self, path):
        self.nc, wind, pairs = load_prt(path)
        P = self.P = 2 * len(wind)
        self.normal = np.zeros((P, 3), np.float32)
        self.dist = np.zeros(P, np.float32)
        self.leaf = np.zeros(P, np.int32)
        self.windings = []
        self.leaf_portals = [[] for _ in range(self.nc)]
        for i, (w, (l0, l1)) in enumerate(zip(wind, pairs)):
            n, d = plane_from_winding(w)
            self.normal[2 * i], self.dist[2 * i], self.leaf[2 * i] = -n, -d, l1       # forward portal
            self.windings.append(w)
            self.leaf_portals[l0].append(2 * i)
            self.normal[2 * i + 1], self.dist[2 * i + 1], self.leaf[2 * i + 1] = n, d, l0   # backward portal
            self.windings.append(w[::-1].copy())
            self.leaf_portals[l1].append(2 * i + 1)


def _dots(pts, normal):
    """DotProduct(point, normal) in float, left to right, for points [J, K, 3] x normals [Pn, 3] -> [Pn, J, K]"""
    f = np.float32
    x = (pts[None, :, :, 0] * normal[:, None, None, 0]).astype(f)
    y = (pts[None, :, :, 1] * normal[:, None, None, 1]).astype(f)
    z = (pts[None, :, :, 2] * normal[:, None, None, 2]).astype(f)
    return ((x + y).astype(f) + z).astype(f)


def base_portal_vis(pt: Portals, chunk=128):
    P = pt.P
    maxp = max(len(w) for w in pt.windings)
    pts = np.zeros((P, maxp, 3), np.float32)
    valid = np.zeros((P, maxp), bool)
    for i, w in enumerate(pt.windings):
        pts[i, :len(w)] = w
        valid[i, :len(w)] = True
    front = np.zeros((P, P), bool)
    for a in range(0, P, chunk):
        b = min(P, a + chunk)
        d1 = (_dots(pts, pt.normal[a:b]) - pt.dist[a:b, None, None]).astype(np.float32)          # [p, j, k]
        anyfront = ((d1.astype(np.float64) > 0.1) & valid[None]).any(2)
        d2 = (_dots(pts[a:b], pt.normal) - pt.dist[:, None, None]).astype(np.float32)             # [j, p, k]
        anyback = ((d2.astype(np.float64) < -0.1) & valid[None, a:b]).any(2).T                    # [p, j]
        front[a:b] = anyfront & anyback
    front[np.arange(P), np.arange(P)] = False
    return front


def simple_flood(pt: Portals, front):
    P = pt.P
    flood = np.zeros((P, P), bool)
    for p in range(P):
        seen = np.zeros(pt.nc, bool)
        stack = [int(pt.leaf[p])]
        f, fr = flood[p], front[p]
        while stack:
            l = stack.pop()
            if seen[l]:
                continue          # a second visit finds every fronted portal of the leaf flooded already
            seen[l] = True
            for q in pt.leaf_portals[l]:
                if fr[q] and not f[q]:
                    f[q] = True
                    stack.append(int(pt.leaf[q]))
    return flood


def sorted_rank(might):
    order = np.argsort(might, kind="stable")
    rank = np.empty(len(might), np.int64)
    rank[order] = np.arange(len(might))
    return order, rank


def dependency_levels(flood, order, rank):
    level = np.zeros(len(order), np.int64)
    for k in order:
        js = np.flatnonzero(flood[k] & (rank < rank[k]))
        level[k] = 1 + (level[js].max() if len(js) else 0)
    return level


def read_dump(path):
    """out.bin of oracle/qvis_harness.c"""
    d = Path(path).read_bytes()
    magic, nc, P, pb = np.frombuffer(d, "<i4", 4, 0)
    assert magic == 0x31445651
    rec = 16 + 12 + 3 * pb
    out = dict(nc=int(nc), P=int(P), plane=np.zeros((P, 4), np.float32), leaf=np.zeros(P, np.int32), might=np.zeros(P, np.int32),
               rank=np.zeros(P, np.int32), front=np.zeros((P, P), bool), flood=np.zeros((P, P), bool), vis=np.zeros((P, P), bool))
    for k in range(P):
        o = 16 + k * rec
        out["plane"][k] = np.frombuffer(d, "<f4", 4, o)
        out["leaf"][k], out["might"][k], out["rank"][k] = np.frombuffer(d, "<i4", 3, o + 16)
        for j, name in enumerate(("front", "flood", "vis")):
            bits = np.unpackbits(np.frombuffer(d, np.uint8, pb, o + 28 + j * pb), bitorder="little")[:P]
            out[name][k] = bits.astype(bool)
    return out


# ---- stage 2: PortalFlow through the C restatement (oracle/qvis_oracle.c) ------------------------------------------------
def _lib():
    import ctypes as C
    so = Path(__file__).resolve().parent / "libqvis_oracle.so"
    lib = C.CDLL(str(so))
    lib.qvo_portal_flow.restype = C.c_long
    return lib


def pack_bits(m):
    """bool [P, P] -> the reference's bit vectors [P, portalbytes] (portalbytes = ((P + 63) & ~63) >> 3, qvis3.c:336)"""
    P = m.shape[0]
    pb = ((P + 63) & ~63) >> 3
    out = np.zeros((P, pb), np.uint8)
    packed = np.packbits(m, axis=1, bitorder="little")
    out[:, :packed.shape[1]] = packed
    return out


class FlowOracle:
    """portal_flow(k, done_vis): the portalvis vector of portal k when every portal ranked before k is finished with the
    vectors in done_vis (bool [P, P]) and every other portal is not."""

    def __init__(self, pt: Portals, flood, rank):
        import ctypes as C
        self.C, self.pt, self.lib = C, pt, _lib()
        self.planes = np.ascontiguousarray(np.concatenate([pt.normal, pt.dist[:, None]], 1), np.float32)
        self.leaf = np.ascontiguousarray(pt.leaf, np.int32)
        cnt = np.array([len(w) for w in pt.windings], np.int32)
        self.wfirst = np.zeros(pt.P + 1, np.int32)
        self.wfirst[1:] = np.cumsum(cnt)
        self.points = np.ascontiguousarray(np.concatenate(pt.windings, 0), np.float32)
        self.leaf_first = np.zeros(pt.nc + 1, np.int32)
        self.leaf_first[1:] = np.cumsum([len(l) for l in pt.leaf_portals])
        self.leaf_portals = np.array([q for l in pt.leaf_portals for q in l], np.int32)
        self.flood = pack_bits(flood)
        self.rank = np.ascontiguousarray(rank, np.int32)
        self.pb = self.flood.shape[1]

    def portal_flow(self, k, done_bits):
        C = self.C
        out = np.zeros(self.pb, np.uint8)
        p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        chains = self.lib.qvo_portal_flow(
            C.c_int(self.pt.P), C.c_int(self.pt.nc), C.c_int(self.pb), p(self.planes, C.c_float), p(self.leaf, C.c_int32),
            p(self.wfirst, C.c_int32), p(self.points, C.c_float), p(self.leaf_first, C.c_int32), p(self.leaf_portals, C.c_int32),
            p(self.flood, C.c_uint8), p(done_bits, C.c_uint8), p(self.rank, C.c_int32), C.c_int(int(k)), p(out, C.c_uint8))
        assert chains >= 0
        return out, chains

    def full(self, order):
        """the reference's run: portals one after the other in sorted order, each seeing the earlier ones finished"""
        done = np.zeros_like(self.flood)
        for k in order:
            done[k], _ = self.portal_flow(k, done)
        return done
