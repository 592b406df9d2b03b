"""N > 1 path on CPU.  The geometry of the transfer matrix and of its distribution over ranks is product code
(qengine_b200/csrc/plan.cpp, the object make_transfers itself builds) with C entry points that need no GPU, so these
tests drive the library's own addressing:

  * structural properties of the plan at 1, 2, 3, 8 ranks: every word of the matrix has exactly one home, row groups and
    receiver slices are each owned once, the pieces the ranks exchange tile the row buffers;
  * the exchange protocol over gloo, world size 2: every rank fills its row buffer with a known function of (source,
    slice), the ranks exchange the contiguous pieces the plan names, and every rank then walks its slices' column lists
    exactly as pass 3 does -- each word must be the one the source's owner computed, and the sources must come in
    ascending patch index (the order ShootLight adds them, qrad3.c:419-441).
"""
from __future__ import annotations

import ctypes as C
import os
import socket
import struct

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_util as G
from qengine_b200 import qrad as Q


class Plan:
    def __init__(self, cluster, visdata, world, rank, nopvs=0):
        L = Q.lib()
        L.qrad_plan_last_error.restype = C.c_char_p
        L.qrad_plan_word_address.restype = C.c_int64
        L.qrad_plan_destroy.restype = None
        self.L = L
        self.h = C.c_void_p()
        cl = np.ascontiguousarray(cluster, np.int32)
        self.n = cl.size
        self.world = world
        rc = L.qrad_plan_create(C.c_int32(cl.size), cl.ctypes.data_as(C.POINTER(C.c_int32)), visdata, C.c_int32(len(visdata)),
                                C.c_int(nopvs), C.c_int(world), C.c_int(rank), C.byref(self.h))
        assert rc == 0, L.qrad_plan_last_error()
        info = (C.c_int64 * 8)()
        L.qrad_plan_info(self.h, info)
        (self.slots, self.nslices, self.rowbuf_words, self.ngroups, self.own_groups, self.runs, self.nc,
         self.total_words) = list(info)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.qrad_plan_destroy(self.h)
            self.h = None

    def slices(self):
        lo = np.zeros(self.world, np.int32)
        hi = np.zeros(self.world, np.int32)
        self.L.qrad_plan_slices(self.h, lo.ctypes.data_as(C.POINTER(C.c_int32)), hi.ctypes.data_as(C.POINTER(C.c_int32)))
        return lo, hi

    def slot_tables(self):
        owner = np.zeros(self.slots, np.int32)
        pslot = np.zeros(self.n, np.int32)
        self.L.qrad_plan_slots(self.h, owner.ctypes.data_as(C.POINTER(C.c_int32)), pslot.ctypes.data_as(C.POINTER(C.c_int32)))
        return owner, pslot

    def piece(self, q, r):
        first, count = C.c_int64(), C.c_int64()
        assert self.L.qrad_plan_piece(self.h, C.c_int(q), C.c_int(r), C.byref(first), C.byref(count)) == 0
        return first.value, count.value

    def word_address(self, slot, sl):
        return self.L.qrad_plan_word_address(self.h, C.c_int32(slot), C.c_int32(sl))

    def column_list(self, sl):
        cap = self.n
        slots = np.zeros(cap, np.int32)
        owner = np.zeros(cap, np.int32)
        addr = np.zeros(cap, np.int64)
        cnt = C.c_int64()
        self.L.qrad_plan_column_list(self.h, C.c_int32(sl), slots.ctypes.data_as(C.POINTER(C.c_int32)),
                                     owner.ctypes.data_as(C.POINTER(C.c_int32)), addr.ctypes.data_as(C.POINTER(C.c_int64)),
                                     C.c_int64(cap), C.byref(cnt))
        k = cnt.value
        return slots[:k], owner[:k], addr[:k]


def scene_inputs(case, tmp):
    """patch clusters (host half of the product) and the raw visibility lump of a golden map"""
    meta, flags = G.load_case(case)
    bsp = G.stage_case(case, tmp)
    sc = Q.Scene(bsp).prepare(G.params_from_flags(flags))
    cluster = sc.patches()["cluster"]
    d = bsp.read_bytes()
    ofs, ln = struct.unpack_from("<ii", d, 8 + 3 * 8)
    return cluster, d[ofs:ofs + ln]


def word_value(slot, sl):
    return np.uint32((slot * 2654435761 + sl * 40503 + 12345) & 0xffffffff)


@pytest.mark.parametrize("case", ["samplec", "grid3_chop16", "manystyles"])
def test_plan_does_not_depend_on_its_threads(case, tmp_path, monkeypatch):
    """build_plan runs its two counting sorts (patches by cluster, visible pairs by receiver) on several threads; slot
    order, ownership, pieces and word addresses must come out as with one."""
    cluster, vis = scene_inputs(case, tmp_path)
    ref = None
    for threads in ("1", "3", "8"):
        monkeypatch.setenv("QRAD_PLAN_THREADS", threads)
        p = Plan(cluster, vis, 3, 1)
        owner, pslot = p.slot_tables()
        lo, hi = p.slices()
        rng = np.random.default_rng(11)
        slots = rng.integers(0, p.slots, 400)
        sls = rng.integers(0, p.nslices, 400)
        addr = np.array([p.word_address(int(a), int(b)) for a, b in zip(slots, sls)])
        cols = [np.concatenate(p.column_list(int(sl))) for sl in range(0, p.nslices, max(1, p.nslices // 8))]
        got = (p.slots, p.nslices, p.rowbuf_words, p.ngroups, p.total_words, owner.tobytes(), pslot.tobytes(), lo.tobytes(),
               hi.tobytes(), [p.piece(q, r) for q in range(3) for r in range(3)], addr.tobytes(), [c.tobytes() for c in cols])
        if ref is None:
            ref = got
        assert got == ref, threads


@pytest.mark.parametrize("case", ["samplec", "grid3_chop16", "angled"])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_every_word_has_one_home(case, world, tmp_path):
    cluster, vis = scene_inputs(case, tmp_path)
    plans = [Plan(cluster, vis, world, r) for r in range(world)]
    p0 = plans[0]
    lo, hi = p0.slices()
    assert lo[0] == 0 and hi[-1] == p0.nslices and (lo[1:] == hi[:-1]).all() and (hi >= lo).all()
    owner, pslot = p0.slot_tables()
    assert sum(p.rowbuf_words for p in plans) == p0.total_words
    # every rank computes the same plan
    for p in plans[1:]:
        assert (p.slots, p.nslices, p.total_words) == (p0.slots, p0.nslices, p0.total_words)
        o2, s2 = p.slot_tables()
        assert np.array_equal(o2, owner) and np.array_equal(s2, pslot)
    # addresses: distinct inside each owner's buffer and inside its bounds; invisible pairs have no word
    seen = [set() for _ in range(world)]
    slots = pslot[pslot >= 0]
    step = max(1, len(slots) // 150)
    nwords = 0
    for s in slots[::step]:
        q = owner[s]
        for sl in range(p0.nslices):
            a = p0.word_address(int(s), sl)
            if a < 0:
                continue
            assert 0 <= a < plans[q].rowbuf_words
            assert a not in seen[q]
            seen[q].add(a)
            nwords += 1
    assert nwords > 0
    # the pieces a rank sends tile its row buffer in slice order
    for q in range(world):
        pos = 0
        for r in range(world):
            first, count = plans[q].piece(q, r)
            assert first == pos and count >= 0
            pos += count
        assert pos == plans[q].rowbuf_words
    # column lists: ascending patch index, one entry per (visible) source, owner = the row group's owner
    slot_patch = np.full(p0.slots, -1, np.int64)
    slot_patch[pslot[pslot >= 0]] = np.nonzero(pslot >= 0)[0]
    for sl in range(0, p0.nslices, max(1, p0.nslices // 12)):
        ss, oo, aa = p0.column_list(sl)
        assert (np.diff(slot_patch[ss]) > 0).all()
        assert np.array_equal(oo, owner[ss])
        r = int(np.searchsorted(hi, sl, side="right"))
        for q in range(world):
            m = oo == q
            if m.any():
                first, count = p0.piece(q, r)
                assert (aa[m] >= 0).all() and (aa[m] < count).all()
                # the same word, addressed through the owner's buffer
                k = int(np.nonzero(m)[0][0])
                assert p0.word_address(int(ss[k]), sl) == first + aa[k]


def _worker(rank, world, port, case, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cluster, vis = scene_inputs(case, tmpdir / f"rank{rank}")
        p = Plan(cluster, vis, world, rank)
        owner, pslot = p.slot_tables()
        lo, hi = p.slices()
        # "pass 1": this rank fills the words of ITS source rows for every slice
        rowbuf = np.zeros(p.rowbuf_words, np.uint32)
        mine = [int(s) for s in pslot[pslot >= 0] if owner[s] == rank]
        for s in mine:
            for sl in range(p.nslices):
                a = p.word_address(s, sl)
                if a >= 0:
                    rowbuf[a] = word_value(s, sl)
        # the exchange before pass 3: contiguous pieces, no packing
        pieces = {}
        reqs = []
        for q in range(world):
            if q == rank:
                continue
            first, count = p.piece(rank, q)
            send = torch.from_numpy(rowbuf[first:first + count].view(np.int32).copy())
            reqs.append(dist.isend(send, q))
            _, rcount = p.piece(q, rank)
            buf = torch.zeros(rcount, dtype=torch.int32)
            pieces[q] = buf
            reqs.append(dist.irecv(buf, q))
        for r in reqs:
            r.wait()
        first_own, _ = p.piece(rank, rank)
        # "pass 3": walk every own slice's column list; each word must be what the source's owner computed
        checked = 0
        for sl in range(int(lo[rank]), int(hi[rank])):
            ss, oo, aa = p.column_list(sl)
            for s, q, a in zip(ss, oo, aa):
                w = rowbuf[first_own + a] if q == rank else np.uint32(pieces[int(q)][int(a)].item() & 0xffffffff)
                assert w == word_value(int(s), sl), (rank, sl, int(s), int(q))
                checked += 1
        t = torch.tensor([checked], dtype=torch.int64)
        dist.all_reduce(t)
        assert t.item() > 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["samplec", "grid2"])
def test_matrix_exchange_protocol_gloo(case, tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, case, tmp_path), nprocs=2, join=True)


def _prepare_worker(rank, world, port, case, tmpdir):
    """what a process-per-GPU host does before the device phases: its share of the dicing, an all-gather of the ranks'
    blobs (different lengths: sizes first, then padded payloads), the merge -- and then it must hold the scene a plain
    prepare builds"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        meta, flags = G.load_case(case)
        bsp = G.stage_case(case, tmpdir / f"rank{rank}")
        sc = Q.Scene(bsp)
        blob = sc.prepare_part(G.params_from_flags(flags), rank, world, copy=False)
        n = torch.tensor([blob.size], dtype=torch.int64)
        sizes = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(sizes, n)
        sizes = [int(x.item()) for x in sizes]
        mx = (max(sizes) + 255) & ~255
        mine = torch.zeros(mx, dtype=torch.uint8)
        mine[:blob.size] = torch.from_numpy(blob)
        parts = [torch.zeros(mx, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(parts, mine)
        sc.prepare_merge([parts[q].numpy()[:sizes[q]] for q in range(world)])
        want = Q.Scene(bsp).prepare(G.params_from_flags(flags))
        a, b = sc.patches(), want.patches()
        for k in b:
            assert np.array_equal(a[k], b[k]), (rank, k)
        (ca, pa), (cb, pb) = sc.windings(), want.windings()
        assert np.array_equal(ca, cb) and np.array_equal(pa, pb)
        la, lb = sc.lights(), want.lights()
        for k in lb:
            assert np.array_equal(la[k], lb[k]), (rank, k)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["grid3_chop16", "manystyles"])
def test_shared_prepare_over_gloo(case, tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_prepare_worker, args=(2, port, case, tmp_path), nprocs=2, join=True)
