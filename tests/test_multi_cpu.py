"""N > 1 path on CPU: world_size-2 gloo processes run the exchange protocol of the sharded RadWorld
(qengine_b200.qrad.rad_world_sharded) on host buffers.  The data are the reference's own transfer lists for the
sample map (tests/golden/samplec.transfers.npz); the sharded, order-exact gather bounce must reproduce the
reference's totallight bit for bit, and the shard geometry must tile the work exactly once."""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_util as G
from qengine_b200 import sharding as S


def test_task_granules_partition_exactly_once():
    for world in (1, 2, 3, 8):
        for num_tasks in (1, 2047, 2048, 2049, 100000):
            seen = np.zeros(num_tasks, int)
            for rank in range(world):
                for lt in range(S.num_local_tasks(num_tasks, rank, world)):
                    t = S.local_to_global_task(lt, rank, world)
                    if t < num_tasks:
                        assert S.task_owner(t, world) == rank
                        seen[t] += 1
            assert (seen == 1).all()


def test_slice_shards_tile():
    for world in (1, 2, 4, 8):
        for n in (1, 7, 8, 31457):
            cover = np.zeros(n, int)
            for rank in range(world):
                lo, hi, shard, padded = S.slice_shard(n, rank, world)
                cover[lo:hi] += 1
                assert shard * world == padded and padded >= n * 32 and (hi - lo) * 32 <= shard
            assert (cover == 1).all()


def _worker(rank, world, port, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        z = np.load(G.GOLDEN / "samplec.transfers.npz")
        cnt, patch, tr = z["numtransfers"], z["patch"].astype(np.int64), z["transfer"].astype(np.float32)
        n = len(cnt)
        row_ptr = np.concatenate([[0], np.cumsum(cnt)])
        src = np.repeat(np.arange(n), cnt)
        # --- exchange 1: every rank contributes the accept bits of its own rows, sum all-reduce = full matrix ---
        dense = np.zeros((n, n), np.int32)
        mine = np.arange(n) % world == rank
        m = mine[src]
        dense[src[m], patch[m]] = 1
        t = torch.from_numpy(dense)
        dist.all_reduce(t)
        full = np.zeros((n, n), np.int32)
        full[src, patch] = 1
        assert np.array_equal(t.numpy(), full)
        # --- gather form: receivers sharded in equal runs, sources ascending (the order ShootLight adds them) ---
        nsl = (n + 31) // 32
        lo, hi, shard, padded = S.slice_shard(nsl, rank, world)
        order = np.lexsort((src, patch))          # by receiver, then source
        rcv, s_src, s_w = patch[order], src[order], tr[order]
        col_ptr = np.concatenate([[0], np.cumsum(np.bincount(rcv, minlength=n))])
        # patch data and the expected answer come from the C oracle (test infrastructure), one private copy per rank
        sys.path.insert(0, str(G.ROOT / "oracle"))
        import oracle as O
        o = O.Oracle(G.stage_case("samplec", tmpdir / f"rank{rank}"), numbounce=8)
        o.make_patches()
        o.create_direct_lights()
        o.build_facelights()
        p = o.patches()
        area = p["area"].astype(np.float32)
        total = p["totallight"].astype(np.float32).copy()
        o.make_transfers()
        o.bounce()
        want = o.patches()["totallight"]
        faces = np.frombuffer(o.lumps["faces"], dtype=np.dtype([("planenum", "<u2"), ("side", "<i2"), ("firstedge", "<i4"),
                                                                 ("numedges", "<i2"), ("texinfo", "<i2"), ("styles", "u1", 4),
                                                                 ("lightofs", "<i4")]))
        prefl = o.reflectivity()[faces["texinfo"][p["face"]]].astype(np.float32)   # patch reflectivity = its texture's
        rad = np.zeros((padded, 3), np.float32)
        rad[:n] = p["samplelight"].astype(np.float32) * prefl * area[:, None]
        tot = np.zeros((padded, 3), np.float32)
        tot[:n] = total
        for b in range(8):
            send = (rad / np.float32(65536)).astype(np.float32)
            new = np.zeros((padded, 3), np.float32)
            for j in range(lo * 32, min(hi * 32, n)):
                ill = np.zeros(3, np.float32)
                for k in range(col_ptr[j], col_ptr[j + 1]):
                    ill = ill + send[s_src[k]] * s_w[k]
                tot[j] = tot[j] + ill / area[j]
                new[j] = ill * prefl[j]
            tr_new = torch.from_numpy(new)
            dist.all_gather_into_tensor(tr_new, tr_new[rank * shard:(rank + 1) * shard].clone())
            rad = tr_new.numpy().copy()
        tt = torch.from_numpy(tot)
        dist.all_gather_into_tensor(tt, tt[rank * shard:(rank + 1) * shard].clone())
        assert np.array_equal(tt.numpy()[:n], want), "sharded gather bounce differs from the reference order"
        assert np.array_equal(want, np.load(G.GOLDEN / "samplec.totallight.npy"))
    finally:
        dist.destroy_process_group()


def test_sharded_exchange_protocol_gloo(tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, tmp_path), nprocs=2, join=True)
