#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU): RadWorld spread over the ranks (exchanges inside
libqrad_cuda, NCCL) must write the same lighting lump and `added` energies as the reference on every rank.  usage:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/multi_gpu_check.py [golden case | big golden (c3_grid13, c4_hall2000, c4_hall2000_b2) ...]
"""
import gzip
import json
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import golden_util as G  # noqa: E402
from qengine_b200 import qrad as Q  # noqa: E402
from qengine_b200 import workloads as W  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for case in sys.argv[1:] or ["samplec", "grid2", "grid3_chop16", "hall", "styles"]:
    with tempfile.TemporaryDirectory() as tmp:
        if (G.GOLDEN / (case + ".lump.gz")).exists():     # large configs: fixtures of tests/golden/make_golden_big.py
            meta = json.loads((G.GOLDEN / (case + ".json")).read_text())
            flags = meta["flags"]
            bsp = W.stage_workload(meta["workload"], Path(tmp))
            want = np.frombuffer(gzip.open(G.GOLDEN / (case + ".lump.gz")).read(), np.uint8)
            want_nt = np.frombuffer(gzip.open(G.GOLDEN / (case + ".numtransfers.gz")).read(), "<i4") \
                if (G.GOLDEN / (case + ".numtransfers.gz")).exists() else None
        else:
            meta, flags = G.load_case(case)
            bsp = G.stage_case(case, Path(tmp))
            want = np.frombuffer((G.GOLDEN / (case + ".lump")).read_bytes(), np.uint8)
            want_nt = G.golden_numtransfers(case)
        p = G.params_from_flags(flags)
        sc = Q.Scene(bsp).prepare(p)
        dev = Q.Device(local)
        Q.join_communicator(dev, rank, world)
        added = Q.rad_world(sc, dev, p)
        data = np.frombuffer(sc.lighting(), np.uint8)
        same = data.size == want.size and bool((data == want).all()) and \
            bool(np.array_equal(added, np.array(meta["added"], np.float32)))
        same_nt = True
        if want_nt is not None and p.numbounce > 0:
            same_nt = bool(np.array_equal(dev.numtransfers(), want_nt))
        st = dev.stats()
        print(f"rank {rank}/{world} {case}: lump identical={same} numtransfers identical={same_nt} "
              f"trace {st['ms_transfers_trace']:.1f} ms exchange {st['ms_x_matrix']:.2f} ms", flush=True)
        ok &= same and same_nt
        dev.close()
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTI_GPU_CHECK", "PASS" if t.item() == 1 else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if t.item() == 1 else 1)
