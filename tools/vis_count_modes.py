import json, os, sys, tempfile
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from qengine_b200 import qrad as Q
from qengine_b200 import workloads as W
name = "c5_grid16_chop8"
with tempfile.TemporaryDirectory() as tmp:
    bsp = W.stage_workload(name, Path(tmp))
    p = Q.params_from_flags(W.WORKLOADS[name][1])
    sc = Q.Scene(bsp).prepare(p)
    dev = Q.Device(0).upload(sc)
    for m in [int(x) for x in sys.argv[1:]]:
        os.environ["QRAD_VIS_MODE"] = str(m)
        dev.set_count_nodes(False); dev.make_transfers(0); t = dev.stats()["ms_transfers_trace"]
        dev.set_count_nodes(True); dev.make_transfers(0); st = dev.stats()
        print(json.dumps({"mode": m, "trace_ms": t, "count_trace_ms": st["ms_transfers_trace"], "rays": st["rays_transfers"], "full_walks": st["full_walks"], "boxed": st["box_pruned_rays"], "visits": st["node_visits"]}), flush=True)
