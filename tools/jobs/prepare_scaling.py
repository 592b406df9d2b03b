#!/usr/bin/env python
"""Host prepare (MakePatches + SubdividePatches + CreateDirectLights) of a workload for several thread counts, alone and
with `procs` copies running at once (what N processes per box do).  usage: prepare_scaling.py [workload] [procs]"""
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from qengine_b200 import qrad as Q  # noqa: E402
from qengine_b200 import workloads as W  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else W.DEFAULT_WORKLOAD
if len(sys.argv) > 3:   # child: prepare with argv[3] threads, three times, print the best
    threads = int(sys.argv[3])
    with tempfile.TemporaryDirectory() as tmp:
        bsp = W.stage_workload(name, Path(tmp))
        p = Q.params_from_flags(W.WORKLOADS[name][1])
        p.threads = threads
        best = 1e9
        for _ in range(4):
            sc = Q.Scene(bsp)
            t0 = time.perf_counter()
            sc.prepare(p)
            best = min(best, time.perf_counter() - t0)
        print(f"{best * 1e3:.1f}")
    sys.exit(0)
procs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ncpu = os.cpu_count()
print("cpus", ncpu)
for t in (ncpu, ncpu // 2, ncpu // 4, ncpu // 8, 4, 1):
    t = max(1, t)
    out = subprocess.run([sys.executable, __file__, name, "1", str(t)], capture_output=True, text=True).stdout.strip()
    print(f"alone, {t} threads: {out} ms")
t = max(1, ncpu // procs)
ps = [subprocess.Popen([sys.executable, __file__, name, "1", str(t)], stdout=subprocess.PIPE, text=True) for _ in range(procs)]
print(f"{procs} at once, {t} threads each:", [p.communicate()[0].strip() for p in ps], "ms")
