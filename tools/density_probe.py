import sys, json, tempfile
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from pathlib import Path
import bench, golden_util as G
from qengine_b200 import qrad as Q
for name in sys.argv[1:]:
    _,flags,_=bench.WORKLOADS[name]
    p=G.params_from_flags(flags)
    with tempfile.TemporaryDirectory() as tmp:
        bsp=bench.stage_workload(name, Path(tmp))
        sc=Q.Scene(bsp).prepare(p); dev=Q.Device(0); dev.upload(sc)
        dev.build_facelights(p.extrasamples,p.numbounce); dev.make_transfers(0)
        st=dev.stats()
        print(name, 'nnz',st['total_transfers'],'nzwords',st['nonzero_words'],'bits/word',st['total_transfers']/max(1,st['nonzero_words']),'padded',st['padded_entries'], 'pad ratio', st['padded_entries']/st['total_transfers'])
        dev.bounce(p.numbounce, want_added=False); print(dev.stats()['ms_bounce_kernel'])
