#!/usr/bin/env python
"""Benchmark of the qrad3 hot path (BASELINE.json metric) -- see DESIGN.md "Measurement".

One step = lighting one map: BuildFacelights + MakeTransfers + BounceLight + FinalLightFace.
  value  : visibility + shadow rays per second of the device phases, inputs already resident in HBM
           (BSP, patches, lights, faces uploaded before the timed region).  Maps are larger than L2 in
           their transfer matrix; nothing is cached between steps (every step recomputes all phases).
  e2e    : the same metric through the public C entry point qrad_rad_world() with HOST buffers: the
           timed region contains every host->device upload and the device->host lump read-back.
  roofline: the BounceLight gather SpMV (k_bounce), algorithmic bytes = nnz*b + 64*N per bounce
            over its CUDA-event time, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline / --impl reference: the reference qrad3 (oracle/_ref, compiled from /root/reference)
           on a bounded sample of the same generator family, 1 host core (the reference has no Linux
           threading: SURVEY.md finding F2).

Launch:  python bench.py [--gpus N --steps K --warmup W] ; N > 1 under torchrun (one rank per GPU).
"""
from __future__ import annotations

import argparse
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

METRIC = "qrad3 s/map; visibility+shadow rays/s; BounceLight SpMV GB/s vs HBM peak"
MAPS = ROOT / "bench_maps"

# name -> (bsp file, qrad3 flags, description)
WORKLOADS = {
    "c3_grid13": ("grid13r512.bsp", ["-bounce", "8", "-chop", "32"],
                  "synthetic 13x13 grid of 512-unit rooms (~8k faces / ~150k patches; 14x14 exceeds MAX_MAP_LIGHTING), "
                  "-chop 32, 8 bounces [BASELINE configs[2]]"),
    "c5_grid16_chop8": ("grid16c128.bsp", ["-bounce", "16", "-chop", "8"],
                        "synthetic 16x16 room grid, qbsp3 -chop 128, qrad3 -chop 8 (~1M patches), 16 bounces [configs[4]]"),
    "c5_chop16": ("grid16c128.bsp", ["-bounce", "16", "-chop", "16"],
                  "the C5 map at -chop 16 (~250k patches): profiling stand-in whose kernels fit an ncu replay"),
    "c4_hall2000": ("hall2000.bsp", ["-bounce", "0", "-extra"],
                    "synthetic direct-light stress hall: 2 000 light / light_spot entities, -extra, direct only [configs[3]]"),
    "probe_grid10": ("grid10.bsp", ["-bounce", "8", "-chop", "32"],
                     "synthetic 10x10 room grid (~37k patches), -chop 32, 8 bounces"),
    "ref_sample_grid6": ("grid6.bsp", ["-bounce", "8", "-chop", "32"],
                         "synthetic 6x6 room grid (~13k patches), -chop 32, 8 bounces (bounded CPU sample)"),
}
DEFAULT_WORKLOAD = "c5_grid16_chop8"
REF_SAMPLE = "ref_sample_grid6"


def stage_workload(name: str, tmp: Path) -> Path:
    """Unpack bench_maps/<bsp>.gz into tmp/assets/maps with the synthetic palette/textures beside it."""
    from qengine_b200 import mapgen
    bspname, _, _ = WORKLOADS[name]
    assets = mapgen.write_assets(tmp)
    dst = assets / "maps" / bspname
    src = MAPS / (bspname + ".gz")
    if not src.exists():
        raise SystemExit(f"{src} is missing (tools/build_bench_maps.py builds it with the reference qbsp3/qvis3)")
    with gzip.open(src, "rb") as f, open(dst, "wb") as g:
        shutil.copyfileobj(f, g)
    return dst


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu: int):
        self.gpu = gpu
        self.rows = []
        self.proc = None

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        sm.sort()
        # median over samples under load (upper half of the samples)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def run_reference_cpu(workload: str, steps: int, warmup: int):
    """Time the reference qrad3 (oracle/_ref/ref_harness, built from /root/reference) on one host core."""
    import refdump as R
    if not R.have_ref():
        return None
    _, flags, desc = WORKLOADS[workload]
    times, info = [], None
    with tempfile.TemporaryDirectory() as tmp:
        bsp = stage_workload(workload, Path(tmp))
        unlit = bsp.read_bytes()
        for i in range(warmup + steps):
            bsp.write_bytes(unlit)
            t0 = time.perf_counter()
            info = R.run_ref_harness(bsp, flags)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    rc = MAPS / "raycounts.json"
    rays = json.loads(rc.read_text()).get(workload) if rc.exists() else None
    if rays is None:
        return None
    return dict(times=times, info=info, rays=rays, desc=desc, flags=flags)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("QRAD_BENCH_WORKLOAD", DEFAULT_WORKLOAD), choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        r = run_reference_cpu(REF_SAMPLE, max(1, args.steps), 0)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (reference build) not present on this box"}))
            return
        dt = sum(r["times"]) / len(r["times"])
        rays = r["rays"]["rays_transfers"] + r["rays"]["rays_direct"]
        val = rays / dt
        sample = f"{REF_SAMPLE}: {r['desc']}; {r['info']['num_patches']} patches, {rays} rays, {dt:.2f} s per step"
        line = {
            "metric": METRIC, "value": val, "unit": "rays/s", "n_gpus": args.gpus, "steps": len(r["times"]), "warmup": 0,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOADS[args.workload][2], "reference_sample": sample},
            "s_per_map": dt,
            "cpu_baseline": {"value": val, "unit": "rays/s", "cores": 1, "kind": "reference", "sample": sample},
            "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from qengine_b200 import qrad as Q

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    _, flags, desc = WORKLOADS[args.workload]
    import golden_util as G
    p = G.params_from_flags(flags)
    tmp = Path(tempfile.mkdtemp(prefix="qrad_bench_"))
    try:
        bsp = stage_workload(args.workload, tmp)
        sc = Q.Scene(bsp).prepare(p)
        counts = sc.counts()
        dev = Q.Device(local_rank)
        if world > 1:
            dev.set_shard(rank, world)

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        def device_step():
            if world > 1:
                return Q.rad_world_sharded(sc, dev, p, rank, world, upload=False)
            dev.build_facelights(p.extrasamples, p.numbounce)
            if p.numbounce > 0:
                dev.make_transfers(p.nopvs)
                dev.bounce(p.numbounce, want_added=False)
            return dev.final_light(p.ambient, p.maxlight, p.lightscale, p.subdiv, p.numbounce)

        # ---- device-resident value ----
        dev.upload(sc)
        for _ in range(args.warmup):
            device_step()
        barrier()
        launches0 = dev.stats()["kernel_launches"]
        bounce_ms, trace_ms, phase = [], [], []
        step_wall = []
        with ClockSampler(local_rank) as clk:
            t0 = time.perf_counter()
            for _ in range(args.steps):
                barrier()
                ts = time.perf_counter()
                device_step()
                barrier()
                step_wall.append(time.perf_counter() - ts)
                st = dev.stats()
                bounce_ms.append(st["ms_bounce_kernel"])
                trace_ms.append(st["ms_transfers_trace"])
                phase.append([st["ms_facelights"], st["ms_transfers"], st["ms_bounce"], st["ms_final"]])
            barrier()
            wall = time.perf_counter() - t0
        st = dev.stats()
        launches = st["kernel_launches"] - launches0
        # N = 1: the library times each phase with CUDA events on its own stream; the step time is their sum.
        # N > 1: the exchanges run on torch's NCCL streams between the phases, so the step is timed from barrier +
        # synchronize to barrier + synchronize and the max over ranks is taken.
        if world > 1:
            step_ms = float(np.mean(step_wall)) * 1e3
        else:
            step_ms = float(np.mean([sum(x) for x in phase]))
        t = torch.tensor([step_ms, float(st["rays_transfers"])], device="cuda", dtype=torch.float64)
        if world > 1:
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            step_ms = float(tmax[0].item())
            st["rays_transfers"] = int(t[1].item())     # each rank traced its own granules of the matrix
        rays = st["rays_transfers"] + st["rays_direct"]
        value = rays / (step_ms * 1e-3)

        # ---- end to end through qrad_rad_world (uploads + lump read-back inside the timed region) ----
        _, pv, lv, fv = sc.views()
        bv = sc.views()[0]
        h2d = (bv.numplanes * 20 + bv.numnodes * 28 + bv.numleafs * 28 + bv.visdatasize
               + pv.num_patches * (12 * 5 + 4 * 4 + 1 + 24) + pv.numfaces * 4
               + lv.numlights * (4 * 4 + 36) + (lv.numclusters + 1) * 4
               + fv.numfaces * (1 + 12 * 3 + 24 + 8 * 4 + 24 + 4 * 3))
        e2e_times = []
        for i in range(1 + max(1, min(args.steps, 3))):   # 1 untimed + up to 3 timed end-to-end maps
            barrier()
            t0 = time.perf_counter()
            if world > 1:
                data, lightofs, styles = Q.rad_world_sharded(sc, dev, p, rank, world)
                sc.store_lighting(data, lightofs, styles)
            else:
                Q.rad_world(sc, dev, p, want_added=False)
            lump = sc.lighting()
            barrier()
            if i:
                e2e_times.append(time.perf_counter() - t0)
        e2e_s = float(np.mean(e2e_times))
        t = torch.tensor([e2e_s], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
        d2h = len(lump) + counts["numfaces"] * 8
        import hashlib
        lump_md5 = hashlib.md5(bytes(lump)).hexdigest()   # the same at every GPU count: sharding changes no byte

        # ---- roofline of the dominant bandwidth kernel (k_bounce) ----
        peak, peak_src = measured_peak()
        b = 4 if st["num_patches"] <= 65535 else 6   # transfer_t is 4 B up to 65 535 patches
        # per GPU: a launch of k_bounce processes this rank's share of the matrix (1/world of the slices)
        alg_bytes = (st["total_transfers"] * b + 64 * st["num_patches"]) / world
        kb_ms = float(np.mean(bounce_ms))
        achieved = alg_bytes / (kb_ms * 1e-3) / 1e9 if kb_ms > 0 else 0.0
        stored_bytes = st["transfer_bytes"]     # what this rank's launch actually streams (dense-slice format)
        roofline = {"bound": "hbm", "kernel": "k_bounce", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "stored_bytes_per_launch": stored_bytes,
                    "stored_gbs": stored_bytes / (kb_ms * 1e-3) / 1e9 if kb_ms > 0 else 0.0,
                    "stored_frac": stored_bytes / (kb_ms * 1e-3) / 1e9 / peak if kb_ms > 0 else 0.0,
                    "note": "achieved = algorithmic bytes (nnz*b + 64*N, SURVEY 8d) / CUDA-event time of k_bounce. The "
                            "stored dense-slice layout is smaller than the algorithmic record (3.6 B against 6 B per "
                            "transfer on C5), so frac can exceed 1; stored_frac = bytes the launch really streams / time "
                            "/ peak is the HBM utilisation, and traffic is ncu's DRAM bytes of one launch. The step "
                            "is dominated by k_visibility (trace_kernel_ms), which is instruction-issue bound, not "
                            "HBM bound: see trace_rays_per_s and profiles/",
                    "ms_per_launch": kb_ms, "launches_per_step": p.numbounce, "per_gpu": world > 1}
        prof = ROOT / "profiles" / "bounce_traffic.json"
        if prof.exists():
            try:
                roofline["traffic"] = json.loads(prof.read_text()).get(args.workload)
            except Exception:
                pass

        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "flags": " ".join(flags), "patches": st["num_patches"],
                       "faces": counts["numfaces"], "luxels": st["num_luxels"], "clusters": st["num_clusters"],
                       "transfers_nnz": st["total_transfers"], "rays_per_step": rays,
                       "l2_policy": "inputs larger than L2 / every phase recomputed each step",
                       "lump_bytes": len(lump), "lump_md5": lump_md5,
                       "parallelism": f"rows x{world}"},
            "s_per_map": step_ms * 1e-3,
            "phase_ms": dict(zip(["facelights", "transfers", "bounce", "final"], [float(x) for x in np.mean(phase, axis=0)])),
            "exchange_ms": getattr(dev, "exchange_ms", None),
            "trace_kernel_ms": float(np.mean(trace_ms)),
            "transfers_breakdown_ms": {k: st[k] for k in ("ms_tr_setup", "ms_transfers_trace", "ms_tr_lists", "ms_tr_rows", "ms_tr_columns")},
            "trace_rays_per_s": st["rays_transfers"] / (float(np.mean(trace_ms)) * 1e-3) if np.mean(trace_ms) > 0 else None,
            "bounce_gbs": achieved,
            "wall_s_per_step": wall / args.steps,
            "e2e": {"value": rays / e2e_s, "unit": "rays/s", "s_per_map": e2e_s, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "clocks": clk.summary(),
        }
        if rank == 0 and not args.no_cpu_baseline and world == 1:
            r = run_reference_cpu(REF_SAMPLE, 1, 0)
            if r is not None:
                dt = r["times"][0]
                rr = r["rays"]["rays_transfers"] + r["rays"]["rays_direct"]
                line["cpu_baseline"] = {
                    "value": rr / dt, "unit": "rays/s", "cores": 1, "kind": "reference",
                    "sample": f"{REF_SAMPLE}: {r['desc']}; {r['info']['num_patches']} patches, {rr} rays, {dt:.2f} s "
                              f"(reference qrad3 is single-threaded on Linux)",
                    "s_per_map": dt,
                    "phases_s": {k: r["info"].get(k) for k in ("facelights_s", "transfers_s", "bounce_s", "final_s")},
                }
        if rank == 0:
            print(json.dumps(line))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
