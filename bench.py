#!/usr/bin/env python
"""Benchmark of the qrad3 hot path (BASELINE.json metric) -- see DESIGN.md "Measurement".

One step = lighting one map: BuildFacelights + MakeTransfers + BounceLight + FinalLightFace.
  value   : visibility + shadow rays per second over the device phases, inputs already resident in HBM
            (BSP, patches, lights, faces uploaded before the timed region).  Nothing is cached between steps
            (every step recomputes all phases) and the transfer matrix is far larger than L2.
  e2e     : the same metric through the drop-in's own call sequence with HOST buffers and FILES: the timed region is
            qrad_host_load (LoadBSPFile + ParseEntities + CalcTextureReflectivity) + qrad_host_prepare (MakePatches,
            SubdividePatches, CreateDirectLights) + qrad_rad_world (every upload, all device phases, lump read-back)
            + qrad_host_write (WriteBSPFile) -- what `qrad3 map.bsp` does; bytes are counted by the library.
  roofline: the BounceLight gather SpMV (k_bounce), the one HBM-bound kernel of the path, CUDA-event timed; plus
            `roofline_visibility` for k_visibility, the kernel that is ~90 % of the step and is bound by instruction
            issue, not by memory.
  cpu_baseline / --impl reference: the REFERENCE's own MakeTransfers / BuildFacelights (oracle/_ref/ref_harness_big:
            the reference translation units with the 65 000-patch limit lifted, byte-identical on every golden) on a
            deterministic sample of the rows / faces of THE SAME MAP, 1 host core (the reference has no Linux
            threading: SURVEY.md finding F2); rays are counted by the harness itself.

Launch:  python bench.py [--gpus N --steps K --warmup W] ; N > 1 under torchrun (one rank per GPU).
"""
from __future__ import annotations

import argparse
import hashlib
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

from qengine_b200.workloads import DEFAULT_WORKLOAD, WORKLOADS, stage_workload  # noqa: E402

METRIC = "qrad3 s/map; visibility+shadow rays/s; BounceLight SpMV GB/s vs HBM peak"
REF_HARNESS = ROOT / "oracle" / "_ref" / "ref_harness_big"


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
        load = sm[len(sm) // 2:] if sm else []   # median over the samples under load (upper half)
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


def run_reference_cpu(workload: str, steps: int):
    """The reference's own MakeTransfers + BuildFacelights on a deterministic sample of the SAME map (1 host core).
    Step s lights the rows i = s mod stride and the faces f = s mod stride, so `steps` steps touch disjoint samples.
    Returns None when oracle/_ref is not on this box."""
    if not REF_HARNESS.exists():
        return None
    _, flags, desc = WORKLOADS[workload]
    with tempfile.TemporaryDirectory() as tmp:
        bsp = stage_workload(workload, Path(tmp))
        # stride 0 = the harness's automatic choice: ceil(num_patches / 128), about 128 rows (a few seconds) per step
        cmd = [str(REF_HARNESS), *flags, "-benchsteps", str(steps), "-rows", "0:0", str(bsp)]
        t0 = time.perf_counter()
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=3000)
        wall = time.perf_counter() - t0
    if r.returncode:
        raise RuntimeError("reference harness failed: " + r.stdout[-1000:] + r.stderr[-500:])
    rows = [json.loads(l[8:]) for l in r.stdout.splitlines() if l.startswith("REFSTEP ")]
    return dict(steps=rows, stride=rows[0]["stride"], desc=desc, flags=flags, wall=wall)


def reference_summary(r, use):
    """Aggregate the harness's REFSTEP lines `use` (a slice) into the bench fields."""
    st = r["steps"][use]
    secs = sum(s["facelights_s"] + s["transfers_s"] for s in st)
    rays = sum(s["rays_direct"] + s["rays_transfers"] for s in st)
    n = len(st)
    s0 = st[0]
    stride = r["stride"]
    sample = (f"reference MakeTransfers + BuildFacelights on 1/{stride} of the rows and faces of the same map per step "
              f"({s0['rows']} of {s0['num_patches']} rows, {s0['faces']} of {s0['numfaces']} faces; step s takes i = s mod "
              f"{stride}); {rays // n} rays and {secs / n:.2f} s per step; BounceLight / FinalLightFace (about 2 % of the "
              f"reference's time) are not sampled; 1 core (the reference is single-threaded on Linux)")
    return dict(value=rays / secs, ms_per_step=secs / n * 1e3, sample=sample, rays_per_step=rays / n,
                s_per_map_extrapolated=s0["setup_s"] + secs / n * stride,
                phases_s={"setup_s": s0["setup_s"], "facelights_s_sample": sum(s["facelights_s"] for s in st) / n,
                          "transfers_s_sample": sum(s["transfers_s"] for s in st) / n})


def visibility_issue_block(workload, trace_ms, rays):
    """k_visibility is bound by instruction issue (the tree is L1/L2 resident).  The issue rate comes from
    the committed ncu capture of the same kernel (profiles/visibility_issue.json); the time and rays are live."""
    out = {"bound": "issue", "kernel": "k_visibility", "ms_per_launch": trace_ms,
           "rays_per_s": rays / (trace_ms * 1e-3) if trace_ms > 0 else None, "unit": "thread-instructions/clk/SM"}
    p = ROOT / "profiles" / "visibility_issue.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            out.update({"achieved": d["warp_inst_per_clk_per_sm"] * d["lanes_active"], "peak": 4.0 * 32,
                        "frac": d["warp_inst_per_clk_per_sm"] * d["lanes_active"] / 128.0,
                        "source": d.get("source"), "measured_on": d.get("workload")})
        except Exception:
            pass
    return out


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
    _, flags, desc = WORKLOADS[args.workload]
    workload_name = f"{args.workload}: {desc}"

    if args.impl == "reference":
        if rank != 0:
            return
        steps, warm = max(1, args.steps), max(0, args.warmup)
        r = run_reference_cpu(args.workload, steps + warm)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_harness_big (reference build) not present on this box"}))
            return
        s = reference_summary(r, slice(warm, None))
        line = {
            "metric": METRIC, "value": s["value"], "unit": "rays/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": s["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name, "flags": " ".join(flags), "sampled": True},
            "s_per_map_extrapolated": s["s_per_map_extrapolated"], "rays_per_step": s["rays_per_step"],
            "cpu_baseline": {"value": s["value"], "unit": "rays/s", "cores": 1, "kind": "reference", "sample": s["sample"]},
            "e2e": {"value": s["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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

    p = Q.params_from_flags(flags)
    # host threads of the set-up (the tool's own -threads flag): the ranks of one box share its cores
    p.threads = max(1, min(64, (os.cpu_count() or 8) // max(1, world)))
    tmp = Path(tempfile.mkdtemp(prefix="qrad_bench_"))
    try:
        bsp = stage_workload(args.workload, tmp)
        unlit = bsp.read_bytes()
        t0 = time.perf_counter()
        sc = Q.Scene(bsp)
        load_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        sc.prepare(p)
        prepare_s = time.perf_counter() - t0
        counts = sc.counts()
        dev = Q.Device(local_rank)
        if world > 1:
            Q.join_communicator(dev, rank, world)

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        pinned = {}

        def gather_blobs(blob, world, xbytes):
            """all-gather of the ranks' byte blobs (different lengths) through the GPUs; xbytes += [h2d, d2h]"""
            n = torch.tensor([blob.size], device="cuda", dtype=torch.int64)
            sizes = [torch.zeros_like(n) for _ in range(world)]
            dist.all_gather(sizes, n)
            sizes = [int(x.item()) for x in sizes]
            mx = (max(sizes) + 255) & ~255
            mine = torch.zeros(mx, dtype=torch.uint8, device="cuda")
            mine[:blob.size].copy_(torch.from_numpy(blob), non_blocking=True)
            out = torch.empty(world * mx, dtype=torch.uint8, device="cuda")
            dist.all_gather_into_tensor(out, mine)
            if pinned.get("n", 0) < world * mx:
                pinned["buf"] = torch.empty(world * mx, dtype=torch.uint8, pin_memory=True)
                pinned["n"] = world * mx
            host = pinned["buf"][:world * mx]
            host.copy_(out)
            torch.cuda.synchronize()
            xbytes[0] += int(blob.size)
            xbytes[1] += world * mx
            a = host.numpy()
            return [a[q * mx:q * mx + sizes[q]] for q in range(world)]

        def device_step():
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
        bounce_ms, trace_ms, phase, step_wall = [], [], [], []
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
        # N > 1: the exchanges (NCCL, inside the library) overlap phases, so the step is timed from barrier + synchronize
        # to barrier + synchronize and the max over ranks is taken.
        step_ms = float(np.mean(step_wall)) * 1e3 if world > 1 else float(np.mean([sum(x) for x in phase]))
        t = torch.tensor([step_ms, float(st["rays_transfers"]), float(st["rays_direct"])], device="cuda", dtype=torch.float64)
        if world > 1:
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            step_ms = float(tmax[0].item())
            st["rays_transfers"] = int(t[1].item())     # each rank traced its own rows / lit its own faces
            st["rays_direct"] = int(t[2].item())
        rays = st["rays_transfers"] + st["rays_direct"]
        value = rays / (step_ms * 1e-3)
        trace_by_rank = None
        if world > 1:   # pass 1 is the bulk of the step: how evenly did it divide?
            tr = torch.tensor([float(np.mean(trace_ms))], device="cuda", dtype=torch.float64)
            allr = [torch.zeros_like(tr) for _ in range(world)]
            dist.all_gather(allr, tr)
            trace_by_rank = [round(float(x.item()), 1) for x in allr]
        # the library's own phase clocks of the last step, rank by rank (waits for other ranks show up in the phase that waits)
        PHASE_KEYS = ("ms_facelights", "ms_tr_setup", "ms_transfers_trace", "ms_tr_lists", "ms_tr_rows", "ms_tr_columns",
                      "ms_bounce", "ms_bounce_kernel", "ms_final", "ms_x_matrix", "ms_x_radiosity")
        phases_by_rank = None
        if world > 1:
            pv = torch.tensor([float(st.get(k) or 0.0) for k in PHASE_KEYS], device="cuda", dtype=torch.float64)
            allp = [torch.zeros_like(pv) for _ in range(world)]
            dist.all_gather(allp, pv)
            phases_by_rank = {k: [round(float(x[i].item()), 1) for x in allp] for i, k in enumerate(PHASE_KEYS)}

        # ---- end to end: what `qrad3 map.bsp` does, files and host buffers, every copy inside the timed region ----
        e2e_times, e2e_parts, h2d, d2h = [], [], 0, 0
        for i in range(1 + max(1, min(args.steps, 3))):   # 1 untimed + up to 3 timed maps
            bsp.write_bytes(unlit)   # every rank lights its own staged copy; rank 0 writes the result
            barrier()
            s0 = dev.stats()
            t0 = time.perf_counter()
            sc2 = Q.Scene(bsp)
            t1 = time.perf_counter()
            xb = [0, 0]
            if world > 1:
                # the processes share the dicing (90 % of prepare): every rank dices its range of the faces, the blobs
                # travel through the GPUs (NCCL all-gather), every rank stitches all of them -- the same scene everywhere
                blob = sc2.prepare_part(p, rank, world, copy=False)
                sc2.prepare_merge(gather_blobs(blob, world, xb))
            else:
                sc2.prepare(p)
            t2 = time.perf_counter()
            Q.rad_world(sc2, dev, p, want_added=False)
            t3 = time.perf_counter()
            if rank == 0:
                sc2.write(bsp)
            lump = sc2.lighting()
            barrier()
            t4 = time.perf_counter()
            s1 = dev.stats()
            if i:
                e2e_times.append(t4 - t0)
                e2e_parts.append([t1 - t0, t2 - t1, t3 - t2, t4 - t3])
                h2d, d2h = s1["h2d_bytes"] - s0["h2d_bytes"] + xb[0], s1["d2h_bytes"] - s0["d2h_bytes"] + xb[1]
            sc2.close()
        e2e_s = float(np.mean(e2e_times))
        t = torch.tensor([e2e_s], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
        parts = np.mean(e2e_parts, axis=0)
        lump_md5 = hashlib.md5(bytes(lump)).hexdigest()   # the same at every GPU count: sharding changes no byte

        # ---- roofline of the HBM-bound kernel (k_bounce) ----
        peak, peak_src = measured_peak()
        b = 4 if st["num_patches"] <= 65535 else 6   # transfer_t is 4 B up to 65 535 patches
        alg_bytes = (st["total_transfers"] * b + 64 * st["num_patches"]) / world   # per GPU: this rank's slices
        kb_ms = float(np.mean(bounce_ms))
        stored_bytes = st["transfer_bytes"]     # what this rank's launch actually streams (dense-slice format)
        sec = kb_ms * 1e-3
        roofline = {"bound": "hbm", "kernel": "k_bounce", "achieved": stored_bytes / sec / 1e9 if sec > 0 else 0.0,
                    "peak": peak, "unit": "GB/s", "frac": stored_bytes / sec / 1e9 / peak if sec > 0 else 0.0,
                    "traffic": None, "peak_source": peak_src,
                    "stored_bytes_per_launch": stored_bytes, "algorithmic_bytes_per_launch": alg_bytes,
                    "algorithmic_gbs": alg_bytes / sec / 1e9 if sec > 0 else 0.0,
                    "algorithmic_frac": alg_bytes / sec / 1e9 / peak if sec > 0 else 0.0,
                    "note": "achieved / frac = bytes the launch streams from HBM (the stored dense-slice matrix: u32 source "
                            "+ 32 x u16 weights per list entry, 3.6 B per transfer on C5) / CUDA-event time of k_bounce / "
                            "measured copy bandwidth. algorithmic_* uses SURVEY 8d's record (nnz*b + 64*N, b = 6 B): that "
                            "count is not a lower bound for this layout, so its fraction can exceed 1. traffic: ncu "
                            "dram__bytes of one launch, from the capture named in traffic_source (one GPU only; null on shards).",
                    "ms_per_launch": kb_ms, "launches_per_step": p.numbounce, "per_gpu": world > 1}
        # DRAM traffic of one launch from an ncu --set full capture: on file only for the one-GPU run of this workload
        tfile = ROOT / "profiles" / "bounce_traffic.json"
        if world == 1 and tfile.exists():
            try:
                t = json.loads(tfile.read_text()).get(args.workload)
                if t and t.get("n_gpus") == 1:
                    roofline["traffic"] = t["dram_bytes_read"] + t["dram_bytes_write"]
                    roofline["traffic_source"] = t["source"]
            except Exception:
                pass
        tr_ms = float(np.mean(trace_ms))

        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name, "flags": " ".join(flags), "patches": st["num_patches"],
                       "faces": counts["numfaces"], "luxels": st["num_luxels"], "clusters": st["num_clusters"],
                       "transfers_nnz": st["total_transfers"], "rays_per_step": rays,
                       "l2_policy": "inputs larger than L2 / every phase recomputed each step",
                       "lump_bytes": len(lump), "lump_md5": lump_md5, "parallelism": f"rows+faces x{world}"},
            "s_per_map": step_ms * 1e-3,
            "phase_ms": dict(zip(["facelights", "transfers", "bounce", "final"], [float(x) for x in np.mean(phase, axis=0)])),
            "exchange_ms": {k: st[k] for k in ("ms_x_matrix", "ms_x_small", "ms_x_radiosity")} if world > 1 else None,
            "trace_kernel_ms": tr_ms, "trace_kernel_ms_by_rank": trace_by_rank, "phases_ms_by_rank": phases_by_rank,
            "transfers_breakdown_ms": {k: st[k] for k in ("ms_tr_setup", "ms_transfers_trace", "ms_tr_lists", "ms_tr_rows", "ms_tr_columns")},
            "trace_rays_per_s": st["rays_transfers"] / (tr_ms * 1e-3) if tr_ms > 0 else None,
            "wall_s_per_step": wall / args.steps,
            "e2e": {"value": rays / e2e_s, "unit": "rays/s", "s_per_map": e2e_s, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h),
                    "parts_s": dict(zip(["host_load", "host_prepare", "rad_world", "host_write"], [float(x) for x in parts])),
                    "includes": "LoadBSPFile+ParseEntities+CalcTextureReflectivity, MakePatches+SubdividePatches+"
                                "CreateDirectLights, uploads + all device phases + lump read-back, WriteBSPFile"
                                + ("; the ranks share the dicing (qrad_host_prepare_part / _merge, blobs all-gathered "
                                   "through the GPUs, counted in the byte figures)" if world > 1 else "")},
            "host_prepare_s": float(parts[1]), "host_load_s": float(parts[0]),
            "gpu_launches": int(launches),
            "roofline": roofline,
            "roofline_visibility": visibility_issue_block(args.workload, tr_ms, st["rays_transfers"] / world),
            "clocks": clk.summary(),
        }
        if rank == 0 and not args.no_cpu_baseline and world == 1:
            r = run_reference_cpu(args.workload, 3)
            if r is not None:
                s = reference_summary(r, slice(1, None))
                line["cpu_baseline"] = {"value": s["value"], "unit": "rays/s", "cores": 1, "kind": "reference",
                                        "sample": s["sample"], "s_per_map_extrapolated": s["s_per_map_extrapolated"],
                                        "phases_s": s["phases_s"], "same_map": True}
        if rank == 0:
            print(json.dumps(line))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
