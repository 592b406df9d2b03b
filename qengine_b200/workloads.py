"""The BASELINE.json configurations as committed inputs: bench_maps/*.bsp.gz are deterministic synthetic maps
(qengine_b200.mapgen) compiled by the reference qbsp3 + qvis3 (tools/build_bench_maps.py), so that the reference
qrad3 and libqrad_cuda light identical BSPs.  Used by bench.py and by the large-config parity tests."""
from __future__ import annotations

import gzip
import shutil
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
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
                         "synthetic 6x6 room grid (~13k patches), -chop 32, 8 bounces"),
}
DEFAULT_WORKLOAD = "c5_grid16_chop8"


def stage_workload(name: str, tmp: Path) -> Path:
    """Unpack bench_maps/<bsp>.gz into tmp/assets/maps with the synthetic palette/textures beside it."""
    from . import mapgen
    bspname, _, _ = WORKLOADS[name]
    assets = mapgen.write_assets(Path(tmp))
    dst = assets / "maps" / bspname
    src = MAPS / (bspname + ".gz")
    if not src.exists():
        raise FileNotFoundError(f"{src} is missing (tools/build_bench_maps.py builds it with the reference qbsp3/qvis3)")
    with gzip.open(src, "rb") as f, open(dst, "wb") as g:
        shutil.copyfileobj(f, g)
    return dst
