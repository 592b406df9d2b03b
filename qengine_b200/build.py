"""In-tree build of libqrad_cuda.so (+ the qrad3 executable) for sm_100a.

nvcc cross-compiles without a GPU.  -fmad=false keeps every float/double operation a
separately rounded IEEE op, as in the reference's x86-64 build (SURVEY.md finding F6);
-prec-div / -prec-sqrt are nvcc's defaults and are stated for clarity.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = HERE / "_build"
LIB = HERE / "libqrad_cuda.so"
EXE = HERE / "qrad3"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-Xptxas", "-v"]
CXX_FLAGS = ["-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-Wall", "-I/usr/local/cuda/include"]
CU = ["api.cu", "transfers.cu", "direct.cu", "final.cu", "comm.cu", "qvis.cu"]
CPP = ["host_bsp.cpp", "host_setup.cpp", "host_api.cpp", "host_multi.cpp", "plan.cpp", "host_walkers.cpp", "qvis_host.cpp"]


def _newer(src: Path, dst: Path, deps=()) -> bool:
    if not dst.exists():
        return True
    t = dst.stat().st_mtime
    return any(p.stat().st_mtime > t for p in (src, *deps))


def _run(cmd, log=None):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if log is not None:
        log.write_text(r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-4000:] + r.stderr[-8000:])
        raise RuntimeError("build step failed: " + " ".join(map(str, cmd)))
    return r


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    hdrs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.hpp")) + list((HERE.parent / "include").glob("*.h"))
    objs = []
    for name in CU:
        src, obj = CSRC / name, OBJ / (name + ".o")
        if force or _newer(src, obj, hdrs):
            if verbose:
                print("nvcc", name)
            _run([NVCC, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)], log=OBJ / (name + ".ptxas.log"))
        objs.append(obj)
    for name in CPP:
        src, obj = CSRC / name, OBJ / (name + ".o")
        if force or _newer(src, obj, hdrs):
            if verbose:
                print("g++", name)
            _run(["g++", *CXX_FLAGS, "-c", str(src), "-o", str(obj)])
        objs.append(obj)
    if force or any(_newer(o, LIB) for o in objs):
        _run([NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
              "-cudart", "static"])
    if force or _newer(CSRC / "qrad3_main.cpp", EXE, [LIB]):
        _run(["g++", "-O2", "-o", str(EXE), str(CSRC / "qrad3_main.cpp"), "-L" + str(HERE), "-lqrad_cuda",
              "-Wl,-rpath,$ORIGIN"])
    if force or _newer(CSRC / "qvis3_main.cpp", HERE / "qvis3", [LIB]):
        _run(["g++", "-O2", "-o", str(HERE / "qvis3"), str(CSRC / "qvis3_main.cpp"), "-L" + str(HERE), "-lqrad_cuda",
              "-Wl,-rpath,$ORIGIN"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
