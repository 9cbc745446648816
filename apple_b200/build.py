"""Builds ``apple_b200/libapple_b200.so`` in-tree with nvcc for sm_100a.

Every translation unit is compiled with ``-gencode arch=compute_100a,code=sm_100a -lineinfo`` (no
other architecture, no JIT cache) and the objects are linked into one shared library next to this
file, so that the library travels with the source tree.  ``python -m apple_b200.build`` or
``apple_b200.build.build()``; rebuilds only what is stale.
"""

from __future__ import annotations

import concurrent.futures
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
# tuning experiments: APL_BUILD_TAG=x APL_NVCC_FLAGS="-DAPL_SLOT_BUFS=1" builds libapple_b200_x.so next to the product
# library (own object directory); APL_LIB=<path> makes apple_b200._lib load it.  Never the product build.
_TAG = os.environ.get("APL_BUILD_TAG", "")
OBJ = CSRC / ("build_" + _TAG if _TAG else "build")
LIB = ROOT / (f"libapple_b200_{_TAG}.so" if _TAG else "libapple_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
COMMON = [
    "-std=c++17",
    "-O3",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
    "-Xcompiler",
    "-O3",
]
COMMON += os.environ.get("APL_NVCC_FLAGS", "").split()


def _units():
    units = [
        ("tiling", "tiling.cpp", []),
        ("capi", "capi.cu", []),
        ("pncg", "pncg.cu", []),
        ("pcg", "pcg.cu", []),
        ("setup", "setup.cu", []),
        ("xchg", "xchg.cu", []),
    ]
    for tname, t in (("f32", "float"), ("f64", "double")):
        for kind in (0, 1, 2, 3):
            units.append((f"fem_{tname}_k{kind}", "fem_inst.cu", [f"-DAPL_INST_T={t}", f"-DAPL_INST_KIND={kind}"]))
            # the opt-in supersets (block off-diagonals / PSD projection) in their own unit: see csrc/fem_inst.cu
            units.append((f"fem_{tname}_k{kind}_sup", "fem_inst.cu",
                          [f"-DAPL_INST_T={t}", f"-DAPL_INST_KIND={kind}", "-DAPL_INST_SUPERSET"]))
    return units


def _deps():
    return [p for p in CSRC.iterdir() if p.suffix in (".h", ".cuh")] + [ROOT.parent / "include" / "apple_b200.h"]


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def _compile(name: str, src: str, defs) -> str:
    obj = OBJ / f"{name}.o"
    source = CSRC / src
    if _stale(obj, [source, *_deps()]):
        cmd = [NVCC, *COMMON, *defs, "-x", "cu", "-c", str(source), "-o", str(obj)]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed for {name}:\n{' '.join(cmd)}\n{proc.stdout}\n{proc.stderr}")
    return str(obj)


def build(verbose: bool = False, jobs: int | None = None) -> Path:
    OBJ.mkdir(parents=True, exist_ok=True)
    units = _units()
    jobs = jobs or min(len(units), os.cpu_count() or 4)
    with concurrent.futures.ThreadPoolExecutor(jobs) as pool:
        objs = list(pool.map(lambda u: _compile(*u), units))
    if _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *objs]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError(f"link failed:\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        print(f"built {LIB} ({LIB.stat().st_size / 1e6:.1f} MB)")
    return LIB


if __name__ == "__main__":
    build(verbose=True)
    sys.exit(0)
