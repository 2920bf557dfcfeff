"""Build the in-tree native library.

  c99_vectordb_b200/_b200flat.so   the C-ABI library (include/b200_flat.h) — sm_100a CUDA

nvcc cross-compiles without a GPU.  The CUDA runtime is linked statically so the library loads
(and exports its symbols) on a box with no driver; compute entry points then fail loudly.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "c99_vectordb_b200" / "csrc"
LIB = ROOT / "c99_vectordb_b200" / "_b200flat.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "--cudart", "static",
]


def _newer(target: Path, sources: list[Path]) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(s.stat().st_mtime <= t for s in sources)


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    """Every csrc/*.cu is compiled to an object in parallel, then linked into the shared library."""
    from concurrent.futures import ThreadPoolExecutor

    cus = sorted(CSRC.glob("*.cu"))
    sources = cus + sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "b200_flat.h"]
    if not force and _newer(LIB, sources):
        return LIB
    objdir = ROOT / "c99_vectordb_b200" / "build"
    objdir.mkdir(exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f not in ("-shared",)]
    nvcc = nvcc_path()

    # the scan kernel templates are compiled once per (metric, store) pair so that the build parallelises
    jobs: list[tuple[Path, list[str], str]] = []
    for cu in cus:
        if cu.stem in ("scan_bulk", "scan_ldg"):
            for m in (0, 1):
                for st in (0, 1):
                    jobs.append((cu, [f"-DSCAN_M={m}", f"-DSCAN_S={st}"], f"{cu.stem}_m{m}_s{st}"))
        else:
            jobs.append((cu, [], cu.stem))

    def compile_one(job) -> Path:
        cu, defines, name = job
        obj = objdir / (name + ".o")
        cmd = [nvcc, *compile_flags, *defines, "-c"]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        cmd += ["-o", str(obj), str(cu)]
        subprocess.run(cmd, check=True)
        return obj

    jobs.sort(key=lambda j: 0 if j[0].stem.startswith("scan_") or j[0].stem == "cabi" else 1)  # long jobs first
    with ThreadPoolExecutor(max_workers=max(1, os.cpu_count() or 4)) as pool:
        objs = list(pool.map(compile_one, jobs))
    subprocess.run([nvcc, *NVCC_FLAGS, "-o", str(LIB)] + [str(o) for o in objs], check=True)
    return LIB


if __name__ == "__main__":
    force = "--force" in sys.argv
    verbose = "-v" in sys.argv
    print(build_cuda(force, verbose))
