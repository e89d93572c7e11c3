"""Builds ``libfos_b200.so`` (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  nvcc
cross-compiles without a GPU, so this runs in the CPU-only build container too.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libfos_b200.so"
SOURCES = ["capi.cu", "solver.cu", "matop.cu", "psd.cu", "psd_large.cu", "batch.cu", "direct.cu"]
HEADERS = ["common.cuh", "kernels.cuh", "matvec.cuh", "solver.cuh", "batch.cuh", "../../include/fos_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the fos_b200 CUDA library cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [(CSRC / h).resolve() for h in HEADERS] + [Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    for s in SOURCES:
        obj = CSRC / (Path(s).stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / s), "-o", str(obj)]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {s} ---\n{out}\n")
        elif verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc compilation of the fos_b200 CUDA library failed")
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
            "-Xlinker", "--exclude-libs,ALL", "-o", str(LIB), *objs, "-ldl"]
    subprocess.run(link, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
