"""In-tree build of libversband_sm100.so (nvcc, sm_100a only).

The shared object is written next to this file so that it travels with the repo snapshot to
the GPU box (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libversband_sm100.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
    "-shared", "-lcudart",
]


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*")) + [HERE.parent / "include" / "versband.h", Path(__file__)]
    return any(p.stat().st_mtime > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    objs = []
    build_dir = HERE / "_build"
    build_dir.mkdir(exist_ok=True)
    procs = []
    for src in sources():
        obj = build_dir / (src.stem + ".o")
        objs.append(obj)
        deps = [src] + list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "versband.h"]
        if not force and obj.exists() and all(obj.stat().st_mtime > d.stat().st_mtime for d in deps):
            continue
        cmd = [NVCC, *[f for f in FLAGS if f not in ("-shared", "-lcudart")], "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{out}")
        if verbose and out:
            print(out)
    cmd = [NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
