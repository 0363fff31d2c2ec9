"""Compile csrc/*.cu into image_recommender_b200/libb2k.so (sm_100a only, in-tree)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libb2k.so"
SOURCES = ["api.cu", "pack.cu", "scan.cu", "score_tc.cu", "score_tc2.cu", "score_tn.cu", "select.cu", "xchg.cu", "ingest.cu", "group.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    return cand if Path(cand).exists() else "nvcc"


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + \
        [PKG.parent / "include" / "b2k.h"]
    return any(p.stat().st_mtime > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    obj_dir = PKG / "build"
    obj_dir.mkdir(exist_ok=True)
    nvcc = _nvcc()
    # the default host compiler wrapper of this image lacks a few spec files; use the system g++
    ccbin = ["-ccbin", "/usr/bin/g++"] if Path("/usr/bin/g++").exists() else []
    flags = list(NVCC_FLAGS) + os.environ.get("B2K_NVCC_EXTRA", "").split()

    def compile_one(src: str) -> Path:
        obj = obj_dir / (src[:-3] + ".o")
        cmd = [nvcc, *ccbin, *flags, "-I", str(PKG.parent / "include"), "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    # cudart is linked statically and libcuda is resolved at run time through
    # cudaGetDriverEntryPoint, so the library loads (and exports its symbols) on a GPU-less host.
    cmd = [nvcc, *ccbin, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB), *map(str, objs),
           "-cudart", "static", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
