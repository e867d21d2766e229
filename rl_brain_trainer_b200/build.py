"""Build ``libkin_b200.so`` in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

    python -m rl_brain_trainer_b200.build [--force] [--verbose]
"""

from __future__ import annotations

import argparse
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
OBJ_DIR = CSRC / "_obj"
LIB = PKG / "libkin_b200.so"

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-I", str(INCLUDE), "-I", str(CSRC)]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; the CUDA path cannot be built (there is no CPU fallback)")


def _stale(target: Path, deps: list[Path]) -> bool:
    return (not target.exists()) or any(d.stat().st_mtime > target.stat().st_mtime for d in deps)


def source_files() -> list[Path]:
    """Every file the library is compiled from (kernels, internal headers, the public C header)."""
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + sorted(INCLUDE.glob("*.h"))


def source_hash() -> str:
    """sha256 over the names and contents of ``source_files()``.  ``build`` bakes it into the library (``kin_source_hash()``),
    ``_lib.lib()`` compares it with the tree it is loaded from: a library built from other sources is rebuilt or refused, so
    "the tests ran the committed source" is checkable, not an mtime convention."""
    h = hashlib.sha256()
    for f in source_files():
        h.update(f.name.encode())
        h.update(b"\0")
        h.update(f.read_bytes())
        h.update(b"\0")
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + sorted(INCLUDE.glob("*.h"))
    OBJ_DIR.mkdir(exist_ok=True)
    nvcc = _nvcc()
    digest = source_hash()
    stamp = OBJ_DIR / "source_hash.txt"      # the hash the object files were compiled from
    if not stamp.exists() or stamp.read_text().strip() != digest:
        force_capi = True                    # kin_capi.o carries the hash: recompile it whenever any source changed
    else:
        force_capi = False
    env = dict(os.environ)
    env.pop("CC", None)  # the image's $CC points at a gcc without the usual spec files; let nvcc pick the host compiler
    env.pop("CXX", None)
    jobs = []
    for src in sources:
        obj = OBJ_DIR / (src.stem + ".o")
        if force or _stale(obj, [src] + headers) or (force_capi and src.stem == "kin_capi"):
            cmd = [nvcc, *NVCC_FLAGS, *( ["-Xptxas", "-v"] if verbose else []), f'-DKIN_SOURCE_HASH="{digest}"', "-c", str(src), "-o", str(obj)]
            jobs.append((src, cmd))

    def run(job):
        src, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        return src, r

    with ThreadPoolExecutor(max_workers=min(8, max(len(jobs), 1))) as pool:
        for src, r in pool.map(run, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src.name}")
    objs = [OBJ_DIR / (s.stem + ".o") for s in sources]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs), "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link of libkin_b200.so failed")
    stamp.write_text(digest + "\n")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
