"""Builds libb200whisper.so (all sm_100a kernels + the C ABI) in-tree with nvcc.

The .so lands next to the sources so that it travels with the repo snapshot to the GPU box; it is
git-ignored.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
SOURCES = ["api.cu", "gemm.cu", "gemm2.cu", "chain.cu", "small.cu", "small_mma.cu", "absorb.cu", "align.cu", "attention.cu", "elementwise.cu",
           "logmel.cu"]
HEADERS = ["common.cuh", "kernels.h", "logmel_core.h", os.path.join(REPO, "include", "b200_whisper.h")]
LIB = os.path.join(HERE, "libb200whisper.so")
STAMP = os.path.join(HERE, ".build_stamp")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-shared",
    "-I" + os.path.join(REPO, "include"), "-I" + HERE,
    "-cudart", "static",
] + os.environ.get("B200W_NVCC_EXTRA", "").split()  # e.g. -DB200W_CROSS_UNROLL=4 for A/B builds (tools/time_cross.py)


def _digest(sources=None) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for f in (SOURCES if sources is None else sources) + HEADERS + [os.path.abspath(__file__)]:
        p = f if os.path.isabs(f) else os.path.join(HERE, f)
        with open(p, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libb200whisper.so")


def build(force: bool = False, verbose: bool = False) -> str:
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest:
        return LIB
    # compile translation units in parallel, then link
    # (an object is reused when its source, the headers and the flags are unchanged: .<name>.stamp beside it)
    objs, procs, stamps = [], [], {}
    for src in SOURCES:
        obj = os.path.join(HERE, src.replace(".cu", ".o"))
        objs.append(obj)
        st = os.path.join(HERE, "." + src + ".stamp")
        dg = _digest([src])
        if not force and not verbose and os.path.exists(obj) and os.path.exists(st) and open(st).read().strip() == dg:
            continue
        stamps[st] = dg
        cmd = [nvcc_path()] + [f for f in NVCC_FLAGS if f != "-shared"] + ["-c", os.path.join(HERE, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    for st, dg in stamps.items():
        with open(st, "w") as f:
            f.write(dg)
    cmd = [nvcc_path(), "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    subprocess.run(cmd, check=True)
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
