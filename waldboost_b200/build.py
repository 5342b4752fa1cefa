"""Build libwbg.so (the C-ABI CUDA library) in-tree for sm_100a with nvcc.

    python -m waldboost_b200.build [--force] [--verbose]

The shared object is written next to this file so it travels with the source tree; it is git-ignored.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwbg.so")
STAMP = os.path.join(HERE, "csrc", ".build_stamp")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
# per-file extra flags: the pyramid restates float64 expressions whose roundings must not be FMA-contracted
SOURCES = {
    "wbg_api.cu": [],
    "wbg_cascade.cu": [],
    "wbg_pyramid.cu": ["-fmad=false"],
}


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libwbg cannot be built (there is no CPU fallback)")


def _extra_defs():
    """tuning aid: extra nvcc flags (e.g. -DH4_MINB=6) from the environment"""
    return os.environ.get("WBG_NVCC_DEFS", "").split()


def _digest():
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../include/wbg.h"]
    for f in files:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p) and (f.endswith((".cu", ".h", ".cuh"))):
            h.update(f.encode())
            with open(p, "rb") as fh:
                h.update(fh.read())
    h.update(repr((ARCH, COMMON[:4], SOURCES, _extra_defs())).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == digest:
                return LIB
    nvcc = _nvcc()
    objs = []
    for src, extra in SOURCES.items():
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [nvcc] + ARCH + COMMON + extra + _extra_defs() + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
        objs.append(obj)
    cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
