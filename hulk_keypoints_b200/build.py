"""Build libhulk_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m hulk_keypoints_b200.build [--force] [--verbose]

The shared library lands next to this file so it travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(CSRC, "build")
LIB_PATH = os.path.join(PKG_DIR, "libhulk_sm100.so")
DIAG_LIB_PATH = os.path.join(PKG_DIR, "libhulk_sm100_diag.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libhulk_sm100.so cannot be built")


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers() -> list[str]:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(os.path.dirname(PKG_DIR), "include", "hulk_sm100.h"))
    return hs


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _compile(src: str, verbose: bool, diag: bool = False) -> str:
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + (".diag.o" if diag else ".o"))
    if _stale(obj, [src] + _headers()):
        cmd = [_nvcc(), *NVCC_FLAGS, *(["-DHK_DIAG"] if diag else []), "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            sys.stderr.write(res.stderr)
    return obj


def build(force: bool = False, verbose: bool = False, diag: bool = False) -> str:
    """Compile every csrc/*.cu for sm_100a and link libhulk_sm100.so.  Returns the library path.
    diag=True builds libhulk_sm100_diag.so with -DHK_DIAG instead: the timeline stamps and the HK_TC2_DEBUG switches of the tools/diag_*
    scripts (select it with HK_LIB_PATH).  The shipped library contains neither."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = sources()
    lib_path = DIAG_LIB_PATH if diag else LIB_PATH
    if force:
        for f in os.listdir(OBJ_DIR):
            if f.endswith(".diag.o") == diag:
                os.remove(os.path.join(OBJ_DIR, f))
        if os.path.exists(lib_path):
            os.remove(lib_path)
    if not _stale(lib_path, srcs + _headers()):
        return lib_path
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as pool:
        objs = list(pool.map(lambda s: _compile(s, verbose, diag), srcs))
    # -fvisibility=hidden + extern "C" entry points marked default below via the version script
    cmd = [_nvcc(), "-shared", "-o", lib_path, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC", "-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return lib_path


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--diag", action="store_true", help="build libhulk_sm100_diag.so (-DHK_DIAG) for the tools/diag_* scripts")
    args = ap.parse_args()
    print(build(args.force, args.verbose, args.diag))
