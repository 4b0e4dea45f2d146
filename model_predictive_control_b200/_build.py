"""In-tree build of libmpc_b200.so (nvcc, sm_100a only).

``python -m model_predictive_control_b200._build`` or ``__graft_entry__.build()``.
nvcc cross-compiles without a GPU; the built ``.so`` is git-ignored and travels to the GPU
box with the repository snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libmpc_b200.so")
BUILD = os.path.join(PKG, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libmpc_b200 cannot be built")
    return exe


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime() -> float:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    hs.append(os.path.join(os.path.dirname(PKG), "include", "mpc_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src: str, force: bool) -> str:
    obj = os.path.join(BUILD, os.path.basename(src)[:-3] + ".o")
    if (not force and os.path.exists(obj)
            and os.path.getmtime(obj) >= max(os.path.getmtime(src), _headers_mtime())):
        return obj
    cmd = [nvcc(), *NVCC_FLAGS, "-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    with open(obj[:-2] + ".ptxas.log", "w") as fh:
        fh.write(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, res.stderr[-8000:]))
    return obj


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force), srcs))
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(o) for o in objs):
        cmd = [nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + res.stderr[-4000:])
    if verbose:
        print("built", LIB)
    return LIB


def build_variant(name: str, defines, only=("boxqp.cu", "rti.cu")) -> str:
    """Ablation build: the library with extra -D macros (csrc/boxqp_core.cuh: MPC_VAR_*), as lib/variants/<name>.so.
    Sources not in ``only`` reuse the objects of the main build.  Load it with MPC_B200_LIB=<path>."""
    build_library()
    vdir = os.path.join(BUILD, "variants", name)
    os.makedirs(vdir, exist_ok=True)
    os.makedirs(os.path.join(LIBDIR, "variants"), exist_ok=True)
    objs = []

    def one(src):
        base = os.path.basename(src)
        if base not in only:
            return os.path.join(BUILD, base[:-3] + ".o")
        obj = os.path.join(vdir, base[:-3] + ".o")
        res = subprocess.run([nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", src, "-o", obj],
                             capture_output=True, text=True)
        with open(obj[:-2] + ".ptxas.log", "w") as fh:
            fh.write(res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, res.stderr[-8000:]))
        return obj
    with cf.ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(one, sources()))
    out = os.path.join(LIBDIR, "variants", name + ".so")
    res = subprocess.run([nvcc(), "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stderr[-4000:])
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print("built", build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        build_library(force="--force" in sys.argv, verbose=True)
