"""Builds and loads tests/harness/host_harness.cu (CPU loops over the kernels' per-scenario bodies)."""
import ctypes
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libhostharness.so")


def _stale():
    if not os.path.exists(SO):
        return True
    csrc = os.path.join(HERE, "..", "..", "model_predictive_control_b200", "csrc")
    deps = [os.path.join(HERE, "host_harness.cu")] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    return os.path.getmtime(SO) < max(os.path.getmtime(d) for d in deps)


def load():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if _stale():
        if not os.path.exists(nvcc):
            return None
        subprocess.run([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-shared",
                        "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-o", SO,
                        os.path.join(HERE, "host_harness.cu")], check=True, capture_output=True)
    return ctypes.CDLL(SO)
