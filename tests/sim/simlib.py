"""TEST-ONLY: build and load the kernel sources compiled against cusim.h (a CPU model of a
CUDA grid) and expose them through the package's own call layer with numpy memory.  Used to
check kernel logic against the oracle on the GPU-less build box; never used by the product."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "classpose_b200", "csrc")
SIM_LIB = os.path.join(HERE, "libcpb_sim.so")


def build(force=False):
    # CPB_SIM_LIB: a pre-built variant of the library, e.g. one compiled with -fsanitize=address (then run python with
    # LD_PRELOAD=$(gcc -print-file-name=libasan.so) and ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0)
    if os.environ.get("CPB_SIM_LIB"):
        return os.environ["CPB_SIM_LIB"]
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "cusim.h"),
                                                                 os.path.join(ROOT, "include", "classpose_b200.h")]
    if not force and os.path.exists(SIM_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(SIM_LIB) for d in deps):
        return SIM_LIB
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-DCPB_SIM", "-x", "c++",
           "-I", HERE, "-I", os.path.join(ROOT, "include"), "-I", CSRC, os.path.join(CSRC, "cpb_api.cu"),
           "-o", SIM_LIB, "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("sim build failed:\n" + r.stderr)
    return SIM_LIB


class NumpyMem:
    def empty(self, shape, dtype):
        return np.empty(shape, dtype)

    def zeros(self, shape, dtype):
        return np.zeros(shape, dtype)

    def ptr(self, x):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data

    def keep_alive(self, *a):
        pass


_calls = None


def calls():
    global _calls
    if _calls is None:
        from classpose_b200._abi import declare
        from classpose_b200._calls import Calls
        lib = declare(ctypes.CDLL(build()), cuda=False)
        _calls = Calls(lib, NumpyMem())
    return _calls
