"""Locate, build and load libclasspose_b200.so.  There is no CPU fallback: if the CUDA
library is missing or cannot be loaded the import of any compute entry point raises."""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

from ._abi import ClassposeB200Error, declare

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libclasspose_b200.so")
SOURCES = ["cpb_api.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]

_lock = threading.Lock()
_lib = None


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "classpose_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc cross-compiles for sm_100a; works on a GPU-less box."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", LIB_PATH,
           *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise ClassposeB200Error("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB_PATH


def load() -> ctypes.CDLL:
    """Load the CUDA library (building it first if nvcc is available and sources are newer)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if _stale():
            try:
                build()
            except (FileNotFoundError, ClassposeB200Error) as e:
                if not os.path.exists(LIB_PATH):
                    raise ClassposeB200Error(
                        f"{LIB_PATH} is missing and could not be built ({e}); "
                        "run `python -c 'import __graft_entry__ as g; g.build()'`. There is no CPU fallback.") from e
        try:
            lib = ctypes.CDLL(LIB_PATH)
        except OSError as e:
            raise ClassposeB200Error(f"cannot load {LIB_PATH}: {e}. There is no CPU fallback.") from e
        _lib = declare(lib, cuda=True)
        return _lib
