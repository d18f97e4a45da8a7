"""In-tree build / load of the sm_100a extension (``distribuuuu_b200/_ext/b200_kernels.so``).

``load()`` JIT-builds with ninja through ``torch.utils.cpp_extension`` into a directory *inside*
the package, so the built ``.so`` travels with a repo snapshot to the GPU box, and is a no-op when
the sources are unchanged.  nvcc flags pin ``-gencode arch=compute_100a,code=sm_100a`` (tcgen05 /
TMA / multimem PTX only assembles for the arch-specific target).
"""
from __future__ import annotations

import hashlib
import importlib.util
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
CSRC = os.path.join(_PKG, "csrc")
BUILD_DIR = os.path.join(_PKG, "_ext")
NAME = "b200_kernels"
SOURCES = ["bindings.cpp", "conv_gemm.cu", "elementwise.cu", "comm.cu", "dwconv.cu", "extras.cu", "attention.cu", "conv3x3_halo.cu", "stem_conv.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xptxas", "-v"]

_lock = threading.Lock()
_module = None


def _source_hash() -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh", ".h", ".cpp")):
            with open(os.path.join(CSRC, name), "rb") as f:
                h.update(name.encode() + b"\0" + f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _import_prebuilt():
    """Import ``_ext/b200_kernels.so`` directly when it was built from the current sources (ninja's mtime
    check is useless after a repo snapshot is copied to another machine)."""
    so = os.path.join(BUILD_DIR, NAME + ".so")
    stamp = os.path.join(BUILD_DIR, "source_hash.txt")
    if not (os.path.exists(so) and os.path.exists(stamp)):
        return None
    with open(stamp) as f:
        if f.read().strip() != _source_hash():
            return None
    import torch  # noqa: F401  (loads libtorch / libc10 the extension links against)
    spec = importlib.util.spec_from_file_location(NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load(verbose: bool = False):
    """Build if needed and import the extension module."""
    global _module
    with _lock:
        if _module is not None:
            return _module
        _module = _import_prebuilt()
        if _module is not None:
            return _module
        os.makedirs(BUILD_DIR, exist_ok=True)
        # cpp_extension adds -gencode flags from TORCH_CUDA_ARCH_LIST / visible devices; pin it so the only
        # real target is the explicit sm_100a one in NVCC_FLAGS (10.0a is accepted by torch's parser).
        os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
        os.environ.setdefault("MAX_JOBS", str(min(8, os.cpu_count() or 4)))
        from torch.utils import cpp_extension
        _module = cpp_extension.load(
            name=NAME,
            sources=[os.path.join(CSRC, s) for s in SOURCES],
            extra_cflags=["-O2", "-std=c++17"],
            extra_cuda_cflags=NVCC_FLAGS,
            extra_include_paths=[CSRC],
            build_directory=BUILD_DIR,
            with_cuda=True,
            verbose=verbose,
        )
        with open(os.path.join(BUILD_DIR, "source_hash.txt"), "w") as f:
            f.write(_source_hash())
        return _module


def is_built() -> bool:
    return os.path.exists(os.path.join(BUILD_DIR, NAME + ".so"))
