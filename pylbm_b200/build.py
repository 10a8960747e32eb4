"""
nvcc wrapper + in-tree cache for the static runtime (liblbm_b200.so) and the
generated per-scheme kernel libraries (liblbmk_<hash>.so).

Plays the role of the reference's CythonCodeWrapper (reference:
pylbm/generator/autowrap.py:52-139: write source, build an extension, import
it) for generator='cuda': write the CUDA source, compile it with nvcc for
sm_100a, load it with ctypes.  Libraries are cached by source hash INSIDE the
package tree (pylbm_b200/_build/) so that what is built on the CPU build box
travels to the GPU box; a missing library is compiled on the spot (nvcc is part
of the image).  There is no CPU fallback: if nvcc or the GPU is missing the
caller gets an exception.
"""

import ctypes
import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BUILD_DIR = os.environ.get("PYLBM_B200_BUILD_DIR", os.path.join(HERE, "_build"))
INCLUDE_DIR = os.path.join(ROOT, "include")
CSRC_DIR = os.path.join(HERE, "csrc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    # ONE CUDA runtime for the static runtime library and every generated kernel library
    # (streams/events created by one are used by launches of the other)
    "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64",
]


class BuildError(RuntimeError):
    pass


def nvcc_path():
    path = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(path):
        raise BuildError("nvcc not found: the CUDA backend cannot be built (there is no CPU fallback)")
    return path


def _run(cmd):
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise BuildError("command failed: %s\n%s" % (" ".join(cmd), proc.stdout))
    return proc.stdout


def _file_hash(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def build_runtime(force=False, verbose=False):
    """compile csrc/lbm_runtime.cu -> _build/liblbm_b200.so (rebuilt when sources change)."""
    os.makedirs(BUILD_DIR, exist_ok=True)
    src = os.path.join(CSRC_DIR, "lbm_runtime.cu")
    deps = [src, os.path.join(INCLUDE_DIR, "lbm_b200.h"), os.path.join(INCLUDE_DIR, "lbmk.h")]
    tag = _file_hash(deps, " ".join(NVCC_FLAGS))
    lib = os.path.join(BUILD_DIR, "liblbm_b200.so")
    stamp = lib + ".hash"
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read().strip() == tag:
        return lib
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-I", INCLUDE_DIR, "-o", lib, src, "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    out = _run(cmd)
    if verbose:
        print(out)
    with open(stamp, "w") as fh:
        fh.write(tag)
    return lib


def kernel_library_path(tag):
    return os.path.join(BUILD_DIR, "liblbmk_%s.so" % tag)


def build_kernels(source, tag, force=False, keep_source=True, extra_flags=()):
    """compile a generated translation unit -> _build/liblbmk_<tag>.so (cached by tag)."""
    os.makedirs(BUILD_DIR, exist_ok=True)
    lib = os.path.join(BUILD_DIR, "liblbmk_%s.so" % tag)
    if os.path.exists(lib) and not force:
        return lib
    cu = os.path.join(BUILD_DIR, "lbmk_%s.cu" % tag)
    with open(cu, "w") as fh:
        fh.write(source)
    tmp = lib + ".tmp%d" % os.getpid()
    _run([nvcc_path()] + NVCC_FLAGS + list(extra_flags) + ["-o", tmp, cu])
    os.replace(tmp, lib)
    if not keep_source:
        os.remove(cu)
    return lib


def load_library(path):
    try:
        return ctypes.CDLL(path)
    except OSError as exc:
        raise BuildError(
            "cannot load %s (%s): the CUDA extension is required, there is no CPU fallback" % (path, exc)
        )
