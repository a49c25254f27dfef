"""Builds the C-ABI shared library (hand-written sm_100a kernels) in-tree with nvcc.

The library is plain CUDA/C++ (no torch headers): ``csrc/libmultike_b200.so`` exports exactly the
symbols declared in ``include/multike_b200.h``.  It is built in-tree so that it travels with the
repository snapshot to the GPU box; nothing is JIT-compiled at run time.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libmultike_b200.so")
STAMP = os.path.join(CSRC, ".build_stamp")
SOURCES = ["mke_rel.cu", "mke_rel_q8.cu", "mke_rel_q8p.cu", "mke_triple.cu", "mke_apply.cu", "mke_sampler.cu", "mke_epoch.cu", "mke_dense.cu", "mke_align.cu", "mke_cnn.cu", "mke_sim.cu", "mke_util.cu"]
HEADERS = ["mke_common.cuh", "mke_rel.cuh", "mke_sampler.cuh", "mke_q8.cuh", os.path.join(ROOT, "include", "multike_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-diag-suppress", "177",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def _digest():
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for f in SOURCES + HEADERS:
        path = f if os.path.isabs(f) else os.path.join(CSRC, f)
        with open(path, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def is_fresh():
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fh:
        return fh.read().strip() == _digest()


def build(force=False, verbose=False):
    """Compile every kernel for sm_100a; returns the path of the shared library."""
    if not force and is_fresh():
        return LIB
    nvcc = _nvcc()
    if nvcc is None:
        if os.path.exists(LIB):  # GPU box without sources changed: use the shipped binary
            return LIB
        raise RuntimeError("nvcc not found and %s is missing" % LIB)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed (%d):\n%s" % (proc.returncode, proc.stdout[-4000:]))
    with open(STAMP, "w") as fh:
        fh.write(_digest())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
