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
SOURCES = ["mke_rel.cu", "mke_rel_q8.cu", "mke_rel_q8p.cu", "mke_rel_persist.cu", "mke_triple.cu", "mke_apply.cu", "mke_sampler.cu", "mke_epoch.cu", "mke_sharded.cu", "mke_gemm.cu", "mke_dense.cu", "mke_align.cu", "mke_space.cu", "mke_cnn.cu", "mke_sim.cu", "mke_sim_tc.cu", "mke_stage.cu", "mke_util.cu"]
HEADERS = ["mke_common.cuh", "mke_rel.cuh", "mke_sampler.cuh", "mke_q8.cuh", "mke_rel_q8p.cuh", "mke_apply.cuh", "mke_rel_persist.cuh", "mke_umma.cuh", os.path.join(ROOT, "include", "multike_b200.h")]
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


def _file_digest(src):
    """digest of one translation unit: flags + the source + every header (headers are few and shared)"""
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for f in [src] + HEADERS:
        path = f if os.path.isabs(f) else os.path.join(CSRC, f)
        with open(path, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every kernel for sm_100a; returns the path of the shared library.

    One nvcc process per translation unit (in parallel, objects cached under csrc/_obj by content
    digest), then one link.  The library is written to a temporary name and renamed, under a file
    lock, so that concurrent ranks of one launch never read a half-written file."""
    if not force and is_fresh():
        return LIB
    nvcc = _nvcc()
    if nvcc is None:
        if os.path.exists(LIB):  # GPU box without a compiler: the shipped binary is all there is
            if os.path.exists(STAMP):
                with open(STAMP) as fh:
                    if fh.read().strip() != _digest():
                        sys.stderr.write("multike_b200.build: WARNING: sources differ from the shipped %s and nvcc "
                                         "is missing; using the shipped binary\n" % LIB)
            return LIB
        raise RuntimeError("nvcc not found and %s is missing" % LIB)
    import fcntl
    from concurrent.futures import ThreadPoolExecutor
    obj_dir = os.path.join(CSRC, "_obj")
    os.makedirs(obj_dir, exist_ok=True)
    with open(os.path.join(obj_dir, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and is_fresh():  # another rank built it while this one waited
            return LIB
        compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]

        def compile_one(src):
            obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
            stamp = obj + ".stamp"
            dig = _file_digest(src)
            if not force and os.path.exists(obj) and os.path.exists(stamp):
                with open(stamp) as fh:
                    if fh.read().strip() == dig:
                        return obj, 0, ""
            cmd = [nvcc] + compile_flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(CSRC, src)]
            proc = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if proc.returncode == 0:
                with open(stamp, "w") as fh:
                    fh.write(dig)
            return obj, proc.returncode, proc.stdout

        with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
            results = list(ex.map(compile_one, SOURCES))
        out = "".join(r[2] for r in results)
        if verbose or any(r[1] for r in results):
            sys.stderr.write(out)
        if any(r[1] for r in results):
            raise RuntimeError("nvcc failed:\n%s" % out[-6000:])
        tmp = LIB + ".tmp.%d" % os.getpid()
        cmd = [nvcc, "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp]
        cmd += [r[0] for r in results]
        proc = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if proc.returncode != 0:
            raise RuntimeError("link failed (%d):\n%s" % (proc.returncode, proc.stdout[-4000:]))
        os.replace(tmp, LIB)
        with open(STAMP, "w") as fh:
            fh.write(_digest())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
