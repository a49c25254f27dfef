"""PyTorch C++ extension over the C-ABI (csrc/torch_ext/multike_torch_ext.cpp): `torch.ops.multike_b200.rel_step`,
`.rows_apply_adagrad`, `.sim_rank` take torch tensors, run on the current CUDA stream and call libmultike_b200.so.
Built in-tree by `python -m multike_b200.torch_ops` (or __graft_entry__.build()); no JIT at run time."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EXT_DIR = os.path.join(HERE, "csrc", "torch_ext")
EXT_LIB = os.path.join(EXT_DIR, "multike_torch_ext.so")
_loaded = False


def build(verbose=False):
    """g++ against the torch headers of this interpreter, linked to libmultike_b200.so next door (rpath $ORIGIN/..)"""
    import subprocess
    import sysconfig
    import torch
    from torch.utils import cpp_extension as ce
    from . import build as mke_build
    mke_build.build()
    src = os.path.join(EXT_DIR, "multike_torch_ext.cpp")
    stamp = EXT_LIB + ".stamp"
    import hashlib
    with open(src, "rb") as fh, open(os.path.join(ROOT, "include", "multike_b200.h"), "rb") as hh:
        digest = hashlib.sha256(fh.read() + hh.read() + torch.__version__.encode()).hexdigest()
    if os.path.exists(EXT_LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return EXT_LIB
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    inc = ce.include_paths() + [os.path.join(cuda_home, "include"), os.path.join(ROOT, "include"), sysconfig.get_paths()["include"]]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI), src, "-o", EXT_LIB + ".tmp"]
    cmd += ["-I" + p for p in inc]
    cmd += ["-L" + libdir, "-L" + os.path.join(HERE, "csrc"), "-lmultike_b200", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch",
            "-Wl,-rpath,$ORIGIN/..", "-Wl,-rpath," + libdir]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout[-4000:])
    if proc.returncode != 0:
        raise RuntimeError("building the torch extension failed")
    os.replace(EXT_LIB + ".tmp", EXT_LIB)
    with open(stamp, "w") as fh:
        fh.write(digest)
    return EXT_LIB


def load():
    """registers torch.ops.multike_b200.*; raises if the extension was not built"""
    global _loaded
    import torch
    if not _loaded:
        if not os.path.exists(EXT_LIB):
            raise RuntimeError("%s is missing: python -m multike_b200.torch_ops" % EXT_LIB)
        torch.ops.load_library(EXT_LIB)
        _loaded = True
    return torch.ops.multike_b200


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
