"""TEST INFRASTRUCTURE (see oracle/__init__.py): byte-compiles the reference's own Python modules from where
they lie under /root/reference/code into oracle/_ref/code/ as sourceless bytecode files (suffix .refbc: the
snapshot tool that ships the tree to the GPU box drops *.pyc) -- the Python
counterpart of compiling a C reference into oracle/_ref/*.so.  No source file is copied; oracle/_ref/ is
git-ignored and travels to the GPU box with the snapshot, where /root/reference does not exist.

What the compiled tree is used for: tests/test_gpu_run_scripts.py executes the UNCHANGED launch scripts
(run_ITC.refbc / run_SSL.refbc: code/run_ITC.py:14-20, code/run_SSL.py:14-20) with multike_b200/refapi first on
sys.path, so that `utils`, `data_model`, `predicate_alignment`, `base.kgs` ... are the reference's own host
code and `MultiKE_CSL`, `MultiKE_Late`, `MultiKE_model`, `losses`, `literal_encoder` are the B200 path.
Only tests may execute anything under oracle/_ref/.
"""
import os
import py_compile
import sys

REF_CODE = "/root/reference/code"
SUFFIX = ".refbc"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "code")


def build(ref_code=REF_CODE, out=OUT):
    """returns the list of compiled files, or None when the reference tree is absent (GPU box)"""
    if not os.path.isdir(ref_code):
        return None
    done = []
    for dirpath, dirnames, filenames in os.walk(ref_code):
        dirnames[:] = [d for d in dirnames if d != "__pycache__"]
        rel = os.path.relpath(dirpath, ref_code)
        for name in sorted(filenames):
            if not name.endswith(".py"):
                continue
            src = os.path.join(dirpath, name)
            dst = os.path.normpath(os.path.join(out, rel, name[:-3] + SUFFIX))  # importable without sources (path hook)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            # dfile: the path shown in tracebacks stays the reference's
            py_compile.compile(src, cfile=dst, dfile=os.path.join("<reference>/code", rel, name), doraise=True,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            done.append(dst)
    with open(os.path.join(out, "PYTHON_VERSION"), "w") as fh:
        fh.write("%d.%d\n" % sys.version_info[:2])
    return done


if __name__ == "__main__":
    files = build()
    print("reference tree absent" if files is None else "compiled %d modules into %s" % (len(files), OUT))
