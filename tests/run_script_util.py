"""Runs the reference's UNCHANGED launch scripts (code/run_ITC.py, code/run_SSL.py) the way SURVEY.md section 7
step 2 prescribes: multike_b200/refapi (the device path under the reference's module names) first on
sys.path, then the stand-ins for the absent third-party packages, then the reference's code directory
(its own host code: utils, data_model, predicate_alignment, base.kgs ...), a repo-local args.json in the
working directory (utils.load_args('args.json') is CWD-relative, code/run_ITC.py:15), runpy as __main__.

Two places hold the reference's code: /root/reference/code (build container, sources) and oracle/_ref/code
(sourceless bytecode compiled from there by oracle/build_ref.py; travels to the GPU box)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference/code"
REF_PYC = os.path.join(ROOT, "oracle", "_ref", "code")

LAUNCH = r'''
import importlib.machinery as M, marshal, os, runpy, sys
root, ref, script, training_data = sys.argv[1:5]
refapi = os.path.join(root, "multike_b200", "refapi")
if script.endswith(".refbc"):   # compiled reference tree: bytecode files under a suffix of their own
    real = os.path.realpath(ref)
    def hook(path):
        if os.path.realpath(path).startswith(real):
            return M.FileFinder(path, (M.SourcelessFileLoader, [".refbc"]))
        raise ImportError(path)
    sys.path_hooks.insert(0, hook)
    sys.path_importer_cache.clear()
sys.path[:0] = [refapi, os.path.join(refapi, "_stubs"), ref, root]
sys.argv = [script, "--training_data", training_data]
path = os.path.join(ref, script)
if script.endswith(".refbc"):
    with open(path, "rb") as fh:
        code = marshal.loads(fh.read()[16:])   # 16-byte .pyc header, then the code object of the unchanged script
    exec(code, {"__name__": "__main__", "__file__": path, "__builtins__": __builtins__})
else:
    runpy.run_path(path, run_name="__main__")
print("SCRIPT RETURNED")
'''


def reference_code_dir(prefer_compiled=False):
    """(directory, suffix) of the reference's code, or (None, None)"""
    have_pyc = os.path.exists(os.path.join(REF_PYC, "run_ITC.refbc"))
    if os.path.isdir(REF_SRC) and not (prefer_compiled and have_pyc):
        return REF_SRC, ".py"
    if have_pyc:
        with open(os.path.join(REF_PYC, "PYTHON_VERSION")) as fh:
            if fh.read().strip() == "%d.%d" % sys.version_info[:2]:
                return REF_PYC, ".refbc"
    return None, None


def run_script(name, workdir, training_data, timeout=900, prefer_compiled=False):
    """name: 'run_ITC' or 'run_SSL'; workdir holds args.json.  Returns (returncode, combined output)."""
    ref, suffix = reference_code_dir(prefer_compiled)
    assert ref is not None
    env = dict(os.environ, PYTHONHASHSEED="0")
    out = subprocess.run([sys.executable, "-c", LAUNCH, ROOT, ref, name + suffix, training_data], cwd=workdir, env=env,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    return out.returncode, out.stdout
