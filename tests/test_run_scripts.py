"""The north-star's acceptance for the drop-in boundary: the reference's UNCHANGED
run_ITC.py / run_SSL.py execute with multike_b200/refapi in front of the reference's code directory.
This file is the CPU half (build container: the reference sources are there, a GPU is not): the scripts
run through the reference's own host code -- argument file, BootEA folder loader, literal cleaning, word
vectors -- up to the first device call and stop there LOUDLY: there is no CPU fallback.  The GPU half
(tests/test_gpu_run_scripts.py) runs them to 'Training ends'."""
import os

import pytest
import torch

import bootea_fixture as bf
import run_script_util as ru


@pytest.mark.skipif(not os.path.isdir(ru.REF_SRC), reason="the reference tree only exists in the build container")
@pytest.mark.skipif(torch.cuda.is_available(), reason="with a GPU the scripts run to the end (test_gpu_run_scripts.py)")
@pytest.mark.parametrize("script", ["run_ITC", "run_SSL"])
@pytest.mark.parametrize("compiled", [False, True])
def test_unchanged_scripts_reach_the_device_boundary_and_fail_loudly_without_a_gpu(tmp_path, script, compiled):
    if compiled:
        from oracle import build_ref
        assert build_ref.build() is not None
    data = str(tmp_path / "BootEA_TINY") + "/"
    vec = bf.write(data)
    bf.write_args(str(tmp_path / "args.json"), data, str(tmp_path / "out") + "/", vec)
    rc, out = ru.run_script(script, str(tmp_path), data, prefer_compiled=compiled)
    assert "load arguments:" in out                       # utils.load_args (reference)
    assert "read relation triples:" in out                # base/read.py (reference) on the BootEA folder
    assert "supervised relation triples:" in out          # base/kgs.py swapping mode (reference)
    assert "literal num: 480" in out                      # data_model.py:86 (reference), 2 x 240 local names
    assert "refapi/literal_encoder.py" in out             # ... which instantiates OUR LiteralEncoder
    assert rc != 0 and "SCRIPT RETURNED" not in out
    assert "NVIDIA" in out or "CUDA" in out               # the device path refuses to run without a GPU
