"""GPU half of the drop-in acceptance: the reference's UNCHANGED
launch scripts run to 'Training ends' on the B200 path.  On the GPU box the reference's own modules come from
oracle/_ref/code (sourceless .pyc compiled from /root/reference/code by oracle/build_ref.py in the build
container; /root/reference itself does not exist there)."""
import glob
import os

import numpy as np
import pytest

import bootea_fixture as bf
import run_script_util as ru

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(ru.reference_code_dir()[0] is None,
                    reason="neither /root/reference/code nor oracle/_ref/code (python oracle/build_ref.py) is present")
@pytest.mark.parametrize("script,cls", [("run_ITC", "MultiKE_CV"), ("run_SSL", "MultiKE_Late")])
def test_unchanged_launch_script_trains_to_the_end(tmp_path, script, cls):
    data = str(tmp_path / "BootEA_TINY") + "/"
    vec = bf.write(data)
    out_dir = str(tmp_path / "out") + "/"
    bf.write_args(str(tmp_path / "args.json"), data, out_dir, vec)
    rc, out = ru.run_script(script, str(tmp_path), data)
    assert rc == 0 and "SCRIPT RETURNED" in out, out[-4000:]
    # the reference's host code ran ...
    assert "load arguments:" in out and "literal num: 480" in out
    # ... its literal cache was written in its own format (data_model.py:26-45)
    lit = np.load(data + "literal_vectors.npy")
    with open(data + "literals.txt", encoding="utf-8") as fh:
        assert len(fh.read().split("\n")) - 1 == lit.shape[0] and lit.shape[1] == 75
    # ... and the device path trained, evaluated and saved
    assert "epoch 1 of literal encoder" in out
    assert "epoch 1 of rel. view" in out or "epoch 1" in out
    assert "Training ends." in out
    assert "results: hits@[1, 5, 10, 50]" in out
    saved = glob.glob(out_dir + "**/ent_embeds.npy", recursive=True)
    assert saved, out[-2000:]
    ent = np.load(saved[0])
    assert ent.shape == (480, 75) and np.isfinite(ent).all()
    np.testing.assert_allclose(np.linalg.norm(ent, axis=1), 1.0, atol=1e-4)   # .eval() of a normalised table
