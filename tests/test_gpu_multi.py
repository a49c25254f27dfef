"""Multi-GPU relation view on real GPUs (needs >= 2 devices on the box; skipped otherwise)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("by_kg", ["1", "0"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_relation_view_equals_single_gpu(world, by_kg):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29510 + world), os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600,
                         env=dict(os.environ, MKE_BY_KG=by_kg))
    assert "MULTI_GPU_CHECK PASS" in out.stdout, out.stdout[-4000:]
