"""code/MultiKE_Late.py: `from MultiKE_Late import MultiKE_Late` (run_SSL.py:5) resolves here."""
from multike_b200.refapi.drivers import MultiKE_Late, test, test_WVA, valid, valid_WVA, wva  # noqa: F401
