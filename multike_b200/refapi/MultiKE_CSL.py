"""code/MultiKE_CSL.py: `from MultiKE_CSL import MultiKE_CV` (run_ITC.py:5) resolves here."""
from multike_b200.refapi.drivers import MultiKE_CV, test, valid  # noqa: F401
