"""multike_b200 -- B200-native (sm_100a) training hot path of MultiKE behind a thin C-ABI.

Only what the path needs lives here: ``csrc/`` (CUDA kernels + the C-ABI of
``include/multike_b200.h``), ``_cabi`` (ctypes binding, no fallback), ``tables`` (device tables and
functional wrappers), ``relation_view`` (step / epoch drivers) and ``refapi/`` (the host-side mirror
of the reference's ``losses.py`` / ``MultiKE_model.py`` surface).
"""
__version__ = "0.1.0"
