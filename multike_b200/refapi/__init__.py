"""Host-side mirror of the reference's operator surface for the hot path.

Modules keep the public names, argument meaning and error behaviour of their counterparts under
/root/reference/code, with torch CUDA tensors / device tables in place of TF tensors:

  losses              losses.py (8 functions, differentiable)
  MultiKE_model       class MultiKE: graphs, trainers, reads, save
  MultiKE_CSL         MultiKE_CV (run_ITC.py)          -> drivers.py
  MultiKE_Late        MultiKE_Late, valid/test[_WVA] (run_SSL.py) -> drivers.py
  literal_encoder     AutoEncoderModel
  base.evaluation / base.alignment / base.batch   valid/test, greedy_alignment, generate_neighbours

To let the reference's own scripts import them under their original top-level names, put this
directory first on ``sys.path`` (then, only where TensorFlow 1.x / gensim are not installed,
``_stubs/``, then the reference's ``code/``); `utils`, `data_model` and `predicate_alignment` stay the
reference's own host-side modules.  ``base/`` deliberately has no ``__init__.py``: as a namespace
package it overlays the reference's ``base/`` (evaluation, alignment, batch from here; kgs, kg, read
from there).  tests/test_reference_overlay.py checks the resolution against the reference tree.
"""
