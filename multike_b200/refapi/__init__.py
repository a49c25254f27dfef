"""Host-side mirror of the reference's operator surface for the hot path.

``losses`` and ``MultiKE_model`` keep the public names, argument meaning and error behaviour of
/root/reference/code/losses.py and /root/reference/code/MultiKE_model.py (relation-view part), with
torch CUDA tensors / device tables in place of TF tensors.  To let the reference's own scripts
import them under their original top-level names, put this directory first on ``sys.path``.
"""
