"""CPU oracle for the MultiKE training hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``multike_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs use it, and
only as the checker / reported baseline.

Pinning status (see DESIGN.md "Oracle"):
  * sampler / batcher restatement (``oracle/ref_batch.py``): PINNED -- bit-identical to the
    reference's own ``code/base/batch.py`` and ``code/attr_batch.py`` run in the build container
    under fixed ``random``/``numpy`` seeds (fixtures in tests/golden/, generator
    tests/golden/make_golden.py).
  * TF-1.x arithmetic (losses.py, l2_normalize, Adagrad; ``oracle/tf_semantics.py``,
    ``oracle/relation_view.py``): PARITY UNPINNED -- TensorFlow 1.x is a third-party dependency
    that is neither vendored under /root/reference nor installable here (no wheel for Python
    3.12, no network) and the reference ships no tests or golden vectors.  The restatement follows
    the reference call sites line by line and TF's documented op semantics; its hand-derived
    sparse form is cross-checked against torch autograd through the dense formulation.
"""
