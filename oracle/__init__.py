"""CPU oracle for the MultiKE training hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``multike_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs use it, and
only as the checker / reported baseline.

Pinning status (see DESIGN.md "Oracle"):
  * sampler / batcher restatement (``oracle/ref_batch.py``): PINNED -- bit-identical to the
    reference's own ``code/base/batch.py`` and ``code/attr_batch.py`` run in the build container
    under fixed ``random``/``numpy`` seeds (fixtures in tests/golden/, generator
    tests/golden/make_golden.py).
  * evaluator / neighbour search (``oracle/alignment.py``): PINNED -- equal to the outputs of the
    reference's own ``code/base/alignment.py`` (greedy_alignment, calculate_rank) and
    ``code/base/batch.py`` (find_neighbours) run in the build container (tests/golden/ref_sim.npz,
    generator tests/golden/make_golden_sim.py); ties rank by ascending column (the reference leaves
    that order to numpy's unstable sorts).
  * device sampler (``oracle/device_sampler.py``): bit-exact restatement of the kernels' counter-based
    sampler; its SEMANTICS (one coin per round, no replacement, filter, exactly K) are checked
    against the pinned reference restatement statistically and case by case.  Likewise ``sample_distinct``:
    bit-exact restatement of mke_sample_distinct (the random.sample batches of MultiKE_model.py:355-358, :377, :399,
    :443, :462 as a keyed permutation); the reference's CPython stream is not reproduced, its semantics (distinct,
    uniform indices) are what tests/test_device_sampler_oracle.py and tests/test_gpu_refapi.py check.
  * reference modules and datasets (``oracle/_ref/``, git-ignored, built by __graft_entry__.build() where
    /root/reference exists): the reference's own code byte-compiled for the unchanged launch scripts
    (oracle/build_ref.py) and the DBP-YG-100K digests of its own loader / DataModel / PredicateAlignModel
    (tools/digest_dbp_wd*.py with MKE_DATASET=DBP_YG) for BASELINE configs[3].
  * TF-1.x arithmetic (losses.py, l2_normalize, Adagrad, conv(), the auto-encoder;
    ``oracle/attr_cnn.py``, ``oracle/autoencoder.py``, ``oracle/tf_semantics.py``,
    ``oracle/relation_view.py``): PARITY UNPINNED -- TensorFlow 1.x is a third-party dependency
    that is neither vendored under /root/reference nor installable here (no wheel for Python
    3.12, no network) and the reference ships no tests or golden vectors.  The restatement follows
    the reference call sites line by line and TF's documented op semantics; its hand-derived
    sparse form is cross-checked against torch autograd through the dense formulation, and the
    conv() restatement against an index-level numpy restatement (tests/test_oracle_attr_cnn.py).
"""
