"""The reference's CPU path for the relation-view step, reconstituted from the oracle pieces and
timed as the reported CPU baseline (TEST/BENCH INFRASTRUCTURE -- never imported by the product).

  sampler   oracle/ref_batch.py (bit-identical restatement of code/base/batch.py:33-116), run in
            `batch_threads_num` forked worker processes feeding a queue, exactly like
            MultiKE_model.py:295-301 does;
  step      oracle/relation_view.py: the dense-semantics torch-CPU fp32 restatement of the TF
            graph (l2_normalize of the whole table -> 6 gathers -> losses.py:4-12 -> autograd ->
            dense Adagrad on every row), all host threads.
TensorFlow 1.x itself cannot be installed in this image; this is "kind": "port".
"""
import multiprocessing as mp
import os
import time

import numpy as np
import torch

from . import ref_batch
from . import relation_view as orv

_G = {}


def _worker_batch(step):
    g = _G
    import random
    random.seed(1000 + step)
    np.random.seed(1000 + step)
    pos, neg = ref_batch.relation_triple_batch(g["l1"], g["l2"], g["s1"], g["s2"], g["e1"], g["e2"], g["B"], step,
                                               None, None, g["K"])
    return np.asarray(pos, dtype=np.int64).reshape(-1, 3), np.asarray(neg, dtype=np.int64).reshape(-1, 3)


def run_steps(triples1, triples2, n_ent, n_rel, ent_split, dim, batch_size, neg_num, steps, warmup=1, lr=0.001,
              workers=4, seed=0):
    """Times `steps` reference-style training steps (sampler processes + dense step, pipelined).
    Returns dict(seconds, positives, steps, cores, sampler_workers)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    _G.update(l1=[tuple(int(x) for x in r) for r in triples1], l2=[tuple(int(x) for x in r) for r in triples2],
              e1=list(range(0, ent_split)), e2=list(range(ent_split, n_ent)), B=batch_size, K=neg_num)
    _G["s1"], _G["s2"] = set(_G["l1"]), set(_G["l2"])
    gen = torch.Generator().manual_seed(seed)
    from .tf_semantics import xavier_truncated_normal
    ent = orv.DenseTable(xavier_truncated_normal((n_ent, dim), gen), True, torch.float32)
    rel = orv.DenseTable(xavier_truncated_normal((n_rel, dim), gen), True, torch.float32)
    total = warmup + steps
    workers = max(1, min(workers, cores))
    ctx = mp.get_context("fork")
    positives, t0 = 0, None
    with ctx.Pool(workers) as pool:
        it = pool.imap(_worker_batch, range(total))
        for s in range(total):
            if s == warmup:
                t0 = time.perf_counter()
            pos, neg = next(it)
            orv.relation_view_step(ent, rel, pos[:, 0], pos[:, 1], pos[:, 2], neg[:, 0], neg[:, 1], neg[:, 2], lr)
            if s >= warmup:
                positives += pos.shape[0]
        dt = time.perf_counter() - t0
    return dict(seconds=dt, positives=positives, steps=steps, cores=cores, sampler_workers=workers)
