"""Host-side logic of the staged-rows protocol of multike_b200/sharded_model.py (BASELINE configs[3]) on CPU, world_size 2
under gloo, with the ORACLE doing the arithmetic: every rank holds a shard of the entity table, stages the rows of the
WHOLE batch (peer reads emulated by an all-gather of the shards), computes the full (replicated) attribute-CNN step,
adds the gradient rows of the ids it OWNS to its shard and applies Adagrad there; dense parameters are updated alike on
every rank.  After three steps the shards must equal the single-process oracle run -- duplicated ids in a batch, ids of
both KGs and the batch-wide l2 norm included."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIM, N_ENT, SPLIT, N_ATTR, N_VAL, B, LR = 12, 60, 33, 5, 40, 24, 0.05


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _inputs():
    gen = torch.Generator().manual_seed(5)
    ent0 = torch.randn(N_ENT, DIM, generator=gen, dtype=torch.float64) * 0.1
    att0 = torch.randn(N_ATTR, DIM, generator=gen, dtype=torch.float64) * 0.2
    val = torch.randn(N_VAL, DIM, generator=gen, dtype=torch.float64) * 0.3
    from oracle import attr_cnn as oc
    theta0 = oc.init_theta(DIM, generator=gen)
    rng = np.random.default_rng(2)
    batches = []
    for _ in range(3):
        ih = rng.integers(0, N_ENT, B)
        ih[:4] = ih[4:8]                      # the same entity several times in one batch
        batches.append((ih, rng.integers(0, N_ATTR, B), rng.integers(0, N_VAL, B), rng.choice([1.0, 0.7], B)))
    return ent0, att0, val, theta0, batches


def _step(rows_raw, att, val, theta, ih_local, ia, iv, w):
    """loss and gradients of one attribute step w.r.t. the raw rows handed in, attr table and theta"""
    from oracle import attr_cnn as oc
    from oracle.tf_semantics import l2_normalize
    r = rows_raw.clone().requires_grad_(True)
    a = att.clone().requires_grad_(True)
    t = theta.clone().requires_grad_(True)
    loss = oc.attribute_cnn_loss(l2_normalize(r, 1)[ih_local], a[ia], val[iv], torch.tensor(w), t, DIM)
    return (float(loss.detach()),) + torch.autograd.grad(loss, [r, a, t])


def _worker(rank, world, port, out):
    sys.path[:0] = [ROOT]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multike_b200.sharded import shard_owner
    from oracle.tf_semantics import adagrad_dense_
    ent0, att, val, theta, batches = _inputs()
    owner, local = shard_owner(np.arange(N_ENT), world, split=SPLIT)
    mine = np.arange(N_ENT)[owner == rank]
    shard, shard_acc = ent0[mine].clone(), torch.full((len(mine), DIM), 0.1, dtype=torch.float64)
    att, theta = att.clone(), theta.clone()
    att_acc, th_acc = torch.full_like(att, 0.1), torch.full_like(theta, 0.1)
    losses = []
    for ih, ia, iv, w in batches:
        # stage: rows of the whole batch, one per occurrence, read from their owners' shards
        sizes = [int((owner == r).sum()) for r in range(world)]
        gathered = [torch.zeros(s, DIM, dtype=torch.float64) for s in sizes]
        for r in range(world):              # (shards differ in size: one broadcast per owner instead of an all-gather)
            if r == rank:
                gathered[r] = shard.clone()
            dist.broadcast(gathered[r], r)
        staged = torch.stack([gathered[owner[e]][local[e]] for e in ih])
        dist.barrier()                      # everybody has staged: updates may begin
        loss, g_rows, g_att, g_theta = _step(staged, att, val, theta, np.arange(len(ih)), ia, iv, w)
        # commit: the gradient rows of the ids this rank owns, duplicates summed; phase 2 on the shard
        g_shard = torch.zeros_like(shard)
        for k, e in enumerate(ih):
            if owner[e] == rank:
                g_shard[local[e]] += g_rows[k]
        adagrad_dense_(shard, shard_acc, g_shard, LR)
        adagrad_dense_(att, att_acc, g_att, LR)           # replicated parameters: the same gradient on every rank
        adagrad_dense_(theta, th_acc, g_theta, LR)
        dist.barrier()                      # every shard is updated: the next step may stage
        losses.append(loss)
    np.savez(out % rank, ids=mine, shard=shard.numpy(), att=att.numpy(), theta=theta.numpy(), losses=np.array(losses))
    dist.destroy_process_group()


def test_staged_rows_protocol_equals_one_process(tmp_path):
    world = 2
    out = str(tmp_path / "rank%d.npz")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    from oracle.tf_semantics import adagrad_dense_
    ent, att, val, theta, batches = _inputs()
    ent, att, theta = ent.clone(), att.clone(), theta.clone()
    accs = [torch.full_like(x, 0.1) for x in (ent, att, theta)]
    want_losses = []
    for ih, ia, iv, w in batches:
        loss, g_ent, g_att, g_theta = _step(ent, att, val, theta, ih, ia, iv, w)
        for x, a, g in zip((ent, att, theta), accs, (g_ent, g_att, g_theta)):
            adagrad_dense_(x, a, g, LR)
        want_losses.append(loss)
    seen = np.zeros(N_ENT, bool)
    for rank in range(world):
        got = dict(np.load(out % rank))
        np.testing.assert_allclose(got["losses"], want_losses, rtol=1e-12)
        np.testing.assert_allclose(got["shard"], ent.numpy()[got["ids"]], rtol=0, atol=1e-13)
        np.testing.assert_allclose(got["att"], att.numpy(), rtol=0, atol=1e-13)       # replicas did not drift
        np.testing.assert_allclose(got["theta"], theta.numpy(), rtol=0, atol=1e-13)
        seen[got["ids"]] = True
    assert seen.all() and float(np.abs(ent.numpy() - _inputs()[0].numpy()).max()) > 1e-3
