"""Host-side logic of the multi-GPU relation view (multike_b200/sharded.py) on CPU:
index arithmetic, and a world_size-2 gloo run in which every rank computes its part of a global
step with the ORACLE, the dense gradients are all-reduced, and the result must equal the
single-process oracle step on the whole batch (same negatives through index_base)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rank_parts_partition_the_global_batch():
    from multike_b200.sharded import local_rows, rank_parts, shard_owner
    from multike_b200.relation_view import clipped_slice, split_batch
    for n1, n2, gb, world in [(700, 500, 200, 2), (463294, 448774, 160000, 8), (10, 1000, 64, 4), (1000, 7, 32, 2)]:
        steps = -(-(n1 + n2) // gb)
        for step in list(range(min(steps, 4))) + [steps - 1]:
            b1, b2 = split_batch(n1, n2, gb)
            a1, e1 = clipped_slice(n1, b1, step)
            a2, e2 = clipped_slice(n2, b2, step)
            want = [("1", i) for i in range(a1, e1)] + [("2", i) for i in range(a2, e2)]
            for by_kg in (False, True):
                got, base_expected = [], 0
                for rank in range(world):
                    (s1, l1), (s2, l2), base = rank_parts(n1, n2, gb, step, rank, world, by_kg=by_kg)
                    assert base == base_expected and l1 >= 0 and l2 >= 0
                    assert not by_kg or (l2 == 0 if rank < world // 2 else l1 == 0)
                    got += [("1", i) for i in range(s1, s1 + l1)] + [("2", i) for i in range(s2, s2 + l2)]
                    base_expected += l1 + l2
                assert got == want
    ids = np.arange(23)
    for world in (2, 4, 8):
        owner, local = shard_owner(ids, world)
        assert np.array_equal(owner * 1 + local * world, ids)
        assert sum(local_rows(23, r, world) for r in range(world)) == 23
        for r in range(world):
            assert local_rows(23, r, world) == int((owner == r).sum())
        # KG-block placement: KG1 = ids [0, 9) on the first half of the ranks, KG2 on the second
        owner, local = shard_owner(ids, world, split=9)
        half = world // 2
        assert (owner[:9] < half).all() and (owner[9:] >= half).all()
        for r in range(world):
            mine = ids[owner == r]
            assert local_rows(23, r, world, split=9) == mine.size
            assert np.array_equal(local[owner == r], np.arange(mine.size))  # local rows are dense, in id order


def test_group_parts_own_ranges_are_the_by_kg_partition():
    """negatives where they live: a rank walks its KG's whole slice, and the positions whose
    positive terms it computes are exactly the slice rank_parts(by_kg=True) gives it"""
    from multike_b200.sharded import group_parts, rank_parts
    for n1, n2, gb, world in [(463294, 448774, 80000, 4), (463294, 448774, 160000, 8), (700, 500, 200, 4), (10, 1000, 64, 8)]:
        steps = -(-(n1 + n2) // gb)
        for step in list(range(min(steps, 3))) + [steps - 1]:
            for rank in range(world):
                kg, (a, ln), (lo, hi), base = group_parts(n1, n2, gb, step, rank, world)
                (a1, l1), (a2, l2), b = rank_parts(n1, n2, gb, step, rank, world, by_kg=True)
                assert kg == (1 if rank < world // 2 else 2) and 0 <= lo <= hi <= ln
                assert (a + lo, hi - lo) == ((a1, l1) if kg == 1 else (a2, l2)) and base + lo == b


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, golden_path, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multike_b200.sharded import rank_parts
    from oracle import device_sampler as ds
    from oracle import relation_view as orv
    g = dict(np.load(golden_path))
    n_ent = int(g["n_ent"])
    t1, t2 = g["triples1"], g["triples2"]
    all1, all2 = np.concatenate([t1, g["sup1"]]), np.concatenate([t2, g["sup2"]])
    kg1 = ds.KG(entity_base=0, n_entities=n_ent, triples=all1)
    kg2 = ds.KG(entity_base=n_ent, n_entities=n_ent, triples=all2)
    gen = torch.Generator().manual_seed(1)
    ent0 = torch.randn(2 * n_ent, 16, generator=gen, dtype=torch.float64) * 0.1
    rel0 = torch.randn(5, 16, generator=gen, dtype=torch.float64) * 0.1
    ent, rel = orv.DenseTable(ent0, True, torch.float64), orv.DenseTable(rel0, True, torch.float64)
    K, per_rank, seed = 5, 60, 9
    for step in range(3):
        (a1, l1), (a2, l2), base = rank_parts(len(t1), len(t2), per_rank * world, step, rank, world,
                                              by_kg=(step % 2 == 1))
        p1, p2 = t1[a1:a1 + l1], t2[a2:a2 + l2]
        skey = ds.stream_key(seed, step)
        neg = []
        for i, (h, r, t) in enumerate(np.concatenate([p1, p2])):
            kg = kg1 if i < l1 else kg2
            neg += ds.sample_one(kg, int(h), int(r), int(t), K, skey, base + i)  # RNG coordinate = global position
        neg = np.asarray(neg).reshape(-1, 3)
        pos = np.concatenate([p1, p2])
        loss, ge, gr = orv.relation_view_step(ent, rel, pos[:, 0], pos[:, 1], pos[:, 2], neg[:, 0], neg[:, 1],
                                              neg[:, 2], 0.001, apply=False)
        # "phase 1" of every rank, then the exchange: dense gradients summed over ranks
        buf = torch.cat([ge.reshape(-1), gr.reshape(-1), torch.tensor([loss], dtype=torch.float64)])
        dist.all_reduce(buf)
        ge = buf[: ge.numel()].reshape(ge.shape)
        gr = buf[ge.numel(): ge.numel() + gr.numel()].reshape(gr.shape)
        from oracle.tf_semantics import adagrad_dense_
        adagrad_dense_(ent.var, ent.acc("r"), ge, 0.001)
        adagrad_dense_(rel.var, rel.acc("r"), gr, 0.001)
    if rank == 0:
        np.savez(out, ent=ent.var.numpy(), rel=rel.var.numpy(), loss=float(buf[-1]))
    dist.destroy_process_group()


def test_two_ranks_equal_one_process_on_the_whole_batch(tmp_path):
    golden_path = os.path.join(ROOT, "tests", "golden", "ref_batch_relation.npz")
    out = str(tmp_path / "two_ranks.npz")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), golden_path, out), nprocs=world, join=True)
    got = dict(np.load(out))
    # single process, global batch
    from oracle import device_sampler as ds
    from oracle import relation_view as orv
    from multike_b200.relation_view import clipped_slice, split_batch
    g = dict(np.load(golden_path))
    n_ent = int(g["n_ent"])
    t1, t2 = g["triples1"], g["triples2"]
    all1, all2 = np.concatenate([t1, g["sup1"]]), np.concatenate([t2, g["sup2"]])
    kg1 = ds.KG(entity_base=0, n_entities=n_ent, triples=all1)
    kg2 = ds.KG(entity_base=n_ent, n_entities=n_ent, triples=all2)
    gen = torch.Generator().manual_seed(1)
    ent0 = torch.randn(2 * n_ent, 16, generator=gen, dtype=torch.float64) * 0.1
    rel0 = torch.randn(5, 16, generator=gen, dtype=torch.float64) * 0.1
    ent, rel = orv.DenseTable(ent0, True, torch.float64), orv.DenseTable(rel0, True, torch.float64)
    K, gb, seed = 5, 120, 9
    for step in range(3):
        b1, b2 = split_batch(len(t1), len(t2), gb)
        a1, e1 = clipped_slice(len(t1), b1, step)
        a2, e2 = clipped_slice(len(t2), b2, step)
        p1, p2 = t1[a1:e1], t2[a2:e2]
        neg = ds.sample_batch(p1, kg1, p2, kg2, K, seed, step)
        pos = np.concatenate([p1, p2])
        loss, _, _ = orv.relation_view_step(ent, rel, pos[:, 0], pos[:, 1], pos[:, 2], neg[:, 0], neg[:, 1], neg[:, 2],
                                            0.001, slot="r")
    assert got["loss"] == pytest.approx(loss, rel=1e-12)
    np.testing.assert_allclose(got["ent"], ent.var.numpy(), rtol=0, atol=1e-13)
    np.testing.assert_allclose(got["rel"], rel.var.numpy(), rtol=0, atol=1e-13)


def _owner_worker(rank, world, port, golden_path, out):
    """world-4 gloo emulation of "negatives where they live" with the ORACLE: every rank walks its
    KG's whole slice (group_parts), keeps the negatives whose corrupted entity it owns (shard_owner,
    KG-block placement) and the positive terms of its own range; dense gradients are all-reduced."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multike_b200.sharded import group_parts, shard_owner
    from oracle import device_sampler as ds
    from oracle import relation_view as orv
    g = dict(np.load(golden_path))
    n_ent = int(g["n_ent"])
    t1, t2 = g["triples1"], g["triples2"]
    all1, all2 = np.concatenate([t1, g["sup1"]]), np.concatenate([t2, g["sup2"]])
    kg1 = ds.KG(entity_base=0, n_entities=n_ent, triples=all1)
    kg2 = ds.KG(entity_base=n_ent, n_entities=n_ent, triples=all2)
    gen = torch.Generator().manual_seed(1)
    ent0 = torch.randn(2 * n_ent, 16, generator=gen, dtype=torch.float64) * 0.1
    rel0 = torch.randn(5, 16, generator=gen, dtype=torch.float64) * 0.1
    ent, rel = orv.DenseTable(ent0, True, torch.float64), orv.DenseTable(rel0, True, torch.float64)
    K, per_rank, seed = 5, 30, 9
    for step in range(3):
        kg_no, (a, ln), (lo, hi), base = group_parts(len(t1), len(t2), per_rank * world, step, rank, world)
        pos = (t1 if kg_no == 1 else t2)[a:a + ln]
        kg = kg1 if kg_no == 1 else kg2
        skey = ds.stream_key(seed, step)
        neg = []
        for i, (h, r, t) in enumerate(pos):
            neg += ds.sample_one(kg, int(h), int(r), int(t), K, skey, base + i)
        neg = np.asarray(neg).reshape(-1, 3)
        pos_of_neg = np.repeat(np.arange(ln), K)
        corrupted = np.where(neg[:, 0] != pos[pos_of_neg, 0], neg[:, 0], neg[:, 2])   # the replaced entity
        same = (neg[:, 0] == pos[pos_of_neg, 0]) & (neg[:, 2] == pos[pos_of_neg, 2])  # negative == positive
        corrupted = np.where(same, neg[:, 0], corrupted)
        owner, _ = shard_owner(corrupted, world, split=n_ent)
        mine = neg[owner == rank]
        own_pos = pos[lo:hi]
        loss, ge, gr = orv.relation_view_step(ent, rel, own_pos[:, 0], own_pos[:, 1], own_pos[:, 2], mine[:, 0],
                                              mine[:, 1], mine[:, 2], 0.001, apply=False)
        buf = torch.cat([ge.reshape(-1), gr.reshape(-1), torch.tensor([loss], dtype=torch.float64)])
        dist.all_reduce(buf)
        ge = buf[: ge.numel()].reshape(ge.shape)
        gr = buf[ge.numel(): ge.numel() + gr.numel()].reshape(gr.shape)
        from oracle.tf_semantics import adagrad_dense_
        adagrad_dense_(ent.var, ent.acc("r"), ge, 0.001)
        adagrad_dense_(rel.var, rel.acc("r"), gr, 0.001)
    if rank == 0:
        np.savez(out, ent=ent.var.numpy(), rel=rel.var.numpy(), loss=float(buf[-1]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [4, 8])
def test_ranks_negatives_where_they_live_equal_one_process(tmp_path, world):
    """world 4: two ranks per KG; world 8: four ranks per KG and, with 240 positives per global step
    over 700 + 500 triples, short and empty tail slices"""
    golden_path = os.path.join(ROOT, "tests", "golden", "ref_batch_relation.npz")
    out = str(tmp_path / "ranks.npz")
    mp.spawn(_owner_worker, args=(world, _free_port(), golden_path, out), nprocs=world, join=True)
    got = dict(np.load(out))
    from oracle import device_sampler as ds
    from oracle import relation_view as orv
    from multike_b200.relation_view import clipped_slice, split_batch
    g = dict(np.load(golden_path))
    n_ent = int(g["n_ent"])
    t1, t2 = g["triples1"], g["triples2"]
    all1, all2 = np.concatenate([t1, g["sup1"]]), np.concatenate([t2, g["sup2"]])
    kg1 = ds.KG(entity_base=0, n_entities=n_ent, triples=all1)
    kg2 = ds.KG(entity_base=n_ent, n_entities=n_ent, triples=all2)
    gen = torch.Generator().manual_seed(1)
    ent0 = torch.randn(2 * n_ent, 16, generator=gen, dtype=torch.float64) * 0.1
    rel0 = torch.randn(5, 16, generator=gen, dtype=torch.float64) * 0.1
    ent, rel = orv.DenseTable(ent0, True, torch.float64), orv.DenseTable(rel0, True, torch.float64)
    K, gb, seed = 5, 30 * world, 9
    for step in range(3):
        b1, b2 = split_batch(len(t1), len(t2), gb)
        a1, e1 = clipped_slice(len(t1), b1, step)
        a2, e2 = clipped_slice(len(t2), b2, step)
        p1, p2 = t1[a1:e1], t2[a2:e2]
        neg = ds.sample_batch(p1, kg1, p2, kg2, K, seed, step)
        pos = np.concatenate([p1, p2])
        loss, _, _ = orv.relation_view_step(ent, rel, pos[:, 0], pos[:, 1], pos[:, 2], neg[:, 0], neg[:, 1], neg[:, 2],
                                            0.001, slot="r")
    assert got["loss"] == pytest.approx(loss, rel=1e-12)
    np.testing.assert_allclose(got["ent"], ent.var.numpy(), rtol=0, atol=1e-13)
    np.testing.assert_allclose(got["rel"], rel.var.numpy(), rtol=0, atol=1e-13)
