"""Times one training step of each small-batch graph at the shipped sizes (args.json: attribute_batch_size = entity_batch_size
= batch_size = 5000, dim 75; DBP-WD-100K table sizes) with CUDA events: the attribute-view CNN step
(MultiKE_model.py:134-151 -- mke_attr_cnn_fwd_bwd + the three Adagrad applies), the cross-KG positives-only relation step
(:158-170), the ITC common-space step (:225-239) and the SSL space-mapping step (:241-261).  Synthetic tables / indices."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multike_b200 import _cabi, tables as T  # noqa: E402
from multike_b200.attr_view import AttrCNN  # noqa: E402


def timed(fn, reps=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3   # us per step


def main():
    lib = _cabi.load()
    dev, dim, B = "cuda", 75, 5000
    n_ent, n_rel, n_attr, n_val = 200000, 550, 1086, 909911
    gen = torch.Generator().manual_seed(1)
    rng = np.random.default_rng(1)
    ent = T.EmbeddingTable(n_ent, dim, True, dev, init=T.xavier_truncated_normal(n_ent, dim, gen), flags=True, grad_replicas=1)
    rv = T.EmbeddingTable(n_ent, dim, True, dev, init=T.xavier_truncated_normal(n_ent, dim, gen), flags=True, grad_replicas=1)
    fin = T.EmbeddingTable(n_ent, dim, True, dev, init=T.xavier_truncated_normal(n_ent, dim, gen), flags=True, grad_replicas=1)
    rel = T.EmbeddingTable(n_rel, dim, True, dev, init=T.xavier_truncated_normal(n_rel, dim, gen))
    att = T.EmbeddingTable(n_attr, dim, False, dev, init=T.xavier_truncated_normal(n_attr, dim, gen))
    val = T.EmbeddingTable(n_val, dim, False, dev, init=rng.standard_normal((n_val, dim)).astype(np.float32), trainable=False)
    name = T.EmbeddingTable(n_ent, dim, False, dev, init=rng.standard_normal((n_ent, dim)).astype(np.float32), trainable=False)
    cnn = AttrCNN(dim, dev, generator=gen)
    ih = torch.randint(0, n_ent, (B,), dtype=torch.int32, device=dev)
    ia = torch.randint(0, n_attr, (B,), dtype=torch.int32, device=dev)
    iv = torch.randint(0, n_val, (B,), dtype=torch.int32, device=dev)
    w = torch.rand(B, device=dev)
    acc = T.new_loss_accumulator(dev)
    out = {"batch": B, "dim": dim}

    def attr_step():
        cnn.fwd_bwd(ent, att, val, ih, ia, iv, acc, w=w, scale=1.0)
        ent.apply_adagrad("a", 0.001)
        att.apply_adagrad("a", 0.001)
        cnn.apply_adagrad("a", 0.001)

    out["attribute_cnn_step_us"] = timed(attr_step)
    out["attribute_cnn_fwd_bwd_only_us"] = timed(lambda: cnn.fwd_bwd(ent, att, val, ih, ia, iv, acc, w=w, scale=1.0))
    ent.apply_adagrad("a", 0.001); att.apply_adagrad("a", 0.001); cnn.apply_adagrad("a", 0.001)
    pos = torch.stack([torch.randint(0, n_ent, (B,), device=dev), torch.randint(0, n_rel, (B,), device=dev),
                       torch.randint(0, n_ent, (B,), device=dev)], 1).to(torch.int32).contiguous()

    def ckge_step():
        T.rel_step_structured(rv, rel, pos, None, None, 0, acc, w=None, pos_scale=2.0, variant=3)
        T.apply_adagrad_pair(rv, rv.adagrad_slot("c"), 0.001, rel, rel.adagrad_slot("c"), 0.001)

    out["cross_kg_relation_step_us"] = timed(ckge_step)
    pick = torch.randperm(n_ent, device=dev)[:B].to(torch.int32).contiguous()

    def align_step():
        T.align_fwd_bwd(fin, name, rv, ent, pick, acc, name_weight=1.0, scale=1.0)
        for t in (fin, rv, ent):
            t.apply_adagrad("n", 0.004)

    out["itc_common_space_step_us"] = timed(align_step)
    maps = torch.stack([torch.linalg.qr(torch.randn(dim, dim, generator=gen))[0] for _ in range(3)]).to(dev).contiguous()
    maps_g, maps_a = torch.zeros_like(maps), torch.full_like(maps, 0.1)
    ws = torch.empty(int(lib.mke_space_mapping_workspace_floats(B, dim)), dtype=torch.float32, device=dev)
    total = torch.zeros(1, dtype=torch.float64, device=dev)

    def space_step():
        _cabi.check(lib.mke_space_mapping_fwd_bwd(fin.c, name.c, rv.c, ent.c, pick.data_ptr(), B, maps.data_ptr(),
                                                  maps_g.data_ptr(), 2.0, 0.0001, ws.data_ptr(), total.data_ptr(),
                                                  _cabi.current_stream()))
        fin.apply_adagrad("s", 0.001)
        _cabi.check(lib.mke_dense_apply_adagrad(maps.data_ptr(), maps_g.data_ptr(), maps_a.data_ptr(), maps.numel(), 0.001,
                                                _cabi.current_stream()))

    out["ssl_space_mapping_step_us"] = timed(space_step)
    out["randperm_550k_us"] = timed(lambda: torch.randperm(554173, device=dev)[:B])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
