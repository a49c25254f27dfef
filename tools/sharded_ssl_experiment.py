"""BASELINE configs[3]: DBP-YG-100K (or DBP-WD-100K), SSL mode (run_SSL.py's schedule: MultiKE_Late.run), full
multi-view, entity tables ROW-SHARDED over the GPUs of one box -- and the same run on ONE GPU next to it.

  1 GPU :  python tools/sharded_ssl_experiment.py --dataset dbp_yg --epochs 3 --shared-epochs 2 --out one.json
  N GPUs:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
               tools/sharded_ssl_experiment.py --dataset dbp_yg --epochs 3 --shared-epochs 2 --out n.json
  compare: python tools/sharded_ssl_experiment.py --compare one.json n.json

Both arms run the drivers of multike_b200/refapi/drivers.py (N GPUs: multike_b200/sharded_model.py in front of them)
on the digests of the reference's own loader / DataModel / PredicateAlignModel (tools/digest_dbp_wd*.py; MKE_DATASET=
DBP_YG writes them under oracle/_ref/), the name / value vectors being multike_b200.synthetic.literal_vectors of the
recorded literal ids (the authors' word-vector file is not available).  Same seeds, same batches,
same negatives: the per-epoch losses and Hits@k of the two arms agree up to fp32 summation order.
"""
import argparse
import contextlib
import io
import json
import os
import random
import re
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NUM = re.compile(r"-?\d+\.\d+")


def digest_paths(dataset):
    if dataset == "dbp_wd":
        d = os.path.join(ROOT, "tests", "golden")
        return os.path.join(d, "dbp_wd_100k_relation.npz"), os.path.join(d, "dbp_wd_100k_multiview.npz")
    d = os.path.join(ROOT, "oracle", "_ref")
    return os.path.join(d, "%s_100k_relation.npz" % dataset), os.path.join(d, "%s_100k_multiview.npz" % dataset)


def tuples(a, weights=None):
    rows = a.tolist()
    if weights is None:
        return [tuple(r) for r in rows]
    return [tuple(r) + (float(w),) for r, w in zip(rows, weights.tolist())]


class PredicateAlign:
    """the fields of predicate_alignment.PredicateAlignModel the drivers read (frozen at their initial state)"""

    def __init__(self, mv):
        self.attribute_triples_w_weights1 = tuples(mv["attr1"], mv["attr1_w"])
        self.attribute_triples_w_weights2 = tuples(mv["attr2"], mv["attr2_w"])
        for key in ("sup_relation_alignment_triples1", "sup_relation_alignment_triples2",
                    "sup_attribute_alignment_triples1", "sup_attribute_alignment_triples2"):
            setattr(self, key, tuples(mv[key], mv[key + "_w"]))

    def update_predicate_alignment(self, embeds, predicate_type='relation', w=0.7):
        pass   # (not reached: start_predicate_soft_alignment lies beyond the epochs run here)


def load(dataset, args_ns):
    from multike_b200 import synthetic
    rel_p, mv_p = digest_paths(dataset)
    g, mv = np.load(rel_p), np.load(mv_p)
    ents1, ents2 = g["entities1"].tolist(), g["entities2"].tolist()

    def kg(trip, sup, n_attr_triples, sup_attr, ents):
        k = types.SimpleNamespace()
        k.entities_list, k.entities_num = ents, len(ents)
        k.local_relation_triples_list, k.local_relation_triples_num = tuples(trip), len(trip)
        k.sup_relation_triples_list = tuples(sup)
        k.local_relation_triples_set = set(k.local_relation_triples_list) | set(k.sup_relation_triples_list)
        k.local_attribute_triples_num = int(n_attr_triples)
        k.sup_attribute_triples_list = tuples(sup_attr)
        return k

    valid, test = g["valid_links"], g["test_links"]
    kgs = types.SimpleNamespace(
        kg1=kg(g["triples1"], g["sup1"], len(mv["attr1"]), mv["sup_attr1"], ents1),
        kg2=kg(g["triples2"], g["sup2"], len(mv["attr2"]), mv["sup_attr2"], ents2),
        entities_num=int(g["entities_num"]), relations_num=int(g["relations_num"]), attributes_num=int(mv["attributes_num"]),
        useful_entities_list1=ents1, useful_entities_list2=ents2,
        train_links=g["train_links"].tolist(), valid_links=valid.tolist(), test_links=test.tolist(),
        valid_entities1=valid[:, 0].tolist(), valid_entities2=valid[:, 1].tolist(),
        test_entities1=test[:, 0].tolist(), test_entities2=test[:, 1].tolist())
    data = types.SimpleNamespace(kgs=kgs, local_name_vectors=synthetic.literal_vectors(mv["name_literal"], args_ns.dim),
                                 value_vectors=synthetic.literal_vectors(mv["value_literal"], args_ns.dim))
    return data, PredicateAlign(mv)


def parse(text):
    losses, hits = [], []
    label = None
    for line in text.splitlines():
        if line.endswith("results:"):
            label = line.strip()
        elif "avg. loss" in line:
            nums = NUM.findall(line.split("avg. loss:")[1])   # loss, then the epoch part's wall-clock seconds
            losses.append((line.split(",")[0], float(nums[0]), float(nums[1]) if len(nums) > 1 else None))
        elif "results: hits@" in line:
            body = line.split("] = ")[1].split(", time")[0]   # "[h1 h5 h10 h50]%, mr = .., mrr = .."
            hits.append((label, [float(x) for x in re.findall(r"-?\d+\.?\d*(?:e-?\d+)?", body)]))
    return losses, hits


def compare(a_path, b_path):
    a, b = json.load(open(a_path)), json.load(open(b_path))
    assert len(a["losses"]) == len(b["losses"]) and len(a["hits"]) == len(b["hits"])
    worst_l = max(abs(x[1] - y[1]) / max(abs(x[1]), 1e-9) for x, y in zip(a["losses"], b["losses"]))
    worst_h = max(max(abs(p - q) for p, q in zip(x[1][:4], y[1][:4])) for x, y in zip(a["hits"], b["hits"]))
    out = {"gpus": [a["gpus"], b["gpus"]], "loss_lines": len(a["losses"]), "hits_lines": len(a["hits"]),
           "worst_relative_loss_difference": worst_l, "worst_hits_difference_points": worst_h,
           "seconds": [a["seconds"], b["seconds"]], "final": [a["hits"][-1], b["hits"][-1]],
           "pass": bool(worst_l < 1e-3 and worst_h <= 0.5)}
    print(json.dumps(out))
    return 0 if out["pass"] else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset", default="dbp_yg", choices=["dbp_yg", "dbp_wd"])
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--shared-epochs", type=int, default=2)
    ap.add_argument("--batch", type=int, default=5000, help="args.json batch_size (global batch of the relation view)")
    ap.add_argument("--out", default=None)
    ap.add_argument("--compare", nargs=2, default=None)
    a = ap.parse_args()
    if a.compare:
        return compare(*a.compare)
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    # code/args.json, evaluation once after the last view epoch and once after the last mapping epoch
    args = types.SimpleNamespace(
        alignment_module='swapping', output='/tmp/mke_out/', training_data='x/%s/' % a.dataset, dim=75, seed=7,
        learning_rate=0.001, ITC_learning_rate=0.004, batch_size=a.batch, entity_batch_size=5000, attribute_batch_size=5000,
        neg_triple_num=10, neg_sampling='truncated', truncated_epsilon=0.98, truncated_freq=20, batch_threads_num=4,
        test_threads_num=8, max_epoch=a.epochs, shared_learning_max_epoch=a.shared_epochs, start_valid=1,
        eval_freq=max(a.epochs, 1), top_k=[1, 5, 10, 50], orthogonal_weight=2, cv_name_weight=1, cv_weight=1,
        start_predicate_soft_alignment=10 ** 6, is_save=False)
    t0 = time.time()
    data, pam = load(a.dataset, args)
    t_load = time.time() - t0
    torch.manual_seed(args.seed)
    torch.cuda.manual_seed(args.seed)
    random.seed(args.seed)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        from multike_b200.sharded_model import ShardedMultiKE_Late as Model
        kw = {"group": dist.group.WORLD}
    else:
        from multike_b200.refapi.drivers import MultiKE_Late as Model
        kw = {}
    buf = io.StringIO()
    t1 = time.time()
    with contextlib.redirect_stdout(buf):
        m = Model(data, args, pam, **kw)
        m.save = lambda: None
        m.run()
        torch.cuda.synchronize()
    secs = time.time() - t1
    losses, hits = parse(buf.getvalue())
    if rank == 0:
        rec = {"dataset": a.dataset, "gpus": world, "epochs": a.epochs, "shared_epochs": a.shared_epochs, "batch": a.batch,
               "seconds": secs, "load_seconds": t_load, "losses": losses, "hits": hits}
        if a.out:
            os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
            with open(a.out, "w") as fh:
                json.dump(rec, fh)
        print(json.dumps({k: rec[k] for k in ("dataset", "gpus", "seconds")}), "losses", len(losses), "final hits", hits[-1] if hits else None)
        last = {}
        for label, _, secs_part in losses:   # seconds of each part of the LAST epoch (the first one pays the conversions)
            last[label.split(" of ", 1)[1]] = secs_part
        print("seconds per epoch part (last epoch):", json.dumps(last))
    if world > 1:
        m.close()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
