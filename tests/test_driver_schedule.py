"""The epoch schedules of refapi/drivers.py (MultiKE_CV.run = MultiKE_CSL.py:36-108, MultiKE_Late.run =
MultiKE_Late.py:202-290) as sequences of trainer / evaluation calls, on a recording stand-in for the
model: no device involved.  The expected sequences are written out from the reference's loops."""
import types

import multiview_fixture as mv
from multike_b200.refapi import drivers


class Recorder:
    def __init__(self, data, args, pam):
        self.kgs, self.args, self.predicate_align_model = data.kgs, args, pam
        self.early_stop, self.session, self.log = False, None, []
        table = types.SimpleNamespace(rows=2 * 400, export=lambda idx=None: ("rows", len(idx)), eval=lambda session=None: "E")
        self.rv_ent_embeds = self.rel_embeds = self.attr_embeds = table
        for name in ("train_relation_view_1epo", "train_cross_kg_entity_inference_relation_view_1epo",
                     "train_cross_kg_relation_inference_1epo", "train_attribute_view_1epo",
                     "train_cross_kg_entity_inference_attribute_view_1epo", "train_cross_kg_attribute_inference_1epo",
                     "train_common_space_learning_1epo", "train_shared_space_mapping_1epo"):
            setattr(self, name, self._rec(name))

    def _rec(self, name):
        short = {"train_relation_view_1epo": "rel", "train_cross_kg_entity_inference_relation_view_1epo": "ckge_rel",
                 "train_cross_kg_relation_inference_1epo": "ckgp_rel", "train_attribute_view_1epo": "attr",
                 "train_cross_kg_entity_inference_attribute_view_1epo": "ckge_attr",
                 "train_cross_kg_attribute_inference_1epo": "ckga_attr", "train_common_space_learning_1epo": "common",
                 "train_shared_space_mapping_1epo": "mapping"}[name]
        return lambda epoch, *a: self.log.append("%s%d" % (short, epoch))

    def save(self):
        self.log.append("save")

    def _end_of_epoch_sync(self):   # hook of the multi-GPU model: not part of the reference's schedule
        pass

    def _due(self, i):
        return drivers._Driver._due(self, i)


def _patched(monkeypatch, model):
    monkeypatch.setattr(drivers, "valid", lambda m, embed_choice='avg', w=(1, 1, 1): m.log.append("valid:" + embed_choice))
    monkeypatch.setattr(drivers, "test", lambda m, embed_choice='avg', w=(1, 1, 1): m.log.append("test:" + embed_choice))
    monkeypatch.setattr(drivers, "valid_WVA", lambda m: m.log.append("valid:wva"))
    monkeypatch.setattr(drivers, "test_WVA", lambda m: m.log.append("test:wva"))
    monkeypatch.setattr(drivers.bat, "generate_neighbours",
                        lambda emb, ents, k, threads, table_rows=None: model.log.append("nb:%d" % k) or [0] * len(ents))


def _epoch(i, soft, itc):
    seq = ["rel%d" % i, "ckge_rel%d" % i] + (["ckgp_rel%d" % i] if soft else [])
    seq += ["attr%d" % i, "ckge_attr%d" % i] + (["ckga_attr%d" % i] if soft else [])
    return seq + (["common%d" % i] if itc else [])


def test_itc_schedule(monkeypatch, capsys):
    data, args, pam = mv.make()   # max_epoch 4, start_valid 2, eval_freq 2, soft alignment after epoch 1, truncated_freq 2
    m = Recorder(data, args, pam)
    _patched(monkeypatch, m)
    drivers.MultiKE_CV.run(m)
    want = ["test:nv"] + _epoch(1, False, True)
    want += _epoch(2, True, True) + ["valid:rv", "valid:av", "valid:final", "nb:39", "nb:39"]
    want += _epoch(3, True, True)
    want += _epoch(4, True, True) + ["valid:rv", "valid:av", "valid:final"]          # i == max_epoch: break
    want += ["save", "test:nv", "test:rv", "test:av", "test:final"]
    assert m.log == want
    assert pam.updates == []                         # predicate refresh: epochs that are multiples of 10
    # step counts handed to the trainers (MultiKE_CSL.py:38-41)
    plan = drivers._Schedule(m)
    kg1, kg2 = data.kgs.kg1, data.kgs.kg2
    assert plan.relation_steps == -(-(kg1.local_relation_triples_num + kg2.local_relation_triples_num) // args.batch_size)
    assert plan.attribute_steps == -(-(kg1.local_attribute_triples_num + kg2.local_attribute_triples_num) // args.batch_size)
    assert plan.ckge_relation == kg1.sup_relation_triples_list + kg2.sup_relation_triples_list
    assert plan.entity_list == kg1.entities_list + kg2.entities_list


def test_itc_schedule_refreshes_predicates_every_tenth_epoch(monkeypatch, capsys):
    data, args, pam = mv.make()
    args.max_epoch, args.start_valid, args.start_predicate_soft_alignment, args.neg_sampling = 21, 100, 10, "uniform"
    m = Recorder(data, args, pam)
    _patched(monkeypatch, m)
    drivers.MultiKE_CV.run(m)
    assert [p for p, _ in pam.updates] == ["relation", "attribute"] * 2          # epochs 10 and 20
    assert "ckgp_rel10" not in m.log and "ckgp_rel11" in m.log                     # soft alignment: i > 10
    assert not any(x.startswith("nb:") or x.startswith("valid:") for x in m.log)


def test_ssl_schedule(monkeypatch, capsys):
    data, args, pam = mv.make()   # shared_learning_max_epoch 3
    m = Recorder(data, args, pam)
    _patched(monkeypatch, m)
    drivers.MultiKE_Late.run(m)
    want = ["valid:nv", "valid:avg"] + _epoch(1, False, False)
    want += _epoch(2, True, False) + ["valid:rv", "valid:av", "valid:avg", "valid:wva", "nb:39", "nb:39"]
    want += _epoch(3, True, False)
    want += _epoch(4, True, False) + ["valid:rv", "valid:av", "valid:avg", "valid:wva"]
    want += ["mapping1", "mapping2", "valid:final", "mapping3"]
    want += ["save", "test:nv", "test:rv", "test:av", "test:avg", "test:wva", "test:final"]
    assert m.log == want
    assert [p for p, _ in pam.updates] == ["relation", "attribute"] * 2          # at the evaluations of epochs 2 and 4


def test_wva_weights_match_the_reference_formula(capsys):
    """drivers.wva == MultiKE_Late.py:62-86: per view, the mean over entities of the cosine between the
    view's row and the mean of the three views (the diagonal of normalize(v) @ normalize(mean).T)"""
    import numpy as np
    import torch
    rng = np.random.default_rng(2)
    views = [rng.standard_normal((50, 12)) for _ in range(3)]

    def compute_weight(e1, e2, e3):
        other = (e1 + e2 + e3) / 3
        other = other / np.linalg.norm(other, axis=1, keepdims=True)
        e1 = e1 / np.linalg.norm(e1, axis=1, keepdims=True)
        return float(np.mean(np.diag(e1 @ other.T)))

    want = (compute_weight(views[0], views[1], views[2]), compute_weight(views[1], views[0], views[2]),
            compute_weight(views[2], views[0], views[1]))
    got = drivers.wva(*[torch.from_numpy(v) for v in views])
    assert got == __import__("pytest").approx(want, rel=1e-12)
    # view_rows: 'avg' is the weighted sum of the three exported views, anything unknown is the final view
    tab = lambda v: types.SimpleNamespace(export=lambda idx: torch.from_numpy(v[idx]))
    model = types.SimpleNamespace(name_embeds=tab(views[0]), rv_ent_embeds=tab(views[1]), av_ent_embeds=tab(views[2]),
                                  ent_embeds=tab(views[0] * 2))
    idx = [3, 1, 4]
    avg = drivers.view_rows(model, "avg", idx, w=(0.5, 0.25, 2.0))
    assert np.allclose(avg.numpy(), 0.5 * views[0][idx] + 0.25 * views[1][idx] + 2.0 * views[2][idx])
    assert np.allclose(drivers.view_rows(model, "rv", idx).numpy(), views[1][idx])
    assert np.allclose(drivers.view_rows(model, "final", idx).numpy(), 2 * views[0][idx])
