"""The drop-in boundary of SURVEY.md section 8(b), checked in the build container (the reference tree
is not on the GPU box): with multike_b200/refapi first on sys.path, then the stand-ins for the two
absent third-party packages, then the reference's own code directory, the names run_ITC.py /
run_SSL.py import resolve to this package for the device path and to the reference for host code."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/code"

SCRIPT = r'''
import os, sys
root, ref = sys.argv[1], sys.argv[2]
refapi = os.path.join(root, "multike_b200", "refapi")
sys.path[:0] = [refapi, os.path.join(refapi, "_stubs"), ref, root]
from utils import *                                   # run_ITC.py:3 -- the reference's utils (np, os, json, load_args ...)
assert os.path.samefile(sys.modules["utils"].__file__, os.path.join(ref, "utils.py"))
args = load_args(os.path.join(ref, "args.json"))
assert args.alignment_module == "swapping" and args.dim == 75 and args.neg_triple_num == 10
assert load_session() is not None                     # utils.py:25-28 on the stand-in
from MultiKE_CSL import MultiKE_CV                    # run_ITC.py:6
from MultiKE_Late import MultiKE_Late, valid, test, valid_WVA, test_WVA   # run_SSL.py:6, MultiKE_CSL.py:9
import MultiKE_model, losses, literal_encoder
import base.evaluation, base.alignment, base.batch, base.kgs, base.read, base.kg
mine = [MultiKE_model, losses, literal_encoder, base.evaluation, base.alignment, base.batch,
        sys.modules["MultiKE_CSL"], sys.modules["MultiKE_Late"]]
theirs = [base.kgs, base.read, base.kg]
assert all(os.path.realpath(m.__file__).startswith(os.path.realpath(refapi)) for m in mine)
assert all(os.path.realpath(m.__file__).startswith(os.path.realpath(ref)) for m in theirs)
assert issubclass(MultiKE_CV, MultiKE_model.MultiKE) and issubclass(MultiKE_Late, MultiKE_model.MultiKE)
import data_model                                     # run_ITC.py:4 -- the reference's DataModel on our literal_encoder
assert os.path.samefile(data_model.__file__, os.path.join(ref, "data_model.py"))
assert data_model.LiteralEncoder is literal_encoder.LiteralEncoder
for name in ("relation_logistic_loss", "attribute_logistic_loss", "relation_logistic_loss_wo_negs",
             "attribute_logistic_loss_wo_negs", "logistic_loss_wo_negs", "space_mapping_loss", "orthogonal_loss",
             "alignment_loss"):
    assert callable(getattr(losses, name))
for name in ("train_relation_view_1epo", "train_attribute_view_1epo", "train_cross_kg_entity_inference_relation_view_1epo",
             "train_cross_kg_entity_inference_attribute_view_1epo", "train_cross_kg_relation_inference_1epo",
             "train_cross_kg_attribute_inference_1epo", "train_shared_space_mapping_1epo",
             "train_common_space_learning_1epo", "eval_kg1_useful_ent_embeddings", "eval_kg2_useful_ent_embeddings", "save"):
    assert callable(getattr(MultiKE_model.MultiKE, name)), name
import Levenshtein                                     # stand-in here; predicate_alignment.py:2
assert abs(Levenshtein.ratio("kitten", "sitting") - 8 / 13) < 1e-12 and Levenshtein.ratio("", "") == 1.0
assert Levenshtein.ratio("birthPlace", "birthPlace") == 1.0 and Levenshtein.ratio("abc", "xyz") == 0.0
assert Levenshtein.distance("kitten", "sitting") == 3
import predicate_alignment                            # run_ITC.py:5 -- the reference's PredicateAlignModel
assert os.path.samefile(predicate_alignment.__file__, os.path.join(ref, "predicate_alignment.py"))
assert callable(predicate_alignment.PredicateAlignModel)
print("OVERLAY OK")
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree only exists in the build container")
def test_reference_scripts_resolve_device_modules_here_and_host_modules_there():
    out = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, REF], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True, timeout=300)
    assert "OVERLAY OK" in out.stdout, out.stdout[-3000:]
