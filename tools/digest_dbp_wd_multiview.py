"""Runs ONCE in the build container (PYTHONHASHSEED=0): the reference's OWN host pipeline on
DBP-WD-100K -- base/kgs.py loader, data_model.DataModel (literal cleaning, name / value ids, swapped
attribute triples) and predicate_alignment.PredicateAlignModel (weighted attribute triples, soft
predicate alignment triples) -- imported unmodified through the overlay of multike_b200/refapi
(stand-ins for tensorflow / gensim / Levenshtein), with ONE substitution: the word-vector file the
authors used (wiki-news-300d-1M.vec) is not available, so DataModel._generate_literal_vectors gets a
stand-in that keeps what matters for alignment -- identical literals get identical vectors -- and
costs nothing: literal i gets the unit vector multike_b200.synthetic.literal_vectors([i], dim)
(counter-based Gaussian keyed by the literal's index).  Everything else is the reference's code.

Output: tests/golden/dbp_wd_100k_multiview.npz (integer ids and weights only; vectors are
regenerated from the ids).  The relation-view part is tests/golden/dbp_wd_100k_relation.npz
(tools/digest_dbp_wd.py) from the same loader run, so entity / relation ids agree.
"""
import os
import sys
import tempfile
import zipfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DATASET = os.environ.get("MKE_DATASET", "DBP_WD")   # DBP_YG: BASELINE configs[3], written under oracle/_ref/ (see digest_dbp_wd.py)
OUT = os.path.join(ROOT, "tests", "golden", "dbp_wd_100k_multiview.npz") if DATASET == "DBP_WD" else \
    os.path.join(ROOT, "oracle", "_ref", "%s_100k_multiview.npz" % DATASET.lower())


def main():
    assert os.environ.get("PYTHONHASHSEED") == "0", "run with PYTHONHASHSEED=0 (ids come from set order)"
    refapi = os.path.join(ROOT, "multike_b200", "refapi")
    sys.path[:0] = [refapi, os.path.join(refapi, "_stubs"), os.path.join(REF, "code"), ROOT]
    import data_model
    import predicate_alignment
    from utils import load_args, clear_attribute_triples
    from multike_b200 import synthetic

    folder = os.environ.get("DBP_WD_FOLDER")
    if not folder:
        tmp = tempfile.mkdtemp(prefix="dbpwd_")
        with zipfile.ZipFile(os.path.join(REF, "data", "BootEA_datasets.zip")) as z:
            z.extractall(tmp, [m for m in z.namelist() if "BootEA_%s_100K" % DATASET in m])
        folder = os.path.join(tmp, "BootEA_datasets", "BootEA_%s_100K" % DATASET) + "/"
    args = load_args(os.path.join(REF, "code", "args.json"))
    args.training_data = folder
    args.retrain_literal_embeds = True

    class RecordingMatrix:
        def __init__(self, mat):
            self.mat, self.requests, self.shape = mat, [], mat.shape

        def __getitem__(self, key):
            idx = np.asarray(key[0] if isinstance(key, tuple) else key)
            self.requests.append(idx)
            return self.mat[idx]

    class StandInDataModel(data_model.DataModel):
        def _generate_literal_vectors(self):
            # data_model.py:70-88 up to the word vectors: the same literal list, in the same order
            c1, _, _ = clear_attribute_triples(self.kgs.kg1.local_attribute_triples_list)
            c2, _, _ = clear_attribute_triples(self.kgs.kg2.local_attribute_triples_list)
            value_list = [v for (_, _, v) in c1 + c2]
            local_name_list = list(self.entity_local_name_dict.values())
            self.literal_list = list(set(value_list + local_name_list))
            print('literal num:', len(local_name_list), len(value_list), len(self.literal_list))
            # the two tables the model reads are gathered from this matrix (data_model.py:110, :155):
            # record WHICH literals, so that the fixture can hold ids instead of vectors
            self.literal_vectors_mat = RecordingMatrix(
                synthetic.literal_vectors(np.arange(len(self.literal_list)), self.args.dim))
            self.literal_id_dic = data_model.generate_literal_id_dic(self.literal_list)

    data = StandInDataModel(args)
    kgs = data.kgs
    pam = predicate_alignment.PredicateAlignModel(kgs, args)
    # literal ids behind the two vector tables the model reads (data_model.py:106-114, :150-158)
    uri_of = dict(zip(kgs.kg1.entities_id_dict.values(), kgs.kg1.entities_id_dict.keys()))
    uri_of.update(dict(zip(kgs.kg2.entities_id_dict.values(), kgs.kg2.entities_id_dict.keys())))
    name_literal = np.array([data.literal_id_dic[data.entity_local_name_dict[uri_of[i]]] for i in range(kgs.entities_num)],
                            dtype=np.int32)
    name_req, value_req = data.literal_vectors_mat.requests     # name_ordered_list, value_ordered_list
    assert np.array_equal(name_req, name_literal)
    value_literal = value_req.astype(np.int32)
    assert np.allclose(synthetic.literal_vectors(name_literal[:50], args.dim), data.local_name_vectors[:50], atol=1e-6)
    assert np.allclose(synthetic.literal_vectors(value_literal[:50], args.dim), data.value_vectors[:50], atol=1e-6)

    def arr(lst, weights=False):
        a = np.array(sorted(lst), dtype=np.float64 if weights else np.int64)
        return a.reshape(-1, 4 if weights else 3)

    a1, a2 = arr(pam.attribute_triples_w_weights1, True), arr(pam.attribute_triples_w_weights2, True)
    out = dict(
        attr1=a1[:, :3].astype(np.int32), attr1_w=a1[:, 3].astype(np.float32),
        attr2=a2[:, :3].astype(np.int32), attr2_w=a2[:, 3].astype(np.float32),
        sup_attr1=arr(kgs.kg1.sup_attribute_triples_list).astype(np.int32),
        sup_attr2=arr(kgs.kg2.sup_attribute_triples_list).astype(np.int32),
        name_literal=name_literal, value_literal=value_literal, n_literals=len(data.literal_list),
        attributes_num=kgs.attributes_num, entities_num=kgs.entities_num, relations_num=kgs.relations_num,
    )
    for key in ("sup_relation_alignment_triples1", "sup_relation_alignment_triples2",
                "sup_attribute_alignment_triples1", "sup_attribute_alignment_triples2"):
        a = arr(getattr(pam, key), True)
        out[key] = a[:, :3].astype(np.int32)
        out[key + "_w"] = a[:, 3].astype(np.float32)
    for k, v in out.items():
        print(k, getattr(v, "shape", v))
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
