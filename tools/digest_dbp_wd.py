"""Runs ONCE in the build container: digests DBP-WD-100K (data/BootEA_datasets.zip of the reference)
with the reference's OWN loader (code/base/kgs.py::read_kgs_from_folder, imported unmodified with
stub tensorflow/gensim modules, PYTHONHASHSEED=0) into a compact fixture of integer ids:
local relation triples of both KGs, the swapped "sup" relation triples, and the train/valid/test
links.  The GPU box never sees /root/reference; it reads tests/golden/dbp_wd_100k_relation.npz.
"""
import os
import sys
import tempfile
import types
import zipfile

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# MKE_DATASET=DBP_YG digests DBP-YG-100K (BASELINE configs[3]) instead: written under oracle/_ref/ (git-ignored like the
# compiled reference modules -- produced in the build container by __graft_entry__.build(), travels to the GPU box)
DATASET = os.environ.get("MKE_DATASET", "DBP_WD")
OUT = os.path.join(ROOT, "tests", "golden", "dbp_wd_100k_relation.npz") if DATASET == "DBP_WD" else \
    os.path.join(ROOT, "oracle", "_ref", "%s_100k_relation.npz" % DATASET.lower())


def main():
    assert os.environ.get("PYTHONHASHSEED") == "0", "run with PYTHONHASHSEED=0 (ids come from set order)"
    for name in ("tensorflow", "gensim", "gensim.models", "gensim.models.word2vec"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["gensim.models.word2vec"].Word2Vec = object
    sys.path.insert(0, os.path.join(REF, "code"))
    from base.kgs import read_kgs_from_folder
    tmp = tempfile.mkdtemp(prefix="dbpwd_")
    with zipfile.ZipFile(os.path.join(REF, "data", "BootEA_datasets.zip")) as z:
        members = [m for m in z.namelist() if "BootEA_%s_100K" % DATASET in m]
        z.extractall(tmp, members)
    folder = os.path.join(tmp, "BootEA_datasets", "BootEA_%s_100K" % DATASET) + "/"
    kgs = read_kgs_from_folder(folder, "631/", "swapping", False)
    kg1, kg2 = kgs.kg1, kgs.kg2
    out = dict(
        triples1=np.array(kg1.local_relation_triples_list, dtype=np.int32),
        triples2=np.array(kg2.local_relation_triples_list, dtype=np.int32),
        sup1=np.array(kg1.sup_relation_triples_list, dtype=np.int32),
        sup2=np.array(kg2.sup_relation_triples_list, dtype=np.int32),
        entities1=np.array(sorted(kg1.entities_list), dtype=np.int32),
        entities2=np.array(sorted(kg2.entities_list), dtype=np.int32),
        train_links=np.array(kgs.train_links, dtype=np.int32),
        valid_links=np.array(kgs.valid_links, dtype=np.int32),
        test_links=np.array(kgs.test_links, dtype=np.int32),
        entities_num=kgs.entities_num, relations_num=kgs.relations_num,
    )
    # the filter set aliases relation_triples_set: local + sup triples (base/kg.py:59,134)
    assert len(kg1.local_relation_triples_set) == len(set(map(tuple, out["triples1"])) | set(map(tuple, out["sup1"])))
    for k, v in out.items():
        print(k, getattr(v, "shape", v))
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
