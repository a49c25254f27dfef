"""Mirror of code/base/batch.py:119-150 (generate_neighbours / find_neighbours) on the top-k kernel.

The reference returns dict entity -> list of k entity ids (np.argpartition order); here the
lists stay on the device as the int32 [entity_table_rows, k] matrix the on-device sampler reads
(row[0] == -1: no list, fall back to the KG's entity list -- neighbor.get(e, entities_list),
base/batch.py:93-94).  MultiKE.train_relation_view_1epo accepts either form.
"""
import numpy as np
import torch

from multike_b200 import similarity


class NeighbourTable:
    """Device-resident truncated-epsilon candidate lists; dict-like enough for the drivers'
    `len(neighbors1)` print (MultiKE_CSL.py:101)."""

    def __init__(self, matrix, entities):
        self.matrix, self.entities = matrix, entities

    def __len__(self):
        return int(len(self.entities))

    def __bool__(self):
        return len(self) > 0

    def get(self, entity, default=None):
        row = self.matrix[int(entity)]
        return default if int(row[0]) < 0 else row.cpu().tolist()

    def as_dict(self):
        m = self.matrix.cpu().numpy()
        return {int(e): m[int(e)].tolist() for e in np.asarray(self.entities)}


def generate_neighbours(entity_embeds, entity_list, neighbors_num, threads_num, table_rows=None, chunk_rows=8192):
    """entity_embeds[i] is the (normalised) embedding of entity_list[i]; returns the NeighbourTable
    whose row entity_list[i] holds the neighbors_num entities most similar to it."""
    ids = np.ascontiguousarray(entity_list, dtype=np.int32)
    rows = int(table_rows if table_rows is not None else (ids.max() + 1 if ids.size else 0))
    out = torch.full((rows, int(neighbors_num)), -1, dtype=torch.int32, device="cuda")
    similarity.sim_topk(entity_embeds, int(neighbors_num), id_list=ids, out=out, out_rows=ids, normalize=False,
                        chunk_rows=chunk_rows)
    return NeighbourTable(out, ids)
