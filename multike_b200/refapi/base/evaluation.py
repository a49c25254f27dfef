"""Mirror of code/base/evaluation.py (valid :6-15, test :18-28, early_stop :31-36)."""
import torch

from multike_b200.refapi.base.alignment import greedy_alignment


def _mapped(embeds1, mapping):
    if mapping is None:
        return embeds1
    a = torch.as_tensor(embeds1, dtype=torch.float32, device="cuda")
    return a @ torch.as_tensor(mapping, dtype=torch.float32, device="cuda")  # [n, d] x [d, d]: library GEMM


def valid(embeds1, embeds2, mapping, top_k, threads_num, metric='inner', normalize=False, csls_k=0, accurate=False):
    _, hits1_12, mr_12, mrr_12 = greedy_alignment(_mapped(embeds1, mapping), embeds2, top_k, threads_num,
                                                  metric, normalize, csls_k, accurate)
    return hits1_12, mrr_12


def test(embeds1, embeds2, mapping, top_k, threads_num, metric='inner', normalize=False, csls_k=0, accurate=True):
    alignment_rest_12, hits1_12, mr_12, mrr_12 = greedy_alignment(_mapped(embeds1, mapping), embeds2, top_k,
                                                                  threads_num, metric, normalize, csls_k, accurate)
    return alignment_rest_12, hits1_12, mrr_12


def early_stop(flag1, flag2, flag):
    if flag <= flag2 <= flag1:
        print("\n == should early stop == \n")
        return flag2, flag, True
    else:
        return flag2, flag, False
