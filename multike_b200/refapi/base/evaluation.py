"""Entry points of code/base/evaluation.py on the fused similarity/rank kernel.

  valid(...)      -> (hits@1, mrr)                    evaluation.py:6-15   (quick mode)
  test(...)       -> (alignment pairs, hits@1, mrr)   evaluation.py:18-28  (accurate mode)
  early_stop(...) -> (flag1, flag2, stop?)            evaluation.py:31-36

Both rankers are one call of greedy_alignment; an optional `mapping` (a dim x dim matrix applied to
the first embedding set, evaluation.py:11) is a small library GEMM on the device.
"""
import torch

from multike_b200.refapi.base.alignment import greedy_alignment


def _ranked(embeds1, embeds2, mapping, top_k, threads_num, metric, normalize, csls_k, accurate):
    if mapping is not None:
        dev = torch.device("cuda")
        embeds1 = torch.as_tensor(embeds1, dtype=torch.float32, device=dev) @ \
            torch.as_tensor(mapping, dtype=torch.float32, device=dev)
    return greedy_alignment(embeds1, embeds2, top_k, threads_num, metric, normalize, csls_k, accurate)


def valid(embeds1, embeds2, mapping, top_k, threads_num, metric='inner', normalize=False, csls_k=0, accurate=False):
    pairs, hits1, mr, mrr = _ranked(embeds1, embeds2, mapping, top_k, threads_num, metric, normalize, csls_k, accurate)
    return hits1, mrr


def test(embeds1, embeds2, mapping, top_k, threads_num, metric='inner', normalize=False, csls_k=0, accurate=True):
    pairs, hits1, mr, mrr = _ranked(embeds1, embeds2, mapping, top_k, threads_num, metric, normalize, csls_k, accurate)
    return pairs, hits1, mrr


def early_stop(flag1, flag2, flag):
    """stop when the metric has not improved twice in a row (flag <= flag2 <= flag1)"""
    stop = flag <= flag2 <= flag1
    if stop:
        print("\n == should early stop == \n")
    return flag2, flag, stop
