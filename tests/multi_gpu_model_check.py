"""Launched by torchrun (one process per GPU): the full multi-view drivers on ROW-SHARDED entity tables
(multike_b200/sharded_model.py: ShardedMultiKE_Late = run_SSL.py's schedule, ShardedMultiKE_CV = run_ITC.py's) must
print the same per-epoch losses and evaluation lines as the single-GPU drivers (refapi.drivers) on the same
synthetic three-view dataset, and end in the same tables, up to fp32 summation order.  Prints
MULTI_GPU_MODEL_CHECK PASS on rank 0.  MKE_SAME_GPU=1: all ranks share cuda:0 (see tests/multi_gpu_check.py).
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_model_check.py
"""
import contextlib
import io
import os
import random
import re
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

NUM = re.compile(r"-?\d+\.\d+")


def numbers(text):
    """(label, values) of every loss / evaluation line, times stripped"""
    out = []
    for line in text.splitlines():
        if "avg. loss" in line:
            out.append((line.split(",")[0], [float(NUM.findall(line.split("avg. loss:")[1])[0])]))
        elif "results: hits@" in line:
            body = line.split("] = ")[1].split(", time")[0]   # "[h1 h5 h10 h50]%, mr = .., mrr = .."
            out.append(("hits", [float(x) for x in re.findall(r"-?\d+\.?\d*(?:e-?\d+)?", body)]))
    return out


def run(cls, mode, **kw):
    import multiview_fixture as mf
    data, args, pam = mf.make(n=int(os.environ.get("MKE_CHECK_N", "1500")), seed=3)
    args.batch_size, args.attribute_batch_size, args.entity_batch_size = 480, 400, 200
    args.max_epoch, args.shared_learning_max_epoch = 3, 2
    torch.manual_seed(args.seed)
    torch.cuda.manual_seed(args.seed)
    random.seed(args.seed)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        m = cls(data, args, pam, **kw)
        m.run()
        torch.cuda.synchronize()
    return m, buf.getvalue()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    same_gpu = os.environ.get("MKE_SAME_GPU", "0") == "1"
    torch.cuda.set_device(0 if same_gpu else int(os.environ.get("LOCAL_RANK", rank)))
    if same_gpu:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from multike_b200 import sharded_model as SM
    from multike_b200.refapi import drivers as D
    ok, report = True, []
    for mode, sharded_cls, single_cls in (("SSL", SM.ShardedMultiKE_Late, D.MultiKE_Late), ("ITC", SM.ShardedMultiKE_CV, D.MultiKE_CV)):
        sm, s_out = run(sharded_cls, mode, group=dist.group.WORLD)
        rm, r_out = run(single_cls, mode)
        a, b = numbers(s_out), numbers(r_out)
        same_len = len(a) == len(b) and len(a) > 8
        worst_loss = worst_hits = 0.0
        if same_len:
            for (la, va), (lb, vb) in zip(a, b):
                if la != lb or len(va) != len(vb):
                    same_len = False
                    break
                d = max(abs(x - y) for x, y in zip(va, vb))
                if la == "hits":
                    worst_hits = max(worst_hits, d)
                else:
                    worst_loss = max(worst_loss, d / max(abs(vb[0]), 1e-9))
        d_tab = 0.0
        for name in ("ent_embeds", "rv_ent_embeds", "av_ent_embeds"):
            idx = np.arange(0, rm.kgs.entities_num, 7)
            d_tab = max(d_tab, float(np.abs(getattr(sm, name).eval(idx=idx) - getattr(rm, name).eval(idx=idx)).max()))
        d_dense = float(np.abs(sm.rel_embeds.raw() - rm.rel_embeds.raw()).max())
        d_dense = max(d_dense, float((sm._cnns[0].theta - rm._cnns[0].theta).abs().max()))
        good = same_len and worst_loss < 2e-4 and worst_hits <= 0.5 and d_tab < 5e-4 and d_dense < 5e-4
        ok &= good
        report.append("%s lines %d/%d rel.loss %.1e hits %.2f tables %.1e dense %.1e -> %s" % (
            mode, len(a), len(b), worst_loss, worst_hits, d_tab, d_dense, "ok" if good else "MISMATCH"))
        if not good and rank == 0:
            sys.stdout.write("--- sharded ---\n" + "\n".join(l for l in s_out.splitlines() if "loss" in l or "hits" in l)[:3000] + "\n")
            sys.stdout.write("--- single ---\n" + "\n".join(l for l in r_out.splitlines() if "loss" in l or "hits" in l)[:3000] + "\n")
        sm.close()
    flags = torch.tensor([1.0 if ok else 0.0], device="cpu" if same_gpu else "cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    sys.stdout.write("rank %d: %s\n" % (rank, " | ".join(report)))
    sys.stdout.flush()
    dist.barrier()
    if rank == 0:
        sys.stdout.write("MULTI_GPU_MODEL_CHECK %s world %d same_gpu %s\n" % ("PASS" if float(flags) == 1.0 else "FAIL", world, same_gpu))
        sys.stdout.flush()
    dist.destroy_process_group()
    return 0 if float(flags) == 1.0 else 1


if __name__ == "__main__":
    sys.exit(main())
