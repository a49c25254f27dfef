"""Launched by torchrun (one process per GPU): the row-sharded relation view on G GPUs must equal
a single-GPU run with batch_size G * B (same negatives through index_base) up to fp32 summation
order.  Prints MULTI_GPU_CHECK PASS on rank 0.  Used by tests/test_gpu_multi.py and by hand:
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py
MKE_SAME_GPU=1: all ranks share cuda:0 (a box with ONE GPU): the shards are still separate allocations reached
through CUDA IPC mappings and the flag barriers still cross process boundaries (the two contexts time-slice the
GPU, so it is slow); torch.distributed then runs on gloo, which the data path does not use anyway.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    same_gpu = os.environ.get("MKE_SAME_GPU", "0") == "1"
    torch.cuda.set_device(0 if same_gpu else int(os.environ.get("LOCAL_RANK", rank)))
    if same_gpu:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    dev = "cpu" if same_gpu else "cuda"
    from multike_b200 import synthetic, tables as T
    from multike_b200.relation_view import RelationView
    from multike_b200.sharded import ShardedRelationView

    shape = dict(n_ent=20_000, n_rel=60, n_rel1=35, n_triples1=60_000, n_triples2=50_000)
    kgs = synthetic.make_kgs(shape, seed=5)
    gen = torch.Generator().manual_seed(77)
    dim, K, B, lr, seed = 75, 10, 3000, 0.001, 21
    ent0 = T.xavier_truncated_normal(kgs["n_ent"], dim, gen)
    rel0 = T.xavier_truncated_normal(kgs["n_rel"], dim, gen)
    by_kg = os.environ.get("MKE_BY_KG", "1") == "1"
    sv = ShardedRelationView(kgs["n_ent"], kgs["n_rel"], dim, kgs["triples1"], kgs["triples2"], kgs["ent_split"],
                             batch_size=B, neg_num=K, lr=lr, seed=seed, group=dist.group.WORLD, ent_init=ent0,
                             rel_init=rel0, by_kg=by_kg)
    steps = 5
    sv.loss_acc.zero_()
    # steps 0-2 in one library call, step 3 on its own, step 4 host fed (positives H2D, loss share D2H)
    trained = sv.train_steps(0, 3) + sv.step(3) + sv.train_steps(4, 1, host_fed=True)
    torch.cuda.synchronize()
    host_loss_ok = abs(float(sv.host_losses[0]) - float(sv._step_loss[0])) <= 1e-12 * abs(float(sv._step_loss[0]))
    tot = torch.cat([sv.loss_acc, torch.tensor([float(trained)], dtype=torch.float64, device="cuda")]).to(dev)
    dist.all_reduce(tot)
    # reference: ONE GPU, batch G*B, same seed (every rank computes it for its own comparison)
    rv = RelationView(kgs["n_ent"], kgs["n_rel"], dim, kgs["triples1"], kgs["triples2"], kgs["ent_split"],
                      batch_size=B * world, neg_num=K, lr=lr, seed=seed, ent_init=ent0, rel_init=rel0)
    ref_trained = rv.train_steps(0, steps)
    ref_loss = float(rv.step_losses.sum().item())
    torch.cuda.synchronize()
    ok = True
    mine = sv.ent.raw_local()
    want = rv.ent.raw()[sv.ent.owned_ids()]
    d_ent = float(np.abs(mine - want).max())
    d_rel = float(np.abs(sv.rel.raw() - rv.rel.raw()).max())
    d_exp = float(np.abs(sv.ent.eval(idx=np.arange(0, kgs["n_ent"], 97)) - rv.ent.eval(idx=np.arange(0, kgs["n_ent"], 97))).max())
    moved = float(np.abs(rv.ent.raw() - ent0.numpy()).max())
    ok &= d_ent < 2e-6 and d_rel < 1e-5 and d_exp < 2e-6 and moved > 1e-4
    ok &= int(tot[1]) == ref_trained
    ok &= abs(float(tot[0]) - ref_loss) <= 1e-5 * abs(ref_loss)
    ok &= float(sv.ent.grad.abs().max()) == 0.0 and (sv.ent.touched is None or int(sv.ent.touched.max()) == 0)
    ok &= host_loss_ok
    flags = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    # one write per line: the ranks share a pipe
    sys.stdout.write("rank %d: d_ent %.2e d_rel %.2e d_export %.2e moved %.2e loss %.6f vs %.6f trained %d vs %d\n" % (
        rank, d_ent, d_rel, d_exp, moved, float(tot[0]), ref_loss, int(tot[1]), ref_trained))
    sys.stdout.flush()
    dist.barrier()
    if rank == 0:
        sys.stdout.write("MULTI_GPU_CHECK %s world %d by_kg %s owner_negs %s same_gpu %s\n" % (
            "PASS" if float(flags) == 1.0 else "FAIL", world, by_kg, sv.owner_negs, same_gpu))
        sys.stdout.flush()
    sv.close()
    dist.destroy_process_group()
    return 0 if float(flags) == 1.0 else 1


if __name__ == "__main__":
    sys.exit(main())
