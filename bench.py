#!/usr/bin/env python
"""bench.py -- positive triples/s through the relation-view TRAIN STEP (BASELINE.json metric).

A "step" is one pass of the hot path over one batch: fused gather -> score -> logistic loss ->
gradient scatter-add with on-device negative sampling (phase 1) + per-touched-row
normalise-backward + Adagrad (phase 2) on the entity and relation tables.  Workload at N=1 is
BASELINE.json configs[1]: DBP-WD-100K-shaped (200 000 entities, 550 relations, 912 068 triples,
synthetic with the measured degree laws), dim=75, batch=20 000, neg=10.

  value      whole-job positives/s with the triple lists resident in HBM (CUDA events)
  e2e        same metric through the public step API with HOST (pinned) positives copied H2D and
             the batch loss read D2H every step
  roofline   phase-1 kernel: algorithmic bytes per launch / mean launch duration (CUDA events
             around one phase-1 launch in --p1-every (default 4) of the timed region: a timed event
             pair costs the step ~4 us, profiles/r1_bench_p1_every*.json) vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  oracle port of the reference CPU path on a bounded sample (rank 0, N=1)
`--impl reference` times that CPU path alone and prints the same line shape.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "pos_triples_per_sec_rel_view_train_step"
UNIT = "triples/s"
WORKLOADS = {
    # BASELINE.json configs[1]
    "dwy100k_rel_d75_b20000_k10": dict(shape="DWY100K", dim=75, batch=20000, neg=10),
    # BASELINE.json configs[4]
    "synth1m_rel_d128_b20000_k25": dict(shape="SYNTH_1M", dim=128, batch=20000, neg=25),
}
DEFAULT_WORKLOAD = "dwy100k_rel_d75_b20000_k10"
FALLBACK_HBM_GBS = 6650.0
P1_KERNEL = ["rel_fused_q8_kernel", "rel_fused_tma_kernel", "rel_fused_ldg_kernel", "rel_fused_q8p_kernel",
             "rel_step_persist_kernel"]
VARIANT_NAME = ["q8_ldg_red", "tma_bulk", "warp_ldg_red", "q8_row_stream", "persistent_step_kernel"]


def bytes_per_positive(dim, K):
    """rows read 3+K, gradient rows written 3+K, 12 B of indices; rows at dim*4 B."""
    return 2 * (3 + K) * dim * 4 + 12


def bytes_per_applied_row(dim):
    """SURVEY.md 8(d), phase 2: read g, v, acc + write v, acc + re-zero g per touched row."""
    return 6 * dim * 4


def ncu_traffic(workload, kernel):
    """dram read+write bytes per launch of `kernel` from the committed ncu --set full capture."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                k = json.load(fh)[workload][kernel]
            return k["dram_read_bytes"] + k["dram_write_bytes"]
        except Exception:
            continue
    return None


def touched_rows_per_step(rv, steps=4):
    """Distinct entity rows one step touches (its positives' endpoints and its negatives), averaged over a
    few steps of the epoch, with the product's own sampler: the row count of phase 2."""
    import torch
    from multike_b200 import tables as T
    tot = 0
    for step in range(steps):
        (a1, b1), (a2, b2) = rv.step_slices(step)
        p1, p2 = rv.triples1[a1:b1], rv.triples2[a2:b2]
        ne, _ = T.sample_structured(p1, rv.kg1, p2, rv.kg2, rv.K, rv.seed, step)
        ids = torch.cat([p1[:, 0], p1[:, 2], p2[:, 0], p2[:, 2], ne.reshape(-1)])
        tot += int(torch.unique(ids).numel())
    return tot / steps


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(name, rank=0, world=1):
    from multike_b200 import synthetic
    w = WORKLOADS[name]
    shape = getattr(synthetic, w["shape"])
    # weak scaling: every rank trains its own DWY100K-shaped pair of KGs (seeded by rank)
    kgs = synthetic.make_kgs(shape, seed=1234 + rank)
    return w, kgs


def cpu_baseline_run(name, steps, warmup):
    from oracle import cpu_path
    w, kgs = make_workload(name)
    r = cpu_path.run_steps(kgs["triples1"], kgs["triples2"], kgs["n_ent"], kgs["n_rel"], kgs["ent_split"], w["dim"],
                           w["batch"], w["neg"], steps=steps, warmup=warmup, workers=4)
    r["value"] = r["positives"] / r["seconds"]
    return r


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; TF-1.x is not installable here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    name = args.workload
    w = WORKLOADS[name]
    steps = max(1, min(args.steps, 60))   # bounded sample: one step is 0.13-0.45 s of all host cores (10-30 s in all)
    warmup = max(1, min(args.warmup, 2))
    r = cpu_baseline_run(name, steps, warmup)
    sample = "%d steps of batch %d (K=%d) of the same workload; sampler in %d forked processes + dense torch-CPU step" % (
        steps, w["batch"], w["neg"], r["sampler_workers"])
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * r["seconds"] / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "dim": w["dim"], "batch": w["batch"], "neg": w["neg"],
                   "note": "oracle port of the TF-1.x CPU graph (dense semantics) + reference sampler restatement"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_sharded(args, rank, world, local_rank):
    """N > 1: ONE workload, entity table row-sharded over the ranks with peer gathers / reductions
    over NVLink fused into phase 1, relation-gradient bucket all-reduced with NCCL.  Weak scaling:
    every rank trains `batch` positives per step (global batch N * batch)."""
    import torch
    import torch.distributed as dist
    from multike_b200 import _cabi
    from multike_b200.sharded import ShardedRelationView

    w, kgs = make_workload(args.workload, 0, 1)  # the same KG pair on every rank
    K, dim, B = w["neg"], w["dim"], w["batch"]
    gen = torch.Generator().manual_seed(20190754)
    from multike_b200 import tables as T
    ent0 = T.xavier_truncated_normal(kgs["n_ent"], dim, gen)
    rel0 = T.xavier_truncated_normal(kgs["n_rel"], dim, gen)
    sv = ShardedRelationView(kgs["n_ent"], kgs["n_rel"], dim, kgs["triples1"], kgs["triples2"], kgs["ent_split"],
                             batch_size=B, neg_num=K, lr=0.001, seed=1234, group=dist.group.WORLD, ent_init=ent0,
                             rel_init=rel0)
    spe = sv.triple_steps
    warmup = max(args.warmup, 3)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    import ctypes
    from multike_b200.sharded_model import PeerBarrier
    lib = _cabi.load()
    # a device-side flag barrier in front of each timed region: the ranks leave dist.barrier() + synchronize() hundreds of
    # microseconds apart, and a rank that starts early would otherwise spend that skew INSIDE its timed region, waiting in
    # the first step's barrier for the GPU of the rank whose host is late
    align = PeerBarrier(dist.group.WORLD)
    sv.train_steps(0, warmup)
    step_no = warmup
    clocks = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        clocks.start()
    _cabi.check(lib.mke_timing_enable(args.steps))  # CUDA events around this rank's phase-1 launches
    _cabi.check(lib.mke_timing_stride(max(1, args.p1_every)))
    launches0 = _cabi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    align.wait()
    e0.record()
    positives = sv.train_steps(step_no % spe, args.steps)
    e1.record()
    barrier()
    step_no += args.steps
    launches = _cabi.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    tot, cnt = ctypes.c_double(0), ctypes.c_int32(0)
    _cabi.check(lib.mke_timing_read(ctypes.byref(tot), ctypes.byref(cnt)))
    p1_ms = tot.value / max(cnt.value, 1)
    _cabi.check(lib.mke_timing_enable(0))
    # end-to-end leg: every step each rank copies its positives in from pinned HOST memory and its share of the
    # step loss back out
    sv.train_steps(step_no % spe, 3, host_fed=True)
    step_no += 3
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    align.wait()
    f0.record()
    e2e_pos = sv.train_steps(step_no % spe, args.steps, host_fed=True)
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    assert float(sv.host_losses[: args.steps].sum()) > 0.0
    walked = sv.positives_walked(step_no % spe, args.steps)  # positives whose ids this rank copies in per call
    stats = torch.tensor([ms, p1_ms, e2e_ms], dtype=torch.float64, device="cuda")
    counts = torch.tensor([positives, launches, e2e_pos, walked], dtype=torch.float64, device="cuda")
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    ms, p1_ms, e2e_ms = [float(x) for x in stats.cpu()]
    positives, launches, e2e_pos, walked = [float(x) for x in counts.cpu()]
    if rank == 0:
        peak, peak_kind = measured_peak()
        alg_bytes = positives / world / args.steps * bytes_per_positive(dim, K)
        achieved = alg_bytes / (p1_ms * 1e-3) / 1e9 if p1_ms > 0 else 0.0
        value = positives / (ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "entities": kgs["n_ent"], "relations": kgs["n_rel"],
                       "triples": int(sv.n1 + sv.n2), "dim": dim, "batch": B, "global_batch": B * world, "neg": K,
                       "steps_per_epoch": spe, "variant": VARIANT_NAME[sv.variant],
                       "l2": "no flush: per-rank working set exceeds L2 / rows come over NVLink",
                       "rank_alignment": "each timed region starts behind a device-side flag barrier of all ranks (after "
                                         "dist.barrier() + synchronize()), so host-side exit skew is not timed",
                       "parallelism": "entity table row-sharded over %d GPUs, KG-block placement (each KG on half of "
                                      "the ranks; %s; peer gathers + peer reductions inside phase 1), relation "
                                      "gradient bucket summed through peer memory, flag barriers, C-side step loop" % (
                                          world, "negatives scored on the rank that owns them, endpoint rows over NVLink"
                                          if sv.owner_negs else "positives trained where their rows live")},
            "clocks": clk,
            "e2e": {"value": e2e_pos / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": walked * 12 / args.steps, "d2h_bytes_per_step": 8 * world,
                    "note": "mke_rel_sharded_train_steps with pinned host triple lists: per step every rank copies the "
                            "positives it walks H2D (all ranks together: h2d_bytes_per_step) and its 8-byte share of the "
                            "step loss D2H"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": P1_KERNEL[sv.variant] + " (phase 1, per rank, rows over NVLink)",
                         "peak_source": peak_kind, "launch_ms": p1_ms, "bytes_per_positive": bytes_per_positive(dim, K)},
        }
        print(json.dumps(line))
    align.close()
    sv.close()
    dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=460)
    ap.add_argument("--warmup", type=int, default=46)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--variant", type=int, default=int(os.environ.get("MKE_VARIANT", "4")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="draw negatives inside the fused kernel")
    ap.add_argument("--cpu-steps", type=int, default=40, help="steps of the CPU baseline sample (0.13-0.45 s each)")
    ap.add_argument("--skip-e2e", action="store_true",
                    help="profiling sessions only: no host-fed leg (ncu serialises streams; the line then carries no e2e)")
    ap.add_argument("--p1-every", type=int, default=4,
                    help="put the CUDA-event pair around one phase-1 launch in N of the timed region")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from multike_b200 import _cabi
    from multike_b200.relation_view import RelationView
    _cabi.load()

    if world > 1:
        return run_sharded(args, rank, world, local_rank)

    w, kgs = make_workload(args.workload, rank, world)
    K, dim, B = w["neg"], w["dim"], w["batch"]
    gen = torch.Generator().manual_seed(20190754 + rank)
    rv = RelationView(kgs["n_ent"], kgs["n_rel"], dim, kgs["triples1"], kgs["triples2"], kgs["ent_split"],
                      batch_size=B, neg_num=K, lr=0.001, seed=1234 + rank, variant=args.variant, generator=gen,
                      pipelined=not args.no_pipeline)
    spe = rv.triple_steps
    warmup = max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg ---------------------------------------------------------------
    import ctypes
    lib = _cabi.load()
    rv.train_steps(0, warmup)
    step_no = warmup
    clocks = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        clocks.start()
    persistent = args.variant == 4 and rv.persist_chunk > 0
    _cabi.check(lib.mke_timing_enable(args.steps))  # CUDA events around the phase-1 (variant 4: all) launches
    _cabi.check(lib.mke_timing_stride(1 if persistent else max(1, args.p1_every)))
    launches0 = _cabi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    positives = rv.train_steps(step_no % spe, args.steps)
    e1.record()
    barrier()
    step_no += args.steps
    launches = _cabi.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    tot, cnt = ctypes.c_double(0), ctypes.c_int32(0)
    _cabi.check(lib.mke_timing_read(ctypes.byref(tot), ctypes.byref(cnt)))
    p1_ms = tot.value / max(cnt.value, 1)
    _cabi.check(lib.mke_timing_enable(0))
    phases = None
    if persistent:  # device-side stamps of the last launch: where the step's time goes inside the kernel
        n_last = args.steps - (args.steps - 1) // rv.persist_chunk * rv.persist_chunk
        st = rv.persist_trace(n_last).cpu().numpy().astype("float64")
        d = st[2:] - st[1:-1]
        phases = {"steps_in_last_launch": int(n_last), "prologue_us": (st[1] - st[0]) * 1e-3,
                  "phase1_us": float(d[0::2].mean()) * 1e-3, "phase2_us": float(d[1::2].mean()) * 1e-3,
                  "note": "%globaltimer at the grid barriers of the last launch (phase incl. its barrier)"}
        touched = touched_rows_per_step(rv)

    # how much of a CUDA-event pair is not the kernel: the same pair around a one-row fill kernel
    probe = torch.zeros(8, device="cuda")
    pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(100)]
    ballast = torch.empty(1 << 28, dtype=torch.float32, device="cuda")
    for _ in range(8):  # ~2 ms of queued device work, so that the pairs below are enqueued AHEAD of the GPU
        ballast.zero_()
    for a, b in pairs:
        a.record()
        _cabi.check(lib.mke_fill_rows(probe.data_ptr(), 1, 8, 8, 0.0, _cabi.current_stream()))
        b.record()
    torch.cuda.synchronize()
    event_floor_ms = sorted(a.elapsed_time(b) for a, b in pairs)[len(pairs) // 2]
    del ballast

    # ---- end-to-end leg: every step copies its positives in from pinned HOST memory and its
    # loss back out (mke_rel_view_t.host_triples / host_step_loss); one sync at the end --------
    e2e_pos, e2e_ms, h2d = 0, 1.0, 0
    if not args.skip_e2e:
        rv.use_host_triples()
        rv.train_steps(step_no % spe, 3, host_fed=True)
        step_no += 3
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        e2e_pos = rv.train_steps(step_no % spe, args.steps, host_fed=True)
        f1.record()
        barrier()
        e2e_ms = f0.elapsed_time(f1)
        h2d = e2e_pos * 12
        e2e_loss = float(rv.host_losses.sum())  # read on the host: the copies have landed
        assert e2e_loss > 0.0

    # ---- reduce over ranks: max time, sum positives -----------------------------------------
    stats = torch.tensor([ms, e2e_ms, p1_ms], dtype=torch.float64, device="cuda")
    counts = torch.tensor([positives, e2e_pos, launches], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    ms, e2e_ms, p1_ms = [float(x) for x in stats.cpu()]
    positives, e2e_pos, launches = [float(x) for x in counts.cpu()]

    if rank == 0:
        peak, peak_kind = measured_peak()
        pos_per_launch = positives / world / args.steps  # per rank per phase-1 launch
        alg_bytes = pos_per_launch * bytes_per_positive(dim, K)
        kernel_note = P1_KERNEL[args.variant] + " (phase 1)"
        if persistent:
            # one launch = `steps per launch` whole steps: phase 1 (bytes_per_positive) + phase 2 (6 dim 4 B per
            # touched entity row and per relation row; the sampler's few MB are not counted)
            p1_bytes = alg_bytes
            p2_bytes = (touched + kgs["n_rel"]) * bytes_per_applied_row(dim)
            steps_per_launch = args.steps / max(cnt.value, 1)
            alg_bytes = (p1_bytes + p2_bytes) * steps_per_launch
            kernel_note = "rel_step_persist_kernel (negatives + phase 1 + phase 2 of %.1f steps per launch)" % steps_per_launch
            phases.update({"phase1_algorithmic_bytes": p1_bytes, "phase2_algorithmic_bytes": p2_bytes,
                           "touched_entity_rows_per_step": touched,
                           "phase1_frac_of_peak": p1_bytes / (phases["phase1_us"] * 1e-6) / 1e9 / peak,
                           "phase2_frac_of_peak": p2_bytes / (phases["phase2_us"] * 1e-6) / 1e9 / peak})
        achieved = alg_bytes / (p1_ms * 1e-3) / 1e9 if p1_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": positives / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "entities": kgs["n_ent"], "relations": kgs["n_rel"],
                       "triples": int(rv.n1 + rv.n2), "dim": dim, "batch": B, "neg": K, "steps_per_epoch": spe,
                       "variant": VARIANT_NAME[args.variant],
                       "l2": "no flush: step working set (var+grad+Adagrad slot of the entity table, %d MB) exceeds "
                             "the 126 MB L2 and each step touches a different random row set"
                             % (3 * kgs["n_ent"] * rv.ent.stride * 4 // 2 ** 20),
                       "parallelism": "1 process per GPU; independent KG pair per rank" if world > 1 else "single GPU"},
            "clocks": clk,
            "e2e": {"value": e2e_pos / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d / world / args.steps,
                    "d2h_bytes_per_step": 8,
                    "note": ("mke_rel_train_steps with pinned host triple lists: per step an H2D copy of the batch "
                             "(two cudaMemcpyAsync + a 4-byte arrival flag on a second stream, consumed by the running "
                             "persistent kernel) and the 8-byte step loss stored to pinned host memory by the kernel")
                    if persistent else
                            ("mke_rel_train_steps with pinned host triple lists: per step an H2D copy of the "
                             "batch (overlapped with the previous step) and an 8-byte D2H loss copy")},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": ncu_traffic(args.workload, P1_KERNEL[args.variant]),
                         "traffic_note": "dram bytes per launch, ncu --set full (cold L2), profiles/r1_traffic.json",
                         "algorithmic_bytes": alg_bytes, "kernel": kernel_note,
                         "peak_source": peak_kind, "launch_ms": p1_ms, "timed_launches": int(cnt.value),
                         "bytes_per_positive": bytes_per_positive(dim, K),
                         "event_pair_floor_ms": event_floor_ms,
                         "event_pair_floor_note": "median of the same CUDA-event pair around a one-row fill kernel: "
                                                  "the part of launch_ms that is launch/event latency, not kernel"},
        }
        if phases is not None:
            line["roofline"]["phases"] = phases
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_baseline_run(args.workload, args.cpu_steps, 1)
            line["cpu_baseline"] = {
                "value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                "sample": "%d steps of batch %d (K=%d): reference sampler restatement in %d forked processes + "
                          "dense-semantics torch-CPU step (oracle port of the TF-1.x graph)" % (
                              args.cpu_steps, B, K, r["sampler_workers"])}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
