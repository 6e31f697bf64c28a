#!/usr/bin/env python
"""BASELINE.json configs 3-5 measured the way bench.py measures the headline workload (``python bench.py --workload NAME``):

  dcnv2   DCN-v2 (CrossNetMix, 3 cross layers, low rank 32, 4 experts, parallel MLP 429-256-128) at the synthetic Criteo shape,
          batch 4096 per GPU; tables sharded by field over the ranks (examples/ranking/run_criteo.py:70 wiring)
  din     DIN at the synthetic Amazon-Electronics shape: batch 4096, history length 50, 100 k items, 1 k categories, 190 k users,
          D = 8, attention / final MLP [256, 128] (examples/ranking/run_amazon_electronics.py:23-57); tables replicated
  dssm    DSSM two-tower, MovieLens shape scaled to 1M users x 10k items, towers [256, 128, 64], in-batch negatives (ratio 20)
          through MatchTrainer's in-batch branch (examples/matching/run_ml_dssm.py:52-70 wiring); single GPU

One JSON line per run with the fields of bench.py's line: value (device-resident, CUDA events, max over ranks), e2e (host batches
through the trainer's public ``train_one_epoch``), roofline (the workload's dominant tensor-core GEMMs timed alone against the
measured bf16 peak, fp32-equivalent flops), cpu_baseline (the installed reference's own trainer, bench.py --impl reference
--workload NAME).  A "step" is one full training step (zero_grad, forward, loss, backward, optimiser)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
BATCH = 4096
MLP = {"dims": [256, 128], "dropout": 0.2, "activation": "relu"}
DIN_SHAPE = {"items": 100_000, "cates": 1000, "users": 190_000, "L": 50, "D": 8}
DSSM_SHAPE = {"users": 1_000_000, "items": 10_000, "L": 50, "D": 16, "neg": 20, "dims": [256, 128, 64]}
N_POOL = 32


# ---------------------------------------------------------------------------------------------------------------------
# synthetic batches (numpy on the host: shared by the b200 arm and the reference arm)
# ---------------------------------------------------------------------------------------------------------------------
def numpy_batch(name, seed, batch=BATCH):
    import numpy as np
    rng = np.random.default_rng(seed)
    if name == "dcnv2":
        x = {"I%d" % i: rng.random(batch, dtype=np.float32) for i in range(13)}
        x.update({"C%d" % i: rng.integers(0, 1_000_000, batch, dtype=np.int64) for i in range(26)})
    elif name == "din":
        s = DIN_SHAPE
        lens = rng.integers(1, s["L"] + 1, batch)
        keep = np.arange(s["L"])[None, :] < lens[:, None]  # post-padding with 0 (utils/data.py:175-176)
        x = {"target_item_id": rng.integers(1, s["items"] + 1, batch, dtype=np.int64), "target_cate_id": rng.integers(1, s["cates"] + 1, batch, dtype=np.int64),
             "user_id": rng.integers(1, s["users"] + 1, batch, dtype=np.int64),
             "hist_item_id": rng.integers(1, s["items"] + 1, (batch, s["L"]), dtype=np.int64) * keep, "hist_cate_id": rng.integers(1, s["cates"] + 1, (batch, s["L"]), dtype=np.int64) * keep}
    else:
        s = DSSM_SHAPE
        lens = rng.integers(1, s["L"] + 1, batch)
        keep = np.arange(s["L"])[None, :] < lens[:, None]
        x = {"user_id": rng.integers(0, s["users"], batch, dtype=np.int64), "item_id": rng.integers(0, s["items"], batch, dtype=np.int64),
             "hist_item_id": rng.integers(1, s["items"], (batch, s["L"]), dtype=np.int64) * keep}
    y = rng.integers(0, 2, batch).astype(np.float32)
    return x, y


def build_features(name, F):
    """Feature lists of a workload from a ``features`` module (this package's or the reference's: same constructors)."""
    if name == "dcnv2":
        dense = [F.DenseFeature("I%d" % i) for i in range(13)]
        sparse = [F.SparseFeature("C%d" % i, vocab_size=1_000_000, embed_dim=16) for i in range(26)]
        return dense, sparse
    if name == "din":
        s = DIN_SHAPE
        feats = [F.SparseFeature("target_item_id", vocab_size=s["items"] + 1, embed_dim=s["D"]), F.SparseFeature("target_cate_id", vocab_size=s["cates"] + 1, embed_dim=s["D"]),
                 F.SparseFeature("user_id", vocab_size=s["users"] + 1, embed_dim=s["D"])]
        hist = [F.SequenceFeature("hist_item_id", vocab_size=s["items"] + 1, embed_dim=s["D"], pooling="concat", shared_with="target_item_id"),
                F.SequenceFeature("hist_cate_id", vocab_size=s["cates"] + 1, embed_dim=s["D"], pooling="concat", shared_with="target_cate_id")]
        return feats, hist
    s = DSSM_SHAPE
    user = [F.SparseFeature("user_id", vocab_size=s["users"], embed_dim=s["D"]), F.SequenceFeature("hist_item_id", vocab_size=s["items"], embed_dim=s["D"], pooling="mean", shared_with="item_id")]
    item = [F.SparseFeature("item_id", vocab_size=s["items"], embed_dim=s["D"])]
    return user, item


def build_model(name, F, M_rank, M_match):
    if name == "dcnv2":
        dense, sparse = build_features(name, F)
        return M_rank.DCNv2(dense + sparse, n_cross_layers=3, mlp_params=dict(MLP))
    if name == "din":
        feats, hist = build_features(name, F)
        return M_rank.DIN(features=feats, history_features=hist, target_features=feats, mlp_params={"dims": [256, 128]}, attention_mlp_params={"dims": [256, 128]})
    user, item = build_features(name, F)
    return M_match.DSSM(user, item, temperature=0.02, user_params={"dims": list(DSSM_SHAPE["dims"]), "activation": "prelu"}, item_params={"dims": list(DSSM_SHAPE["dims"]), "activation": "prelu"})


DESCRIBE = {
    "dcnv2": "DCN-v2 synthetic Criteo-shape: 13 dense + 26 sparse x 1M vocab x dim 16, batch 4096 per GPU, CrossNetMix (3 layers, rank 32, 4 experts) || MLP 429-256-128 relu dropout 0.2, LR head",
    "din": "DIN synthetic Amazon-Electronics-shape: batch 4096 per GPU, history 50, 100k items / 1k categories / 190k users, dim 8, attention MLP 32-256-128-1 (Dice) x 2, final MLP 64-256-128-1 (Dice)",
    "dssm": "DSSM two-tower synthetic MovieLens-shape: 1M users x 10k items, dim 16, history 50 (mean-pooled, item table shared), towers 256-128-64 prelu, batch 4096, in-batch negatives ratio 20, cross entropy",
}
METRIC = {"dcnv2": "ctr_samples_per_sec_dcnv2_criteo_train_step", "din": "ctr_samples_per_sec_din_amazon_train_step", "dssm": "match_samples_per_sec_dssm_inbatch_train_step"}
# dominant GEMMs of one step: (M, N, K, a_mn, b_mn, split_k) through rh_gemm_tf32x3
GEMMS = {
    "dcnv2": [(4096, 132, 429, 0, 0, 1), (4096, 128, 128, 0, 0, 1), (4096, 429, 128, 0, 0, 1)] * 3 + [(4096, 128, 429, 0, 1, 1), (429, 128, 4096, 1, 1, 16), (4096, 128, 128, 0, 1, 1), (128, 128, 4096, 1, 1, 16), (4096, 429, 132, 0, 1, 1),
                                                                                                       (132, 429, 4096, 1, 1, 16)] * 3,
    "din": [(204800, 256, 32, 0, 0, 1), (204800, 128, 256, 0, 0, 1), (204800, 256, 128, 0, 1, 1), (256, 32, 204800, 1, 1, 64), (128, 256, 204800, 1, 1, 64)] * 2,
    "dssm": [(4096, 4096, 64, 0, 0, 1)],
}


# ---------------------------------------------------------------------------------------------------------------------
# reference arm: the installed reference's own models + trainers on the host cores (called from bench.py --impl reference)
# ---------------------------------------------------------------------------------------------------------------------
def reference_run(name, steps, budget_s, seed=2022):
    import numpy as np
    import torch
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "torch_rechub")):
        return None
    for p_ in [q for q in sys.path if q.rstrip("/").endswith("torch-rechub_b200")]:
        sys.path.remove(p_)
    sys.path.insert(0, ref_dir)
    import torch_rechub
    assert os.path.abspath(torch_rechub.__file__).startswith(os.path.abspath(ref_dir))
    import torch_rechub.basic.features as F
    import torch_rechub.models.matching as MM
    import torch_rechub.models.ranking as MR
    from torch.utils.data import DataLoader
    from torch_rechub.trainers import CTRTrainer, MatchTrainer
    from torch_rechub.utils.data import TorchDataset
    cores = os.cpu_count() or 1
    torch.manual_seed(seed)
    torch.set_num_threads(min(32, cores))
    model = build_model(name, F, MR, MM)
    if name == "dssm":
        trainer = MatchTrainer(model, mode=0, in_batch_neg=True, in_batch_neg_ratio=DSSM_SHAPE["neg"], sampler_seed=seed, n_epoch=1, device="cpu")
    else:
        trainer = CTRTrainer(model, device="cpu", n_epoch=1)

    def epoch(n_batches, s):
        xs, ys = zip(*[numpy_batch(name, s + i) for i in range(n_batches)])
        x = {k: np.concatenate([b[k] for b in xs]) for k in xs[0]}
        dl = DataLoader(TorchDataset(x, np.concatenate(ys)), batch_size=BATCH, shuffle=False)
        t0 = time.perf_counter()
        trainer.train_one_epoch(dl)
        return time.perf_counter() - t0

    t_start = time.perf_counter()
    epoch(1, seed)
    sweep = {}
    for th in sorted({t for t in (8, 32, cores) if t <= cores}):
        torch.set_num_threads(th)
        sweep[th] = epoch(1, seed + 1)
        if time.perf_counter() - t_start > 0.5 * budget_s:
            break
    best = min(sweep, key=sweep.get)
    torch.set_num_threads(best)
    n = max(3, min(steps, int((budget_s - (time.perf_counter() - t_start)) / max(sweep[best], 1e-3))))
    sec = epoch(n, seed + 2) / n
    return {"samples_per_s": BATCH / sec, "ms_per_step": sec * 1e3, "steps": n, "cores": cores, "threads": best, "kind": "reference", "thread_sweep_s_per_step": {str(k): round(v, 3) for k, v in sweep.items()},
            "sample": "%d training steps at batch %d through the reference's own %s.train_one_epoch (DataLoader(TorchDataset) + fwd + loss + bwd + Adam), %d intra-op threads (best of %s)" %
                      (n, BATCH, "MatchTrainer" if name == "dssm" else "CTRTrainer", best, sorted(sweep))}


# ---------------------------------------------------------------------------------------------------------------------
# b200 arm
# ---------------------------------------------------------------------------------------------------------------------
def _pool(name, n, seed):
    import torch
    from torch_rechub.b200.data import PackedColumns
    out = []
    for i in range(n):
        x, y = numpy_batch(name, seed + i)
        ids = [k for k, v in x.items() if v.ndim == 1 and v.dtype.kind == "i"]
        nums = [k for k, v in x.items() if v.ndim == 1 and v.dtype.kind == "f"]
        seqs = [k for k, v in x.items() if v.ndim == 2]
        import numpy as np
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        pc = PackedColumns(ids, pin(np.stack([x[k] for k in ids], axis=1)), nums, pin(np.stack([x[k] for k in nums], axis=1)) if nums else None, seqs,
                           pin(np.stack([x[k] for k in seqs], axis=1)) if seqs else None)
        out.append((pc, torch.from_numpy(y).pin_memory()))
    return out


def _time_gemms(name, device):
    import torch
    from torch_rechub.b200 import ops
    calls, flops = [], 0
    for M, N, K, am, bm, sk in GEMMS[name]:
        mk = lambda rows, cols, mn: (torch.randn((cols if mn else rows), ops._pad4(rows if mn else cols), device=device)[:, :(rows if mn else cols)])
        A, B = mk(M, K, am), mk(N, K, bm)
        out = torch.zeros(M, ops._pad4(N), device=device)
        calls.append((A, am, B, bm, M, N, K, sk, out))
        flops += 2 * M * N * K
    run = lambda: [ops.gemm3x(A, bool(am), B, bool(bm), M, N, K, split_k=sk, out=out) for (A, am, B, bm, M, N, K, sk, out) in calls]
    run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return flops, sorted(ts)[len(ts) // 2]


def run(args, bench):
    """Called by bench.py (``bench`` = that module: clock sampler, peaks, reference subprocess, exit helper)."""
    import torch
    import torch.distributed as dist
    name = args.workload
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "the b200 arm needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if name == "dssm" and world > 1:
        if rank == 0:
            print(json.dumps({"metric": METRIC[name], "unavailable": "the two-tower path runs on one GPU (cross-device negatives are not wired)", "n_gpus": world}), flush=True)
        return
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    import torch_rechub.basic.features as F
    import torch_rechub.models.matching as MM
    import torch_rechub.models.ranking as MR
    from torch_rechub.b200 import _lib, config
    from torch_rechub.b200.graph import GraphedStep
    from torch_rechub.trainers import CTRTrainer, MatchTrainer
    _lib.lib()
    config.rowwise_optimizer = True
    config.cuda_graph = True
    torch.manual_seed(2022)
    with torch.device(device):  # tables are created on the device (26 x 64 MB for dcnv2)
        model = build_model(name, F, MR, MM)
    model = model.to(device)
    pool = _pool(name, N_POOL, 2022 + 1000 * rank)
    pool_dev = [(x.to(device, non_blocking=False), y.to(device)) for x, y in pool]
    torch.cuda.synchronize()
    if name == "dssm":
        trainer = MatchTrainer(model, mode=0, in_batch_neg=True, in_batch_neg_ratio=DSSM_SHAPE["neg"], sampler_seed=2022, n_epoch=1, device=str(device))

        def step_fn(x, y):
            loss = trainer._loss(x, y)
            trainer.model.zero_grad()
            loss.backward()
            trainer.optimizer.step()
            return loss.detach()
        step_desc = "zero_grad + two towers + (B, B) scores + in-batch negatives + cross entropy + bwd + Adam (eager launches)"
    else:
        trainer = CTRTrainer(model, device=str(device), n_epoch=1)
        step_fn = GraphedStep(trainer)
        trainer._graph_step = step_fn
        step_desc = "zero_grad + fwd + BCE + bwd + optimiser (row-wise Adam on touched table rows, Adam on the dense parameters), whole step replayed as one CUDA graph"
    model.train()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_warm = max(args.warmup, 3) + 4
    for i in range(n_warm):
        step_fn(*pool_dev[i % N_POOL])
    barrier()
    sampler = bench.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        loss = step_fn(*pool_dev[(n_warm + i) % N_POOL])
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    final_loss = float(loss.item())
    _lib.check_errors(device)
    before = _lib.lib().rh_launch_count()
    if name == "dssm":
        step_fn(*pool_dev[0])
    else:
        trainer._train_step(*pool_dev[0])
    per_step_launches = int(_lib.lib().rh_launch_count() - before)
    torch.cuda.synchronize()

    class HostLoader(object):

        def __init__(self, n, start):
            self.n, self.start = n, start

        def __len__(self):
            return self.n

        def __iter__(self):
            for i in range(self.n):
                yield pool[(self.start + i) % N_POOL]

    trainer.train_one_epoch(HostLoader(max(args.warmup, 3), 0))
    epoch_s = []
    for e in range(3):
        barrier()
        t0 = time.perf_counter()
        trainer.train_one_epoch(HostLoader(args.steps, 5 + e * args.steps))
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter() - t0], device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        epoch_s.append(float(tt.item()))
    e2e_s = sorted(epoch_s)[1]
    h2d = pool[0][0].h2d_bytes() + pool[0][1].numel() * 4
    roof = None
    if rank == 0:
        flops, us = _time_gemms(name, device)
        try:
            tpeak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
            src = "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
        except Exception:
            tpeak, src = 1590.0, "fallback 1.59 PFLOP/s (B200_PROFILING.md)"
        roof = {"bound": "tensor", "kernel": "rh::gemm_tf32x3_kernel: the %d dominant GEMMs of one step (%s), graph-replayed back to back" % (len(GEMMS[name]), name), "achieved": flops / us / 1e6, "peak": tpeak,
                "unit": "TFLOP/s", "frac": flops / us / 1e6 / tpeak, "traffic": None, "us_per_step": us, "fp32_flops_per_step": flops, "peak_source": src,
                "note": "fp32-accurate 3xTF32 (3 tensor-core MMAs per fp32 product at half the bf16 rate): 1/6 of the bf16 peak is this scheme's ceiling"}
    if world > 1:
        dist.barrier()
    if rank != 0:
        bench._leave(world)
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = bench.cpu_baseline_subprocess(float(os.environ.get("RECHUB_BENCH_CPU_BUDGET_S", "40")), workload=name)
    total = BATCH * world * args.steps
    line = {
        "metric": METRIC[name], "value": total / (ms_total * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": DESCRIBE[name], "global_batch_per_gpu": BATCH, "step": step_desc, "ids": "uniform int64, %d distinct batches cycled" % N_POOL,
                   "l2": "inputs larger than L2" if name == "dcnv2" else "tables smaller than L2 (%s): the step is compute / latency bound, not HBM bound" % name,
                   "parallelism": "single GPU" if world == 1 else ("tables sharded by field over %d ranks + dp towers" % world if name == "dcnv2" else "replicated tables, data parallel over %d ranks" % world)},
        "clocks": clocks,
        "e2e": {"value": total / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8, "ms_per_step": e2e_s / args.steps * 1e3, "api": "%s.train_one_epoch(loader of pinned PackedColumns batches)" % type(trainer).__name__,
                "epoch_ms": [round(v * 1e3, 3) for v in epoch_s], "reported": "median epoch"},
        "gpu_launches": per_step_launches * args.steps, "gpu_launches_per_step": per_step_launches, "roofline": roof, "cpu_baseline": cpu, "final_loss": final_loss,
    }
    print(json.dumps(line), flush=True)
    bench._leave(world)
