#!/usr/bin/env python
"""bench.py — CTR samples/sec, DeepFM Criteo-shape (13 dense + 26 sparse x 1M vocab x dim 16, batch 4096 per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One JSON line on stdout (rank 0).  A "step" is one full CTRTrainer training step of the hot path over one synthetic
batch: zero_grad, fused gather+FM+LR+tile forward, MLP tower, BCE, backward incl. the sparse-gradient scatter-add into
the tables, and the optimiser (row-wise Adam on touched table rows + Adam on the dense tower).

  value     whole-job samples/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e       the same step through the public API (CTRTrainer.train_one_epoch over a host loader): every step copies
            its batch from pinned host memory (copy stream -> staging -> static inputs) and reads its loss + the
            out-of-range-id flag back to the host (async D2H, consumed by the loop one step later)
  roofline  the fused gather+FM+LR+tile forward kernel (rh_fields_fwd) timed alone with CUDA events:
            algorithmic bytes (3544 B/sample with the tile) / duration vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference   the reference's own CPU training step (oracle/ref_port.py: stock torch modules
            restating the reference's DeepFM + CTRTrainer step, validated bitwise against the live reference) on the
            host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "torch-rechub_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

N_DENSE, N_SPARSE, VOCAB, DIM, BATCH = 13, 26, 1_000_000, 16, 4096
MLP_PARAMS = {"dims": [256, 128], "dropout": 0.2, "activation": "relu"}
ALGO_BYTES_FWD_PER_SAMPLE = 3544  # ids 208 + rows 1664 + outputs 8 + flattened tile 1664
N_POOL = 64  # distinct uniform-id batches cycled: 64 x 6.8 MB of table rows = 435 MB >> 126 MB L2


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = sorted(int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) >= 6 and r[2 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# -----------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU training step (oracle/ref_port.py)
# -----------------------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, budget_s, seed=2022):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(seed)
    model = ref_port.PortDeepFM(N_DENSE, [VOCAB] * N_SPARSE, DIM, mlp_dims=tuple(MLP_PARAMS["dims"]), dropout=MLP_PARAMS["dropout"], deep_includes_sparse=True)
    g = torch.Generator().manual_seed(seed)
    batches = []
    for _ in range(4):
        dense = [torch.rand(BATCH, generator=g) for _ in range(N_DENSE)]
        ids = [torch.randint(0, VOCAB, (BATCH,), generator=g) for _ in range(N_SPARSE)]
        batches.append((dense, ids, torch.randint(0, 2, (BATCH,), generator=g).float()))
    # bounded sample: time one warm-up step, then fit as many of the requested steps as the budget allows (>= 3)
    t_warm, _ = ref_port.time_train_steps(model, batches, steps=1, warmup=max(warmup, 1) - 1 if warmup > 1 else 0)
    est = t_warm[-1]
    n = max(3, min(steps, int(budget_s / max(est, 1e-3))))
    times, loss = ref_port.time_train_steps(model, batches, steps=n, warmup=1)
    sec = sum(times) / len(times)
    return {"samples_per_s": BATCH / sec, "ms_per_step": sec * 1e3, "steps": n, "cores": cores, "loss": loss}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.steps, args.warmup, budget_s=float(os.environ.get("RECHUB_BENCH_CPU_BUDGET_S", "120")))
    sample = "%d full training steps (fwd+BCE+zero_grad+bwd+dense Adam) at batch %d on the full 26x1M x16 tables" % (r["steps"], BATCH)
    line = {
        "impl": "reference",
        "metric": "ctr_samples_per_sec_deepfm_criteo_train_step",
        "value": r["samples_per_s"],
        "unit": "samples/s",
        "n_gpus": args.gpus,
        "steps": r["steps"],
        "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"],
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "fp32",
        "data": "synthetic",
        "config": workload_config("cpu", step="reference CTRTrainer step on CPU: fwd + BCE + zero_grad + bwd (dense per-lookup table gradients) + dense Adam over all 416 M parameters"),
        "cpu_baseline": {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": r["samples_per_s"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(parallelism, step="zero_grad + fwd + BCE + bwd (scatter-add into tables) + optimiser (row-wise Adam on touched rows, Adam on the tower); tower GEMMs on tcgen05 (3xTF32, fp32-accurate); whole step replayed as one CUDA graph"):
    return {
        "workload": "DeepFM synthetic Criteo-shape: 13 dense + 26 sparse x 1M vocab x dim 16, batch 4096 per GPU, tutorial wiring (deep = dense + sparse), MLP 429-256-128-1 relu dropout 0.2",
        "global_batch_per_gpu": BATCH,
        "step": step,
        "ids": "uniform int64, %d distinct batches cycled" % N_POOL,
        "l2": "inputs larger than L2: tables 1.66 GB and %d x 6.8 MB of distinct rows per cycle vs 126 MB L2" % N_POOL,
        "parallelism": parallelism,
    }


# -----------------------------------------------------------------------------------------------------------------
# B200 arm
# -----------------------------------------------------------------------------------------------------------------
def build_model(device, init_std=1e-4, mlp_params=None):
    """The benchmarked DeepFM.  ``init_std`` / ``mlp_params`` let the parity tests build the SAME model with numerically
    non-trivial tables (N(0, 0.05)) and dropout 0 (tests/test_gpu_fullshape.py)."""
    import torch
    from torch_rechub.basic.features import DenseFeature, SparseFeature
    from torch_rechub.models.ranking import DeepFM
    torch.manual_seed(2022)
    dense = [DenseFeature("I%d" % i) for i in range(N_DENSE)]
    sparse = [SparseFeature("C%d" % i, vocab_size=VOCAB, embed_dim=DIM) for i in range(N_SPARSE)]
    # build tables directly on the device (26 x 64 MB): same distribution as RandomNormal(0, 1e-4)
    for f in sparse:
        from torch_rechub.b200.table import FieldTable
        with torch.device(device):
            t = FieldTable(VOCAB, DIM)
        with torch.no_grad():
            t.weight.normal_(0.0, init_std)
        f.embed = t
    model = DeepFM(deep_features=dense + sparse, fm_features=sparse, mlp_params=dict(MLP_PARAMS if mlp_params is None else mlp_params))
    return model.to(device), dense, sparse


ZIPF_ALPHA = 1.05  # the secondary, skewed id distribution


def make_pool(n_pool, seed, ids_dist="uniform"):
    """n_pool synthetic batches in pinned host memory, packed (ids (B,26) int64, dense (B,13) fp32, labels (B,) fp32).
    ``ids_dist``: "uniform" (the primary workload: worst case for L2) or "zipf" (rank r drawn with p ~ r^-1.05, ranks
    scattered over the vocabulary by a fixed multiplicative hash so that hot rows are not neighbours)."""
    import torch
    from torch_rechub.b200.data import PackedColumns
    g = torch.Generator().manual_seed(seed)
    pool = []
    id_names = ["C%d" % i for i in range(N_SPARSE)]
    num_names = ["I%d" % i for i in range(N_DENSE)]
    cdf = None
    if ids_dist == "zipf":
        w = torch.arange(1, VOCAB + 1, dtype=torch.float64).pow_(-ZIPF_ALPHA)
        cdf = torch.cumsum(w / w.sum(), 0)
    for _ in range(n_pool):
        if cdf is None:
            ids = torch.randint(0, VOCAB, (BATCH, N_SPARSE), generator=g)
        else:
            rank = torch.searchsorted(cdf, torch.rand(BATCH, N_SPARSE, generator=g, dtype=torch.float64)).clamp_(max=VOCAB - 1)
            ids = (rank * 2654435761) % VOCAB  # odd multiplier: a bijection on ranks for any vocabulary not divisible by it
        ids = ids.pin_memory()
        nums = torch.rand(BATCH, N_DENSE, generator=g).pin_memory()
        y = torch.randint(0, 2, (BATCH,), generator=g).float().pin_memory()
        pool.append((PackedColumns(id_names, ids, num_names, nums), y))
    return pool


def time_fused_forward_kernel(model, pool_dev, reps=20):
    """Average duration of ONE fused gather+FM+LR+tile forward launch.  The launches (one per pool batch, so every
    launch gathers different rows) are captured into a CUDA graph and the replay is timed with CUDA events on the
    launching stream: no host launch latency in the number, but the ~1 us inter-kernel gaps of a graph are included."""
    import torch
    from torch_rechub.b200 import ops
    w, b = model.linear.fc.weight, model.linear.fc.bias
    with torch.no_grad():
        plans = [model._fused_plan(x) for x, _ in pool_dev]
        for p in plans[:4]:
            ops.fused_tile(p, w, b)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep = [ops.fused_tile(p, w, b) for p in plans]
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1) / len(plans))
        del keep
    ms.sort()
    return sum(ms) / len(ms), ms[len(ms) // 2]


TOWER_GEMMS = [("fwd1", 4096, 256, 429, False, False, 1), ("fwd2", 4096, 128, 256, False, False, 1), ("dX1", 4096, 429, 256, False, True, 1), ("dX2", 4096, 256, 128, False, True, 1),
               ("dW1", 256, 429, 4096, True, True, 16), ("dW2", 128, 256, 4096, True, True, 32)]


def time_tower_gemms(device):
    """The six tower GEMMs of one step (rh_gemm_tf32x3, tcgen05) replayed from a CUDA graph: (sum of fp32-equivalent FLOPs, us)."""
    import torch
    from torch_rechub.b200 import ops
    calls, flops = [], 0
    for name, M, N, K, am, bm, sk in TOWER_GEMMS:
        def mk(rows, cols, mn):
            r, c = (cols, rows) if mn else (rows, cols)
            return torch.randn(r, (c + 3) // 4 * 4, device=device)[:, :c]
        A, B = mk(M, K, am), mk(N, K, bm)
        out = torch.zeros(M, (N + 3) // 4 * 4, device=device)
        calls.append((A, am, B, bm, M, N, K, sk, out))
        flops += 2 * M * N * K
    run = lambda: [ops.gemm3x(A, am, B, bm, M, N, K, split_k=sk, out=out) for (A, am, B, bm, M, N, K, sk, out) in calls]
    run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
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
        ts.append(e0.elapsed_time(e1) * 1e3 / 10)
    return flops, sorted(ts)[len(ts) // 2]


def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "the b200 arm needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from torch_rechub.b200 import _lib, config
    from torch_rechub.trainers import CTRTrainer
    _lib.lib()
    config.rowwise_optimizer = True
    config.cuda_graph = True

    model, dense, sparse = build_model(device)
    trainer = CTRTrainer(model, device=str(device), n_epoch=1)  # defaults: Adam lr 1e-3 weight_decay 1e-5 (ctr_trainer.py:60)
    pool = make_pool(N_POOL, seed=2022 + rank, ids_dist=args.ids)
    pool_dev = [(x.to(device, non_blocking=False), y.to(device)) for x, y in pool]
    torch.cuda.synchronize()

    from torch_rechub.b200.graph import GraphedStep
    use_graph = trainer._dist is None or os.environ.get("RECHUB_B200_DIST_CUDA_GRAPH", "1") == "1"
    if use_graph:
        step_fn = GraphedStep(trainer)
        trainer._graph_step = step_fn
    else:
        step_fn = lambda x, y: trainer._train_step(x, y)
    model.train()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput -------------------------------------------------------------------------
    n_warm = max(args.warmup, 3) + 4  # graph capture needs 3 eager steps + the capture step
    for i in range(n_warm):
        step_fn(*pool_dev[i % N_POOL])
    barrier()
    launches_before = _lib.lib().rh_launch_count()
    sampler = ClockSampler(local_rank)
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

    # kernels of OUR library per step: count one eager step (graph replays do not pass through the C ABI)
    before = _lib.lib().rh_launch_count()
    x0, y0 = pool_dev[0]
    trainer._train_step(x0, y0)
    per_step_launches = int(_lib.lib().rh_launch_count() - before)
    torch.cuda.synchronize()

    # ---- end to end through the public API ---------------------------------------------------------------------
    class HostLoader(object):

        def __init__(self, n, start):
            self.n, self.start = n, start

        def __len__(self):
            return self.n

        def __iter__(self):
            for i in range(self.n):
                yield pool[(self.start + i) % N_POOL]

    trainer.train_one_epoch(HostLoader(max(args.warmup, 3), 0))
    # K steps are ~50 ms of wall clock: one host hiccup moves the number by 20 %.  Three epochs of K steps each, the MEDIAN
    # epoch (max over ranks per epoch) is reported.
    E2E_EPOCHS = 3
    epoch_s = []
    for e in range(E2E_EPOCHS):
        barrier()
        t0 = time.perf_counter()
        trainer.train_one_epoch(HostLoader(args.steps, 7 + e * args.steps))
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        epoch_s.append(float(t.item()))
    e2e_s = sorted(epoch_s)[E2E_EPOCHS // 2]
    h2d = pool[0][0].h2d_bytes() + pool[0][1].numel() * 4

    # ---- roofline of the fused forward kernel -------------------------------------------------------------------
    roof = gemm_roof = None
    if rank == 0:
        kmodel = model if trainer._dist is None else build_model(device)[0]  # sharded run: time the kernel on a private full set of tables
        avg_ms, med_ms = time_fused_forward_kernel(kmodel, pool_dev)
        peak, peak_src = peaks()
        algo = ALGO_BYTES_FWD_PER_SAMPLE * BATCH
        achieved = algo / (avg_ms * 1e-3) / 1e9
        traffic = None
        tf = os.path.join(ROOT, "profiles", "r01_fields_fwd_dram_bytes.json")
        if os.path.exists(tf):
            try:
                traffic = json.load(open(tf)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        roof = {"bound": "hbm", "kernel": "rh::fields_fwd_v4<4,8> (fused 26-field gather + FM + LR + tile, the north_star kernel)", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "algorithmic_bytes_per_launch": algo, "avg_us": avg_ms * 1e3, "median_us": med_ms * 1e3, "peak_source": peak_src,
                "note": "latency floor, not bandwidth: 106 k random 64-B rows cost 8.8 us at any footprint (profiles/r01_microbench_gather.csv); the same kernel reaches 3.16 TB/s at B=262144 = the random-gather ceiling of this part (profiles/r01_sweep_fields_fwd.csv)"}
        gflops, gus = time_tower_gemms(device)
        try:
            tpeak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
        except Exception:
            tpeak = 1590.0
        gemm_roof = {"bound": "tensor", "kernel": "rh::gemm_tf32x3_kernel x6 (tower fwd/dX/dW; largest share of the step)", "achieved": gflops / gus / 1e6, "peak": tpeak, "unit": "TFLOP/s",
                     "frac": gflops / gus / 1e6 / tpeak, "us_per_step": gus, "fp32_flops_per_step": gflops,
                     "note": "fp32-accurate 3xTF32: 3 tensor-core MMAs per fp32 product and TF32 peak is half the bf16 peak, so 1/6 of the bf16 peak is the ceiling of this scheme; ncu tensor-pipe 10-29 % (profiles/r01b_ncu_full_summary.json)"}

    if world > 1:
        dist.barrier()
    if rank != 0:
        _leave(world)
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(steps=5, warmup=1, budget_s=float(os.environ.get("RECHUB_BENCH_CPU_BUDGET_S", "25")))
        cpu = {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["cores"], "kind": "port",
               "sample": "%d full training steps (fwd+BCE+zero_grad+bwd+dense Adam) at batch %d on the full 26x1M x16 tables (oracle/ref_port.py)" % (r["steps"], BATCH), "ms_per_step": r["ms_per_step"]}

    total_samples = BATCH * world * args.steps
    line = {
        "metric": "ctr_samples_per_sec_deepfm_criteo_train_step",
        "value": total_samples / (ms_total * 1e-3),
        "unit": "samples/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "fp32",
        "data": "synthetic",
        "config": dict(workload_config("single GPU" if world == 1 else "tables sharded by field over %d ranks + dp tower" % world),
                       **({} if args.ids == "uniform" else {"ids": "zipf(alpha=%.2f) int64, %d distinct batches cycled (secondary workload)" % (ZIPF_ALPHA, N_POOL)})),
        "clocks": clocks,
        "e2e": {"value": total_samples / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8, "ms_per_step": e2e_s / args.steps * 1e3,
                "api": "CTRTrainer.train_one_epoch(loader of pinned PackedColumns batches)", "epochs_timed": E2E_EPOCHS, "epoch_ms": [round(v * 1e3, 3) for v in epoch_s], "reported": "median epoch"},
        "gpu_launches": per_step_launches * args.steps,
        "gpu_launches_per_step": per_step_launches,
        "roofline": roof,
        "roofline_gemm": gemm_roof if rank == 0 else None,
        "cpu_baseline": cpu,
        "final_loss": final_loss,
    }
    print(json.dumps(line), flush=True)
    _leave(world)


def _leave(world):
    """Multi-rank exit: NCCL communicators captured inside live CUDA graphs can hang destroy_process_group(); every rank has
    passed the final barrier, so flush and leave without tearing the communicator down."""
    if world > 1:
        import torch
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ids", default="uniform", choices=["uniform", "zipf"], help="id distribution of the synthetic batches (uniform = the headline workload)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
