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



def repeat_regions(timed_region, cover_ms=60.0, lo=3, hi=15):
    """Run ``timed_region(r) -> (ms, result)`` for r = 0, 1, ...: once if the first region already covers ``cover_ms``, otherwise
    ceil(cover_ms / first) times clamped to [lo, hi].  The count only depends on the FIRST region's time — which the caller all-reduces
    over the ranks — so every rank runs the same number of regions.  Returns (list of ms, last result)."""
    ms, res = timed_region(0)
    out = [ms]
    n = 1 if ms >= cover_ms else int(min(hi, max(lo, -(-cover_ms // max(ms, 1e-3)))))
    for r in range(1, n):
        ms, res = timed_region(r)
        out.append(ms)
    return out, res


class ClockSampler(object):
    """SM clock + clock-event (throttle) reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML is queried in a
    thread every 2 ms (the timed region of the default run is ~40 ms; `nvidia-smi -lms 100` saw a single sample of it); when NVML
    cannot be loaded the nvidia-smi poller is the fallback."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self.thread, self.nv = index, [], None, None, None
        self.stop_flag = threading.Event()
        self.sm, self.bits, self.sm_max = [], 0, None

    def _nvml_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        try:
            ids = [int(v) for v in vis.split(",") if v.strip() != ""]
            return ids[self.index] if ids else self.index
        except Exception:
            return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._nvml_index())
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.sm.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nv is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            nv = self.nv
            masks = [nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown, nv.nvmlClocksEventReasonSwThermalSlowdown, nv.nvmlClocksEventReasonSwPowerCap]
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max, "reasons": [n for n, m in zip(self.NAMES, masks) if self.bits & m], "samples": len(sm), "source": "nvml, 2 ms period"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = sorted(int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
        reasons = [n for j, n in enumerate(self.NAMES) if any(len(r) >= 6 and r[2 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm), "source": "nvidia-smi -lms 20"}


# -----------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU training step (oracle/ref_port.py)
# -----------------------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def _reference_trainer_run(steps, budget_s, seed):
    """The UNMODIFIED reference (oracle/_ref, installed by oracle/build_ref.py) through its own public API: reference DeepFM +
    reference CTRTrainer.train_one_epoch over a DataLoader(TorchDataset) of synthetic Criteo-shape rows, on the host cores.
    Intra-op threads are swept (8 / 32 / all) on one step each and the best setting is timed."""
    import numpy as np
    import torch
    for p_ in [q for q in sys.path if q.rstrip("/").endswith("torch-rechub_b200")]:
        sys.path.remove(p_)  # this repo's package has the same import name: the reference arm must not see it
    sys.path.insert(0, REF_DIR)
    import torch_rechub
    assert os.path.abspath(torch_rechub.__file__).startswith(os.path.abspath(REF_DIR)), torch_rechub.__file__
    from torch.utils.data import DataLoader
    from torch_rechub.basic.features import DenseFeature, SparseFeature
    from torch_rechub.models.ranking import DeepFM
    from torch_rechub.trainers import CTRTrainer
    from torch_rechub.utils.data import TorchDataset
    cores = os.cpu_count() or 1
    torch.manual_seed(seed)
    torch.set_num_threads(min(32, cores))
    dense = [DenseFeature("I%d" % i) for i in range(N_DENSE)]
    sparse = [SparseFeature("C%d" % i, vocab_size=VOCAB, embed_dim=DIM) for i in range(N_SPARSE)]
    model = DeepFM(deep_features=dense + sparse, fm_features=sparse, mlp_params=dict(MLP_PARAMS))
    trainer = CTRTrainer(model, device="cpu", n_epoch=1)  # its defaults: Adam lr 1e-3 weight_decay 1e-5 (ctr_trainer.py:60)
    rng = np.random.default_rng(seed)

    def loader(n_batches):
        n = n_batches * BATCH
        x = {"I%d" % i: rng.random(n, dtype=np.float32) for i in range(N_DENSE)}
        x.update({"C%d" % i: rng.integers(0, VOCAB, n, dtype=np.int64) for i in range(N_SPARSE)})
        y = rng.integers(0, 2, n).astype(np.float32)
        return DataLoader(TorchDataset(x, y), batch_size=BATCH, shuffle=False)

    def epoch(n_batches):
        dl = loader(n_batches)
        t0 = time.perf_counter()
        trainer.train_one_epoch(dl)
        return time.perf_counter() - t0

    t_start = time.perf_counter()
    epoch(1)  # warm-up: allocates the dense gradients and Adam's state for 416 M parameters
    sweep = {}
    for th in sorted({t for t in (8, 32, cores) if t <= cores}):
        torch.set_num_threads(th)
        sweep[th] = epoch(1)
        if time.perf_counter() - t_start > 0.5 * budget_s:
            break
    best = min(sweep, key=sweep.get)
    torch.set_num_threads(best)
    left = budget_s - (time.perf_counter() - t_start)
    n = max(3, min(steps, int(left / max(sweep[best], 1e-3))))
    sec = epoch(n) / n
    return {"samples_per_s": BATCH / sec, "ms_per_step": sec * 1e3, "steps": n, "cores": cores, "threads": best, "kind": "reference",
            "thread_sweep_s_per_step": {str(k): round(v, 3) for k, v in sweep.items()},
            "sample": "%d training steps at batch %d through the reference's own CTRTrainer.train_one_epoch (DataLoader(TorchDataset) collate + fwd + BCE + zero_grad + bwd with dense per-lookup table gradients + "
                      "dense Adam over all 416 M parameters), %d intra-op threads (best of a one-step sweep over %s)" % (n, BATCH, best, sorted(sweep))}


def _port_run(steps, warmup, budget_s, seed):
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
    return {"samples_per_s": BATCH / sec, "ms_per_step": sec * 1e3, "steps": n, "cores": cores, "threads": cores, "kind": "port",
            "sample": "%d full training steps (fwd+BCE+zero_grad+bwd+dense Adam) at batch %d on the full 26x1M x16 tables (oracle/ref_port.py, the stock-torch port of the reference's step)" % (n, BATCH)}


def cpu_reference_run(steps, warmup, budget_s, seed=2022):
    """The reference's CPU training step on the host cores: the installed reference itself when oracle/_ref travelled with the
    snapshot (kind "reference"), else the stock-torch port of its step (oracle/ref_port.py, kind "port")."""
    if os.path.isdir(os.path.join(REF_DIR, "torch_rechub")):
        return _reference_trainer_run(steps, budget_s, seed)
    return _port_run(steps, warmup, budget_s, seed)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget = float(os.environ.get("RECHUB_BENCH_CPU_BUDGET_S", "150"))
    if args.workload != "deepfm":
        import bench_workloads as bw
        r = bw.reference_run(args.workload, args.steps, budget)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (the installed reference) did not travel with this snapshot; only the DeepFM step has a stock-torch port"}), flush=True)
            return
        print(json.dumps({"impl": "reference", "metric": bw.METRIC[args.workload], "value": r["samples_per_s"], "unit": "samples/s", "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup,
                          "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
                          "config": {"workload": bw.DESCRIBE[args.workload], "global_batch_per_gpu": bw.BATCH, "parallelism": "cpu"},
                          "cpu_baseline": {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["threads"], "host_cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                                           "thread_sweep_s_per_step": r.get("thread_sweep_s_per_step")},
                          "e2e": {"value": r["samples_per_s"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    r = cpu_reference_run(args.steps, args.warmup, budget_s=budget)
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
        "cpu_baseline": {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["threads"], "host_cores": r["cores"], "kind": r["kind"], "sample": r["sample"], "thread_sweep_s_per_step": r.get("thread_sweep_s_per_step")},
        "e2e": {"value": r["samples_per_s"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess(budget_s, workload="deepfm"):
    """cpu_baseline of the b200 arm: the reference arm in its own interpreter (the installed reference shares this package's import
    name, so it cannot live in this process), on a bounded sample."""
    env = dict(os.environ, RECHUB_BENCH_CPU_BUDGET_S=str(budget_s))
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "5", "--warmup", "1", "--workload", workload], env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=budget_s * 4 + 120)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
        d = json.loads(line)
        cb = d["cpu_baseline"]
        cb["ms_per_step"] = d["ms_per_step"]
        return cb
    except Exception as e:  # the baseline is a reported number, not a gate
        return {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": "reference arm failed: %r" % (e,)}


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


def kernel_times(step_fn, pool_dev, rank, steps=20, top=14):
    """Per-kernel device time of `steps` more replays of the timed step (every rank runs them: the step is collective), as seen
    by rank 0: calls per step, average duration, share of the summed kernel time.  Runs AFTER the timed regions."""
    import torch
    from collections import defaultdict
    from torch.profiler import ProfilerActivity, profile
    torch.cuda.synchronize()
    if rank != 0:
        for i in range(steps):
            step_fn(*pool_dev[i % len(pool_dev)])
        torch.cuda.synchronize()
        return None
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(steps):
            step_fn(*pool_dev[i % len(pool_dev)])
        torch.cuda.synchronize()
    agg = defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            agg[ev.name][0] += 1
            agg[ev.name][1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    total = sum(v[1] for v in agg.values())
    if total <= 0:
        return {"error": "no device events recorded"}
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]
    return {"steps": steps, "sum_us_per_step": round(total / steps, 1), "source": "torch.profiler (CUPTI), warm, rank 0, after the timed region",
            "top": [{"kernel": name[:72], "per_step": round(n / steps, 2), "avg_us": round(t / n, 2), "share": round(t / total, 3)} for name, (n, t) in rows]}


def sharded_parity_check(trainer, model, pool_dev, device, world, rank):
    """Checks the N-rank sharded engine ON THE BOX THAT PRODUCES THE NUMBER (run after the timed regions): one forward + backward
    of the field-sharded model (ids scattered to the owners, owner-side gather storing rows into the samples' GPUs, FM + LR on
    the receiver, row gradients RED to the owners, dense all-reduce) against a single-GPU emulation of the reference's
    DataParallel semantics on rank 0 — the un-sharded model with the same weights, each rank's sub-batch through its own
    BatchNorm statistics, loss = mean over the global batch (tests/test_gpu_dist.py does this at world 2 on a toy model).
    Compared: every sample's logit (|d| <= 1e-4 |ref| + 1e-6), the global loss, the all-reduced tower / LR gradients and the
    gradient rows of every table at the ids the global batch touched.  Dropout is switched off for the check on both sides
    (its stream depends on per-module step counters)."""
    import torch
    import torch.distributed as dist
    from torch_rechub.b200 import table as _table
    from torch_rechub.b200.data import PackedColumns
    eng = trainer._dist
    drops = [(m, m.p) for m in model.modules() if isinstance(m, torch.nn.Dropout)]
    for m, _ in drops:
        m.p = 0.0
    model.train()
    x, y = pool_dev[1]
    out = {"ranks": world, "what": "sharded fwd+bwd vs single-GPU DataParallel emulation (same weights, per-rank BatchNorm statistics, global-batch-mean loss)"}
    try:
        # every rank's check batch and the full tables on rank 0
        ids_all = [torch.empty_like(x.ids) for _ in range(world)]
        nums_all = [torch.empty_like(x.nums) for _ in range(world)]
        y_all = [torch.empty_like(y) for _ in range(world)]
        dist.all_gather(ids_all, x.ids.contiguous())
        dist.all_gather(nums_all, x.nums.contiguous())
        dist.all_gather(y_all, y.contiguous())
        full = eng.full_state_dict()  # collective: every rank takes part
        ref = None
        if rank == 0:
            ref = build_model(device)[0]
            ref.load_state_dict({k: v for k, v in full.items()})
            for m in ref.modules():
                if isinstance(m, torch.nn.Dropout):
                    m.p = 0.0
            ref.train()
        del full
        # ---- forward: logits of every rank's samples ----
        with torch.no_grad():
            p_mine = model(x)
        p_all = [torch.empty_like(p_mine) for _ in range(world)]
        dist.all_gather(p_all, p_mine.contiguous())
        # ---- backward of the sharded engine (no optimiser) ----
        for prm in eng.owned:
            prm.grad = None
        for prm in eng.dense_params:
            prm.grad = None
        for f in eng.fronts:
            f.lazy_clean, f.defer_barrier, f.deferred = True, False, None
        loss = trainer._loss(x, y)
        loss.backward(eng._inv_world)
        flat = torch.cat([(prm.grad if prm.grad is not None else torch.zeros_like(prm)).reshape(-1) for prm in eng.dense_params] + [(loss.detach() / world).reshape(1)])
        dist.all_reduce(flat)
        torch.cuda.synchronize()
        dist.barrier()
        # ---- the emulation on rank 0 ----
        logit = lambda pr: torch.log(pr.double()) - torch.log1p(-pr.double())
        res = torch.zeros(6, dtype=torch.float64, device=device)  # logit err, logit excess over tolerance, loss err, dense grad err, table grad err, rows compared
        uniq = []
        if rank == 0:
            ref.zero_grad()
            total = 0.0
            for r in range(world):
                xr = PackedColumns(x.id_names, ids_all[r], x.num_names, nums_all[r])
                pr = ref(xr)
                lr_, lg = logit(pr.detach()), logit(p_all[r])
                d = (lg - lr_).abs()
                res[0] = torch.maximum(res[0], d.max())
                res[1] = torch.maximum(res[1], (d - (1e-4 * lr_.abs() + 1e-6)).max())
                l = torch.nn.BCELoss()(pr, y_all[r]) / world
                l.backward()
                total += float(l.detach())
            res[2] = abs(float(flat[-1]) - total)
            off = 0
            names = dict((id(prm), n) for n, prm in model.named_parameters())
            ref_params = dict(ref.named_parameters())
            for prm in eng.dense_params:
                g = flat[off:off + prm.numel()].view_as(prm)
                off += prm.numel()
                gr = ref_params[names[id(prm)]].grad
                scale = max(float(gr.abs().max()), 1e-6)
                if names[id(prm)].endswith(".bias") and names[id(prm)][:-4] + "weight" in ref_params and ref_params[names[id(prm)][:-4] + "weight"].grad is not None:
                    scale = max(scale, float(ref_params[names[id(prm)][:-4] + "weight"].grad.abs().max()))
                res[3] = max(float(res[3]), float((g - gr).abs().max()) / scale)
        # ---- table gradients: rows at the ids the global batch touched, from each owner ----
        ids_cat = torch.cat(ids_all, dim=0)
        for front in eng.fronts:
            for j, name in enumerate(x.id_names):
                if name not in front.owner:
                    continue
                owner = front.owner[name]
                u = torch.unique(ids_cat[:, j])
                rows = torch.zeros((u.numel(), DIM), dtype=torch.float32, device=device)
                if rank == owner:
                    w = front.layer.embed_dict[name].weight
                    slot = _table.find_slot(w)
                    if slot is not None and slot.buffer is not None:
                        rows = slot.buffer.index_select(0, u).contiguous()
                dist.broadcast(rows, src=owner)
                if rank == 0:
                    gr = ref.embedding.embed_dict[name].weight.grad.index_select(0, u)
                    scale = max(float(gr.abs().max()), 1e-12)
                    res[4] = max(float(res[4]), float((rows - gr).abs().max()) / scale)
                    res[5] += u.numel()
        dist.broadcast(res, src=0)
        r_ = [float(v) for v in res.tolist()]
        # Gradient tolerance.  The check runs on the TRAINED weights (after the timed steps): probabilities are saturated, and with
        # d loss / d logit = (p - y) / N a logit error e moves that factor by up to e RELATIVE (d p = p (1 - p) e against
        # |p - y| ~ 1 - p).  So the logit errors this very check measures (inside their own 1e-4 |ref| + 1e-6 bound) put a floor of
        # ~max_logit_err under every gradient's relative error; 2 x that (two paths: FM / LR head and tower) + the 2e-4 of the tests.
        grad_tol = 2e-4 + 2.0 * r_[0]
        out.update({"max_logit_err": r_[0], "logit_tolerance": "1e-4*|ref| + 1e-6", "max_logit_excess_over_tolerance": r_[1], "loss_err": r_[2], "dense_grad_rel_err": r_[3], "table_grad_rel_err": r_[4],
                    "table_rows_compared": int(r_[5]), "grad_tolerance": grad_tol, "grad_tolerance_rule": "2e-4 + 2 * max_logit_err (of each gradient's max-abs scale)"})
        out["ok"] = bool(r_[1] <= 0.0 and r_[2] <= 1e-5 and r_[3] <= grad_tol and r_[4] <= grad_tol and r_[5] > 0)
    finally:
        for m, pv in drops:
            m.p = pv
        for prm in eng.owned:  # the un-consumed gradient rows of this check must not meet a later step
            _table.clean(prm)
            prm.grad = None
    return out


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

    def timed_region(offset):
        """EXACTLY args.steps steps between two CUDA events, a barrier + synchronize on both sides; max over ranks, in ms."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        last = None
        for i in range(args.steps):
            last = step_fn(*pool_dev[(offset + i) % N_POOL])
        ev1.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), last

    # A K-step region is a few milliseconds at the driver's K (20 steps = 3.3 ms: one NVML sample at best, and one scheduling hiccup
    # moves the number).  The region is therefore repeated — every repetition is exactly K steps between its own barriers — until
    # ~60 ms are covered (3 to 15 regions; the count follows the all-reduced time of the first one, so every rank agrees), and the
    # MEDIAN region is reported.  The clock sampler runs across all of them.
    region_ms, loss = repeat_regions(lambda r: timed_region(n_warm + r * args.steps))
    ms_total = sorted(region_ms)[len(region_ms) // 2]
    clocks = sampler.stop() if rank == 0 else None
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
        import glob
        tfs = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_fields_fwd_dram_bytes.json")))  # the latest round's ncu --set full capture
        traffic_src = None
        if tfs:
            try:
                tj = json.load(open(tfs[-1]))
                traffic = tj.get("dram_bytes_per_launch")
                traffic_src = "profiles/" + os.path.basename(tfs[-1]) + ((": " + tj["note"]) if tj.get("note") else "")
            except Exception:
                traffic = None
        roof = {"bound": "hbm", "kernel": "rh::fields_fwd_v4<4,8> (fused 26-field gather + FM + LR + tile, the north_star kernel)", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": algo, "avg_us": avg_ms * 1e3, "median_us": med_ms * 1e3, "peak_source": peak_src,
                "note": "latency-bound at this batch, not bandwidth-bound: ids arrive 1.9 us after block entry, rows at 2.7 us, the tile is stored at 5.0 us (profiles/r02_fields_trace.txt); a bare gather + store of the same 106 k rows costs 4.9 us per launch of which 2.3 us is an empty launch (profiles/r02b_microbench_gather.csv); the same kernel moves 2.5 TB/s at B=262144 (profiles/r02_sweep_fields_fwd.csv)"}
        gflops, gus = time_tower_gemms(device)
        try:
            tpeak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
        except Exception:
            tpeak = 1590.0
        gemm_roof = {"bound": "tensor", "kernel": "rh::gemm_tf32x3_kernel x6 (tower fwd/dX/dW; largest share of the step)", "achieved": gflops / gus / 1e6, "peak": tpeak, "unit": "TFLOP/s",
                     "frac": gflops / gus / 1e6 / tpeak, "us_per_step": gus, "fp32_flops_per_step": gflops,
                     "note": "fp32-accurate 3xTF32: 3 tensor-core MMAs per fp32 product and TF32 peak is half the bf16 peak, so 1/6 of the bf16 peak is the ceiling of this scheme; ncu tensor-pipe 12-29 % (profiles/r02_ncu_full_summary.json); per-CTA pipeline stamps in profiles/r02c_gemm_trace.txt"}

    # ---- warm per-kernel durations of the step that was just timed (CUPTI through torch.profiler; rank 0's view) ----------------
    ktimes = None
    if not args.no_kernel_times:
        try:
            ktimes = kernel_times(step_fn, pool_dev, rank)
        except Exception as e:  # a profiler that cannot attach must not cost the bench line
            ktimes = {"error": str(e)[:200]}

    parity = {"ok": None, "note": "single GPU: this configuration is checked against the oracle by tests/test_gpu_fullshape.py (logits, gradients, the graph-replayed row-wise Adam step)"}
    if world > 1 and trainer._dist is not None and not args.no_check:
        parity = sharded_parity_check(trainer, model, pool_dev, device, world, rank)
    if world > 1:
        dist.barrier()
    if rank != 0:
        _leave(world)
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_subprocess(float(os.environ.get("RECHUB_BENCH_CPU_BUDGET_S", "30")))

    total_samples = BATCH * world * args.steps
    line = {
        "metric": "ctr_samples_per_sec_deepfm_criteo_train_step",
        "value": total_samples / (ms_total * 1e-3),
        "unit": "samples/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps,
        "timing": {"regions": len(region_ms), "region_ms": [round(v, 4) for v in region_ms], "reported": "median region; each region = exactly `steps` steps between barriers, max over ranks"},
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
        "parity": parity,
        "kernel_times": ktimes,
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
    ap.add_argument("--no-kernel-times", action="store_true", help="skip the CUPTI per-kernel breakdown appended to the line")
    ap.add_argument("--no-check", action="store_true", help="skip the sharded-vs-single-GPU parity check of multi-GPU runs")
    ap.add_argument("--ids", default="uniform", choices=["uniform", "zipf"], help="id distribution of the synthetic batches (uniform = the headline workload)")
    ap.add_argument("--workload", default="deepfm", choices=["deepfm", "dcnv2", "din", "dssm"], help="deepfm = the headline metric of BASELINE.json (default); the others are its configs 3-5 (bench_workloads.py)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload != "deepfm":
        import bench_workloads
        bench_workloads.run(args, sys.modules[__name__])
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
