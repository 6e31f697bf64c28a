"""Batch sweep of the fused gather+FM+LR+tile forward (rh_fields_fwd) at Criteo shape: achieved algorithmic GB/s vs batch,
next to the random-64-B-gather ceiling of tools/microbench_gather.cu.  Timed as CUDA-graph replays of one launch per distinct
id batch (no host launch latency, inter-kernel gaps included)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "torch-rechub_b200"))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    model, dense, sparse = bench.build_model(dev)
    from torch_rechub.b200 import ops
    from torch_rechub.b200.data import PackedColumns
    w, b = model.linear.fc.weight, model.linear.fc.bias
    print("batch,us_per_launch,algorithmic_GBps,frac_of_copy_peak_%.0f" % bench.peaks()[0])
    for B in (1024, 4096, 16384, 65536, 262144):
        n_pool = 8 if B >= 65536 else 32
        g = torch.Generator().manual_seed(B)
        xs = []
        for _ in range(n_pool):
            ids = torch.randint(0, bench.VOCAB, (B, bench.N_SPARSE), generator=g).to(dev)
            nums = torch.rand(B, bench.N_DENSE, generator=g).to(dev)
            xs.append(PackedColumns(["C%d" % i for i in range(bench.N_SPARSE)], ids, ["I%d" % i for i in range(bench.N_DENSE)], nums))
        with torch.no_grad():
            plans = [model._fused_plan(x) for x in xs]
            for p in plans[:2]:
                ops.fused_tile(p, w, b)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                keep = [ops.fused_tile(p, w, b) for p in plans]
            gr.replay()
            torch.cuda.synchronize()
            ts = []
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                gr.replay()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3 / n_pool)
            del keep, gr
        us = sorted(ts)[len(ts) // 2]
        gbs = bench.ALGO_BYTES_FWD_PER_SAMPLE * B / us / 1e3
        print("%d,%.2f,%.1f,%.3f" % (B, us, gbs, gbs / bench.peaks()[0]))
        del xs, plans
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
