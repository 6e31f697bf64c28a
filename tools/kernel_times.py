"""Warm, in-situ per-kernel durations of the training step (torch.profiler / CUPTI) — complements the ncu launch list,
whose per-launch times are cold-cache and serialised.  Prints one line per kernel name: calls/step, avg us, share."""
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "torch-rechub_b200"))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402


def main():
    steps = int(os.environ.get("PROF_STEPS", "20"))
    use_graph = os.environ.get("PROF_GRAPH", "1") == "1"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from torch_rechub.b200 import config
    from torch_rechub.b200.graph import GraphedStep
    from torch_rechub.trainers import CTRTrainer
    config.rowwise_optimizer = True
    config.cuda_graph = use_graph
    model, dense, sparse = bench.build_model(dev)
    trainer = CTRTrainer(model, device=str(dev), n_epoch=1)
    pool = bench.make_pool(16, seed=1 + rank)
    pool_dev = [(x.to(dev), y.to(dev)) for x, y in pool]
    model.train()
    step = GraphedStep(trainer) if use_graph else (lambda x, y: trainer._train_step(x, y))
    for i in range(8):
        step(*pool_dev[i % 16])
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(steps):
            step(*pool_dev[i % 16])
        torch.cuda.synchronize()
    if rank != 0:
        torch.distributed.barrier()
        os._exit(0)
    agg = defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            agg[ev.name][0] += 1
            agg[ev.name][1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    total = sum(v[1] for v in agg.values())
    print("graph=%s steps=%d  sum of kernel time per step: %.1f us" % (use_graph, steps, total / steps))
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%6.2f/step %8.2f us avg %5.1f%%  %s" % (n / steps, t / n, 100 * t / total, name[:110]))
    sys.stdout.flush()
    if world > 1:
        torch.distributed.barrier()
        os._exit(0)


if __name__ == "__main__":
    main()
