"""Profiling driver: a few warm-up steps, then ONE eager DeepFM Criteo-shape training step (the same kernels the CUDA
graph replays) inside cudaProfilerStart/Stop.  Run under ncu with --profile-from-start off (B200_PROFILING.md):

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:fields_fwd -o gpurun_out/prof python tools/profile_step.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "torch-rechub_b200"))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    batch = int(os.environ.get("PROF_BATCH", bench.BATCH))
    bench.BATCH = batch
    dev = torch.device("cuda", 0)
    from torch_rechub.b200 import config
    from torch_rechub.trainers import CTRTrainer
    config.rowwise_optimizer = True
    model, dense, sparse = bench.build_model(dev)
    trainer = CTRTrainer(model, device="cuda:0", n_epoch=1)
    pool = bench.make_pool(8, seed=1)
    pool_dev = [(x.to(dev), y.to(dev)) for x, y in pool]
    model.train()
    for i in range(4):
        trainer._train_step(*pool_dev[i])
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    trainer._train_step(*pool_dev[5])
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
