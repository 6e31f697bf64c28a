"""2-GPU NCCL test of the field-sharded CUDA route (fused owner-side gather -> all-to-all -> fused receive/unpack+FM+LR)
against a single-GPU emulation of the DataParallel semantics.  Needs >= 2 GPUs (skipped on a 1-GPU box)."""
import copy
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "torch-rechub_b200")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make(kind):
    from torch_rechub.basic.features import DenseFeature, SparseFeature
    from torch_rechub.basic.initializers import RandomNormal
    from torch_rechub.models.ranking import DCN, DeepFM
    torch.manual_seed(5)
    init = RandomNormal(0, 0.05)
    dense = [DenseFeature("I%d" % i) for i in range(3)]
    sparse = [SparseFeature("C%d" % i, 301 + i, 16, initializer=init) for i in range(7)]
    if kind == "deepfm":
        return DeepFM(dense + sparse, sparse, {"dims": [32, 16], "dropout": 0.0, "activation": "relu"})
    return DCN(dense + sparse, n_cross_layers=2, mlp_params={"dims": [32, 16]})


def _batch(rank, b=256):
    g = torch.Generator().manual_seed(100 + rank)
    x = {"I%d" % i: torch.rand(b, generator=g) for i in range(3)}
    x.update({"C%d" % i: torch.randint(0, 301, (b,), generator=g) for i in range(7)})
    return x, torch.randint(0, 2, (b,), generator=g).float()


def _worker(rank, world, port, kind, out):
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from torch_rechub.b200 import _lib
    from torch_rechub.trainers import CTRTrainer
    model = _make(kind)
    full_sd = copy.deepcopy(model.state_dict())
    trainer = CTRTrainer(model, optimizer_fn=torch.optim.SGD, optimizer_params={"lr": 0.1}, device=str(dev))
    assert trainer._dist is not None
    x, y = _batch(rank)
    model.train()
    loss = trainer._train_step({k: v.to(dev) for k, v in x.items()}, y.to(dev))
    _lib.check_errors(dev)
    sd = trainer._dist.full_state_dict()
    out[rank] = {"loss": float(loss), "sd": {k: v.detach().cpu().clone() for k, v in sd.items()}, "init": full_sd}
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0) if False else None
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("kind", ["deepfm", "dcn"])
def test_two_gpu_sharded_step(kind):
    world = 2
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), kind, out), nprocs=world, join=True)
    ref = _make(kind)
    ref.load_state_dict(out[0]["init"])
    ref = ref.to("cuda:0").train()
    opt = torch.optim.SGD(ref.parameters(), lr=0.1)
    opt.zero_grad()
    total = 0.0
    bn_mods = [m for m in ref.modules() if isinstance(m, torch.nn.BatchNorm1d)]
    saved = [(m.running_mean.clone(), m.running_var.clone()) for m in bn_mods]
    for r in range(world):
        for m, (rm, rv) in zip(bn_mods, saved):
            m.running_mean.copy_(rm)
            m.running_var.copy_(rv)
        x, y = _batch(r)
        loss = torch.nn.BCELoss()(ref({k: v.to("cuda:0") for k, v in x.items()}), y.to("cuda:0")) / world
        loss.backward()
        total += float(loss.detach())
    opt.step()
    assert abs(out[0]["loss"] - total) < 1e-5 and abs(out[1]["loss"] - total) < 1e-5
    want = {k: v.detach().cpu() for k, v in ref.state_dict().items()}
    for k, v in out[0]["sd"].items():
        if "running_" in k or "num_batches" in k or k.endswith("mlp.0.bias") or k.endswith("mlp.4.bias"):
            continue
        assert torch.allclose(v, want[k], rtol=2e-4, atol=2e-6), (k, (v - want[k]).abs().max())
    for k, v in out[1]["sd"].items():
        if "running_" in k or "num_batches" in k:
            continue
        assert torch.equal(v, out[0]["sd"][k]), k
