"""2-GPU test of the field-sharded CUDA route (ids to the owners, fused owner-side gather storing rows into the samples'
GPUs, fused receive/unpack+FM+LR, row gradients RED back to the owners) against a single-GPU emulation of the
DataParallel semantics, over TWO consecutive steps (the second one exercises the sparse re-zeroing of the gradient buffers
under peer writes).  Both peer-memory variants are covered: field-major ids + direct gradients, and sample-major ids +
staged gradients.  Needs >= 2 GPUs (skipped on a 1-GPU box)."""
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


N_STEPS = 3  # the third step meets id / row buffers and flags that two earlier exchanges have used


def _batch(rank, step=0, b=256):
    g = torch.Generator().manual_seed(100 + rank + 10 * step)
    x = {"I%d" % i: torch.rand(b, generator=g) for i in range(3)}
    x.update({"C%d" % i: torch.randint(0, 301, (b,), generator=g) for i in range(7)})
    return x, torch.randint(0, 2, (b,), generator=g).float()


def _worker(rank, world, port, kind, variant, out):
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    fused = "0" if variant.endswith("-barriers") else "1"  # "-barriers": the hand-overs as separate barrier kernels (the checker of the fused route)
    variant = variant[0]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      RECHUB_B200_P2P_DIRECT_GRADS=variant, RECHUB_B200_P2P_FIELD_MAJOR=variant, RECHUB_B200_P2P_FUSED_SYNC=fused)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from torch_rechub.b200 import _lib
    from torch_rechub.trainers import CTRTrainer
    model = _make(kind)
    full_sd = copy.deepcopy(model.state_dict())
    trainer = CTRTrainer(model, optimizer_fn=torch.optim.SGD, optimizer_params={"lr": 0.1}, device=str(dev))
    from torch_rechub.b200 import config
    assert config.p2p_fused_sync == (fused == "1")
    assert trainer._dist is not None
    assert (trainer._dist.grad_pool is not None) == (variant == "1")
    model.train()
    for step in range(N_STEPS):
        x, y = _batch(rank, step)
        loss = trainer._train_step({k: v.to(dev) for k, v in x.items()}, y.to(dev))
    _lib.check_errors(dev)
    sd = trainer._dist.full_state_dict()
    out[rank] = {"loss": float(loss), "sd": {k: v.detach().cpu().clone() for k, v in sd.items()}, "init": full_sd}
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("kind,variant", [("deepfm", "1"), ("deepfm", "0"), ("dcn", "1"), ("dcn", "0"), ("deepfm", "1-barriers")])
def test_two_gpu_sharded_step(kind, variant):
    world = 2
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), kind, variant, out), nprocs=world, join=True)
    ref = _make(kind)
    ref.load_state_dict(out[0]["init"])
    ref = ref.to("cuda:0").train()
    opt = torch.optim.SGD(ref.parameters(), lr=0.1)
    for step in range(N_STEPS):
        opt.zero_grad()
        total = 0.0
        for r in range(world):  # training-mode BatchNorm normalises with each replica's own batch statistics
            x, y = _batch(r, step)
            loss = torch.nn.BCELoss()(ref({k: v.to("cuda:0") for k, v in x.items()}), y.to("cuda:0")) / world
            loss.backward()
            total += float(loss.detach())
        opt.step()
    assert abs(out[0]["loss"] - total) < 1e-5 and abs(out[1]["loss"] - total) < 1e-5
    want = {k: v.detach().cpu() for k, v in ref.state_dict().items()}
    bad = []
    for k, v in out[0]["sd"].items():
        if "running_" in k or "num_batches" in k or k.endswith("mlp.0.bias") or k.endswith("mlp.4.bias"):
            continue
        moved = (want[k] - out[0]["init"][k].cpu()).abs().max().item()  # size of the two SGD updates: the error scale that matters
        err = (v - want[k]).abs().max().item()
        if err > 2e-3 * moved + 2e-6:
            bad.append((k, err, moved))
    assert not bad, bad
    for k, v in out[1]["sd"].items():
        if "running_" in k or "num_batches" in k:
            continue
        assert torch.equal(v, out[0]["sd"][k]), k


def _worker_hybrid(rank, world, port, allreduce, out):
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), RECHUB_B200_ROWWISE_OPT="1", RECHUB_B200_P2P_ALLREDUCE=allreduce)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from torch_rechub.b200 import _lib
    from torch_rechub.trainers import CTRTrainer
    model = _make("deepfm")
    trainer = CTRTrainer(model, device=str(dev))  # Adam lr 1e-3 weight_decay 1e-5 -> row-wise Adam on the tables + fused dense step
    assert trainer._dist is not None and (trainer._dist.peer_reduce is not None) == (allreduce == "1")
    model.train()
    losses = []
    for step in range(3):
        x, y = _batch(rank, step)
        losses.append(float(trainer._train_step({k: v.to(dev) for k, v in x.items()}, y.to(dev))))
    _lib.check_errors(dev)
    sd = trainer._dist.full_state_dict()
    out[(allreduce, rank)] = {"losses": losses, "sd": {k: v.detach().cpu().clone() for k, v in sd.items()}}
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_peer_memory_allreduce_matches_nccl_route():
    """Three sharded steps with the hybrid optimiser: the engine's own all-reduce + fused dense update over NVLink peer memory
    (rh_dense_pack_signal / rh_dense_reduce_update) against the NCCL all-reduce + rh_dense_update route — same losses, same weights
    (sums in rank order vs a ring: fp32 reassociation only), and bit-identical replicated weights on both ranks."""
    world = 2
    out = mp.get_context("spawn").Manager().dict()
    for variant in ("1", "0"):
        mp.spawn(_worker_hybrid, args=(world, _free_port(), variant, out), nprocs=world, join=True)
    a, b = out[("1", 0)], out[("0", 0)]
    assert all(abs(x - y) <= 1e-6 for x, y in zip(a["losses"], b["losses"])), (a["losses"], b["losses"])
    for k, v in a["sd"].items():
        if "num_batches" in k:
            continue
        assert torch.allclose(v, b["sd"][k], rtol=1e-4, atol=2e-6), (k, (v - b["sd"][k]).abs().max())
    for k, v in out[("1", 1)]["sd"].items():
        if "running_" in k or "num_batches" in k:
            continue
        assert torch.equal(v, a["sd"][k]), k
