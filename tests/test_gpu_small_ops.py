"""rh_bce_fwd / rh_bce_bwd (the CTR trainer's BCELoss(mean), trainers/ctr_trainer.py:68,88) against torch.nn.BCELoss in float64,
including saturated probabilities (the -100 log clamp and the 1e-12 denominator clamp) and CUDA-graph replay."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 7, 4096, 100003])
def test_bce_mean_matches_torch(n):
    from torch_rechub.b200 import ops
    g = torch.Generator().manual_seed(n)
    p = torch.rand(n, generator=g)
    if n >= 7:
        p[0], p[1], p[2] = 0.0, 1.0, 1e-30
    y = torch.randint(0, 2, (n,), generator=g).float()
    pc = p.cuda().requires_grad_(True)
    crit = ops.EngineBCELoss()
    assert isinstance(crit, torch.nn.BCELoss)
    for _ in range(2):  # the ticket must come back to zero
        pc.grad = None
        loss = crit(pc, y.cuda())
        (loss * 3.0).backward()
    pr = p.double().requires_grad_(True)
    ref = torch.nn.BCELoss()(pr, y.double())
    (ref * 3.0).backward()
    assert abs(loss.item() - ref.item()) <= 2e-6 * abs(ref.item()) + 1e-7
    assert torch.allclose(pc.grad.cpu().double(), pr.grad, rtol=1e-5, atol=1e-9 * float(pr.grad.abs().max()))


def test_bce_other_modes_take_the_stock_route():
    from torch_rechub.b200 import ops
    p = torch.rand(64).cuda()
    y = torch.randint(0, 2, (64,)).float().cuda()
    assert torch.allclose(ops.EngineBCELoss(reduction="sum")(p, y), torch.nn.BCELoss(reduction="sum")(p, y))
    w = torch.rand(64).cuda()
    assert torch.allclose(ops.EngineBCELoss(weight=w)(p, y), torch.nn.BCELoss(weight=w)(p, y))
    assert torch.allclose(ops.EngineBCELoss()(p.cpu(), y.cpu()), torch.nn.BCELoss()(p.cpu(), y.cpu()))


def test_copy_segments_moves_a_packed_batch_in_one_launch():
    """rh_copy_segments behind GraphedStep._copy_batch: ids / numerics / labels of a device batch into the static buffers, incl.
    odd byte counts and unaligned (sliced) sources."""
    import ctypes
    from torch_rechub.b200 import _lib
    from torch_rechub.b200.data import PackedColumns
    from torch_rechub.b200.graph import GraphedStep
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, 1000, (513, 7), generator=g).cuda()
    nums = torch.rand(513, 3, generator=g).cuda()
    y = torch.rand(513, generator=g).cuda()
    src = PackedColumns(["a%d" % i for i in range(7)], ids, ["n%d" % i for i in range(3)], nums)
    dst = PackedColumns(src.id_names, torch.zeros_like(ids), src.num_names, torch.zeros_like(nums))
    dy = torch.zeros_like(y)
    before = _lib.lib().rh_launch_count()
    GraphedStep._copy_batch(src, y, dst, dy)
    assert _lib.lib().rh_launch_count() == before + 1
    assert torch.equal(dst.ids, ids) and torch.equal(dst.nums, nums) and torch.equal(dy, y)
    raw = torch.arange(0, 1031, dtype=torch.uint8).cuda()
    out = torch.zeros(1031, dtype=torch.uint8).cuda()
    a, b = raw[3:1030], out[1:1028]  # 1027 bytes, neither side 16-byte aligned
    d = (ctypes.c_void_p * 1)(b.data_ptr()); s = (ctypes.c_void_p * 1)(a.data_ptr()); n = (ctypes.c_int64 * 1)(1027)
    _lib.check(_lib.lib().rh_copy_segments(1, d, s, n, _lib.stream_ptr()), "rh_copy_segments")
    assert torch.equal(out[1:1028], raw[3:1030]) and out[0] == 0 and int(out[1028:].sum()) == 0
