"""CrossNetMix / CrossNetV2 on the engine's kernels (csrc/rh_crossmix.cu + rh_gemm_tf32x3) against the reference-order loop of the
same module on CPU in float64 (basic/layers.py:440-444, 470-506): output, input gradient and every parameter gradient,
for tensor-core shapes (Criteo width 429, rank 32) and thin ones (rank 4, fewer than 128 rows -> library GEMM on the same views)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _compare(mod_cpu, mod_gpu, x, rtol):
    xc = x.double().requires_grad_(True)
    xg = x.to(DEV).requires_grad_(True)
    yc = mod_cpu.double()(xc)
    yg = mod_gpu(xg)
    assert tuple(yg.shape) == tuple(yc.shape)
    scale = yc.detach().abs().max().item()
    assert (yg.detach().cpu().double() - yc.detach()).abs().max().item() <= rtol * scale + 1e-6
    w = torch.randn(yc.shape, generator=torch.Generator().manual_seed(1)).double()
    (yc * w).sum().backward()
    (yg * w.to(DEV).float()).sum().backward()
    gs = xc.grad.abs().max().item()
    assert (xg.grad.cpu().double() - xc.grad).abs().max().item() <= 5 * rtol * gs + 1e-7, "d_x"
    for (n, pc), pg in zip(mod_cpu.named_parameters(), mod_gpu.parameters()):
        s = max(pc.grad.abs().max().item(), 1e-6)
        assert pg.grad is not None, n
        assert (pg.grad.cpu().double() - pc.grad).abs().max().item() <= 5 * rtol * s + 1e-7, (n, (pg.grad.cpu().double() - pc.grad).abs().max().item(), s)


@pytest.mark.parametrize("B,W,L,r,E", [(512, 429, 3, 32, 4), (300, 429, 2, 32, 4), (257, 64, 1, 32, 2), (64, 45, 2, 4, 4), (2, 20, 2, 3, 3), (1000, 100, 2, 8, 4)])
def test_crossnetmix_kernels_match_reference_loop(B, W, L, r, E):
    from torch_rechub.basic.layers import CrossNetMix
    torch.manual_seed(B + W)
    m = CrossNetMix(W, num_layers=L, low_rank=r, num_experts=E)
    with torch.no_grad():
        for b in m.bias:
            b.normal_(0, 0.1)
    x = torch.randn(B, W) * 0.5
    _compare(copy.deepcopy(m), copy.deepcopy(m).to(DEV), x, rtol=3e-5)


@pytest.mark.parametrize("B,W,L", [(512, 429, 3), (200, 96, 2), (50, 30, 2)])
def test_crossnetv2_kernels_match_reference_loop(B, W, L):
    from torch_rechub.basic.layers import CrossNetV2
    torch.manual_seed(B + W)
    m = CrossNetV2(W, L)
    with torch.no_grad():
        for b in m.b:
            b.normal_(0, 0.1)
    x = torch.randn(B, W) * 0.5
    _compare(copy.deepcopy(m), copy.deepcopy(m).to(DEV), x, rtol=3e-5)


def test_crossnetmix_squeeze_quirk_at_batch_one():
    from torch_rechub.basic.layers import CrossNetMix
    m = CrossNetMix(24, num_layers=2, low_rank=4, num_experts=2).to(DEV)
    assert tuple(m(torch.randn(1, 24, device=DEV)).shape) == (24,)  # the reference's squeeze() drops the batch axis (layers.py:505)
