"""rh_gemm_tf32x3 (tcgen05 3xTF32, TMA-fed, TMEM accumulators) against fp64: fp32-level accuracy for every operand-major
combination and the tower's shapes, incl. K / N tails, bias and split-K."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(M, N, K, a_mn, b_mn, bias, split_k, lda_pad=0, ldb_pad=0):
    from torch_rechub.b200 import ops
    g = torch.Generator().manual_seed(M + 7 * N + 13 * K)
    A = (torch.randn(M, K, generator=g) * 1.5 + 0.1).to(DEV)
    B = (torch.randn(N, K, generator=g) * 0.7 - 0.2).to(DEV)
    bvec = torch.randn(N, generator=g).to(DEV) if bias else None

    def store(X, mn, pad):  # operand X (rows, K) stored K-major (rows, K+pad) or MN-major (K, rows+pad)
        S = X.t().contiguous() if mn else X
        r, c = S.shape
        buf = torch.zeros(r, ((c + 3) // 4) * 4 + pad * 4, device=DEV)
        buf[:, :c] = S
        return buf[:, :c]

    As, Bs = store(A, a_mn, lda_pad), store(B, b_mn, ldb_pad)
    C = ops.gemm3x(As, a_mn, Bs, b_mn, M, N, K, bias=bvec, split_k=split_k)
    ref = A.double() @ B.double().t()
    if bias:
        ref = ref + bvec.double()
    err = (C.double() - ref).abs().max().item()
    scale = (A.double().abs() @ B.double().abs().t()).max().item()
    return err / scale, C


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
def test_all_operand_majors(a_mn, b_mn):
    rel, _ = _run(256, 256, 128, a_mn, b_mn, bias=False, split_k=1)
    assert rel < 2e-6, rel


@pytest.mark.parametrize("M,N,K,a_mn,b_mn,bias,split_k", [
    (4096, 256, 429, False, False, True, 1),   # tower layer 1 forward (K tail 429 -> 448, zero filled by TMA)
    (4096, 128, 256, False, False, True, 1),   # layer 2 forward
    (4096, 429, 256, False, True, False, 1),   # dX1 = dH1 W1 (N tail)
    (4096, 256, 128, False, True, False, 1),   # dX2
    (256, 429, 4096, True, True, False, 16),   # dW1 = dH1^T X, split-K
    (128, 256, 4096, True, True, False, 32),   # dW2
    (300, 70, 50, False, False, True, 1),      # ragged everything
    (1000, 36, 32, False, False, True, 1),     # DIN attention first layer (K = 32)
])
def test_tower_shapes(M, N, K, a_mn, b_mn, bias, split_k):
    rel, C = _run(M, N, K, a_mn, b_mn, bias, split_k, lda_pad=1, ldb_pad=2)
    assert C.shape == (M, N)
    assert rel < 2e-6, rel


def test_is_more_accurate_than_tf32_and_matches_fp32_level():
    from torch_rechub.b200 import ops
    g = torch.Generator().manual_seed(1)
    A = torch.randn(512, 512, generator=g).to(DEV)
    B = torch.randn(384, 512, generator=g).to(DEV)
    ref = A.double() @ B.double().t()
    C = ops.gemm3x(A, False, B, False, 512, 384, 512)
    fp32 = A @ B.t()
    e3 = (C.double() - ref).abs().max().item()
    e32 = (fp32.double() - ref).abs().max().item()
    assert e3 < 4 * e32 + 1e-5, (e3, e32)  # same order as cuBLAS fp32; TF32 alone would be ~1e-2 here
