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


@pytest.mark.parametrize("tile_n", [64, 128])
@pytest.mark.parametrize("M,N,K,a_mn,b_mn,bias,split_k", [
    (4096, 256, 429, False, False, True, 1), (4096, 429, 256, False, True, False, 1), (128, 256, 4096, True, True, False, 32), (256, 429, 4096, True, True, False, 16),
    (300, 70, 50, False, False, True, 1), (257, 200, 96, True, False, True, 2), (130, 64, 40, False, True, True, 1)])
def test_both_tile_widths(tile_n, M, N, K, a_mn, b_mn, bias, split_k):
    """The 128 x 64 and 128 x 128 output tiles (rh_gemm_tile_n forces one; 0 = chosen per problem) give fp32-level results for every
    operand-major combination, N / K tails, bias and split-K."""
    from torch_rechub.b200 import _lib
    L = _lib.lib()
    assert L.rh_gemm_tile_n(tile_n) == tile_n
    try:
        rel, C = _run(M, N, K, a_mn, b_mn, bias, split_k, lda_pad=1, ldb_pad=2)
    finally:
        L.rh_gemm_tile_n(0)
    assert rel < 2e-6, rel


@pytest.mark.parametrize("tma_epilogue,concat_b", [(1, 1), (0, 0), (1, 0), (0, 1)])
@pytest.mark.parametrize("tile_n", [64, 128])
@pytest.mark.parametrize("M,N,K,a_mn,b_mn,bias,split_k", [
    (4096, 256, 429, False, False, True, 1), (4096, 429, 256, False, True, False, 1), (256, 429, 4096, True, True, False, 16), (128, 256, 4096, True, True, False, 32),
    (300, 70, 50, False, False, True, 1), (257, 200, 96, True, False, True, 2), (130, 64, 40, False, True, True, 1), (1000, 36, 32, False, False, True, 1)])
def test_kernel_variants(tma_epilogue, concat_b, tile_n, M, N, K, a_mn, b_mn, bias, split_k):
    """rh_gemm_options: the shared-memory + TMA-store epilogue (bulk reduce-add for split-K; rows / columns beyond (M, N) clipped by
    the tensor map) and the concatenated-B MMA (one tcgen05.mma of width 2 BN for hi*hi and hi*lo) against the lane-per-row stores and
    the three-MMA k-step, for both tile widths, every operand-major combination, tails, bias and split-K."""
    from torch_rechub.b200 import _lib
    L = _lib.lib()
    before = L.rh_gemm_options(-1, -1)
    assert L.rh_gemm_options(tma_epilogue, concat_b) == (tma_epilogue | (concat_b << 1))
    L.rh_gemm_tile_n(tile_n)
    try:
        rel, C = _run(M, N, K, a_mn, b_mn, bias, split_k, lda_pad=1, ldb_pad=2)
    finally:
        L.rh_gemm_tile_n(0)
        L.rh_gemm_options(before & 1, (before >> 1) & 1)
    assert C.shape == (M, N)
    assert rel < 2e-6, rel


def test_tma_epilogue_stays_inside_the_16_byte_rows_of_c():
    """C with ldc > N (the tower's padded activations).  The bulk stores are clipped by the tensor map at 16-byte granularity: with
    N % 4 != 0 the tail of the last 16-byte piece of a row (columns N .. pad4(N) - 1, padding by the header's contract) may be
    written; nothing beyond pad4(N), and no row beyond M."""
    from torch_rechub.b200 import ops
    g = torch.Generator().manual_seed(3)
    M, N, K = 300, 70, 64
    A = torch.randn(M, K, generator=g).to(DEV)
    B = torch.randn(N, K, generator=g).to(DEV)
    buf = torch.full((M + 5, 76), 7.0, device=DEV)
    ops.gemm3x(A, False, B, False, M, N, K, out=buf[:M])
    assert torch.all(buf[:M, 72:] == 7.0) and torch.all(buf[M:] == 7.0)
    assert torch.allclose(buf[:M, :N], A @ B.t(), rtol=1e-4, atol=1e-4)
    buf2 = torch.full((M, 72), 7.0, device=DEV)  # N % 4 == 0: exact clipping
    ops.gemm3x(A, False, B[:68], False, M, 68, K, out=buf2)
    assert torch.all(buf2[:, 68:] == 7.0)
    assert torch.allclose(buf2[:, :68], A @ B[:68].t(), rtol=1e-4, atol=1e-4)


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


@pytest.mark.parametrize("M,N,K,bias", [(4096, 256, 429, True), (4096, 128, 256, True), (300, 72, 52, True), (129, 36, 32, False), (1000, 260, 64, True), (128, 128, 32, True)])
def test_gemm_with_column_statistics_epilogue(M, N, K, bias):
    """GEMM + BatchNorm training statistics in one launch vs fp64: C, mean, biased variance, running statistics, step counter;
    two launches in a row (the tickets must come back to zero) and a column with a huge common offset (the per-tile two-pass
    + Chan merge must not cancel)."""
    from torch_rechub.b200 import _lib
    from torch_rechub.b200.ops import _pad4
    L = _lib.lib()
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * 1.5 + 0.1).to(DEV)
    Kp = _pad4(K)
    Wbuf = torch.zeros(N, Kp, device=DEV)
    Wbuf[:, :K] = (torch.randn(N, K, generator=g) * 0.7 - 0.2).to(DEV)
    Abuf = torch.zeros(M, Kp, device=DEV)
    Abuf[:, :K] = A
    bvec = None
    if bias:
        bvec = torch.randn(N, generator=g).to(DEV)
        bvec[0] = 3000.0  # mean >> std in one column
    scratch = torch.zeros(int(L.rh_gemm_stats_scratch_floats(M, N)), device=DEV)
    rm0, rv0 = torch.rand(N, device=DEV), torch.rand(N, device=DEV) + 0.5
    ref = A.double() @ Wbuf[:, :K].double().t() + (bvec.double() if bias else 0.0)
    mean_ref, var_ref = ref.mean(0), ref.var(0, unbiased=False)
    for launch in range(2):
        C = torch.empty(M, N, device=DEV)
        stats = torch.empty(2 * N + 1, device=DEV)
        rm, rv, nbt = rm0.clone(), rv0.clone(), torch.tensor(41 + launch, device=DEV)
        _lib.check(L.rh_gemm_tf32x3_stats(Abuf.data_ptr(), Kp, 0, Wbuf.data_ptr(), Kp, 0, C.data_ptr(), N, M, N, K, _lib.ptr(bvec), stats.data_ptr(), scratch.data_ptr(), rm.data_ptr(), rv.data_ptr(),
                                          nbt.data_ptr(), 0.1, _lib.stream_ptr()), "rh_gemm_tf32x3_stats")
        torch.cuda.synchronize()
        scale = (A.double().abs() @ Wbuf[:, :K].double().abs().t()).max().item() + (3000.0 if bias else 0.0)
        assert (C.double() - ref).abs().max().item() / scale < 2e-6
        mean, var = stats[:N].double(), stats[N:2 * N].double()
        assert ((mean - mean_ref).abs() <= 2e-6 * mean_ref.abs() + 1e-5).all(), (mean - mean_ref).abs().max()
        assert ((var - var_ref).abs() <= 2e-5 * var_ref + 1e-6).all(), ((var - var_ref).abs() / var_ref).max()
        assert int(nbt) == 42 + launch and int(stats[2 * N:].view(torch.int32)) == 42 + launch
        unbiased = var_ref * (M / (M - 1.0))
        assert torch.allclose(rm.double(), 0.9 * rm0.double() + 0.1 * mean_ref, rtol=1e-5, atol=1e-5)
        assert torch.allclose(rv.double(), 0.9 * rv0.double() + 0.1 * unbiased, rtol=1e-4, atol=1e-6)
        assert int(scratch[-((N + 127) // 128):].view(torch.int32).abs().max()) == 0  # tickets back to zero


def test_tower_with_fused_statistics_matches_cpu_route():
    import copy
    from torch_rechub.basic.layers import MLP
    torch.manual_seed(9)
    cpu = MLP(64, dims=[128, 64], dropout=0.0, activation="relu")
    gpu = copy.deepcopy(cpu).to(DEV)
    x = torch.randn(513, 64) * 2 + 0.3
    yc, yg = cpu(x), gpu(x.to(DEV))
    assert (yg.cpu() - yc).abs().max().item() < 2e-5 * yc.abs().max().item() + 2e-6
    w = torch.randn(513, 1)
    (yc * w).sum().backward()
    (yg * w.to(DEV)).sum().backward()
    for (n, p), q in zip(gpu.named_parameters(), cpu.parameters()):
        scale = max(q.grad.abs().max().item(), 1e-3)
        if n.endswith("0.bias") or n.endswith("4.bias"):
            continue
        assert (p.grad.cpu() - q.grad).abs().max().item() <= 2e-4 * scale, n
    for mc, mg in zip(cpu.mlp, gpu.mlp):
        if isinstance(mc, torch.nn.BatchNorm1d):
            assert torch.allclose(mg.running_mean.cpu(), mc.running_mean, rtol=1e-5, atol=1e-6) and torch.allclose(mg.running_var.cpu(), mc.running_var, rtol=1e-5, atol=1e-6)
            assert int(mg.num_batches_tracked) == int(mc.num_batches_tracked)
