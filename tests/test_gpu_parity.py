"""GPU parity: the sm_100a kernels (through the reference-facing Python API and the C ABI underneath) against the
CPU composite path of the same package — which ``test_cpu_reference_parity.py`` pins bit-for-bit to the live
reference — on identical weights and inputs.  Tolerance (BASELINE.md §2 / SURVEY App. A.2):
``|delta| <= 1e-4 * |ref| + 1e-6`` on pre-sigmoid logits; gathered indices bit-exact.
"""
import copy

import numpy as np
import pytest
import torch

from torch_rechub.basic.features import DenseFeature, SequenceFeature, SparseFeature
from torch_rechub.basic.layers import FM, MLP, CrossNetwork, EmbeddingLayer
from torch_rechub.basic.initializers import RandomNormal
from torch_rechub.models.ranking import DCN, DIN, DCNv2, DeepFM

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
INIT = RandomNormal(0, 0.05)  # numerically non-trivial tables (the default std 1e-4 makes every logit ~0)


def logit(p):
    return torch.log(p) - torch.log1p(-p)


def assert_close(got, ref, rtol=1e-4, atol=1e-6, what=""):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    bad = (got - ref).abs() > rtol * ref.abs() + atol
    assert not bad.any(), "%s: %d/%d off, max abs err %.3e (ref max %.3e)" % (what, int(bad.sum()), bad.numel(), float((got - ref).abs().max()), float(ref.abs().max()))


def assert_grads_close(gpu_model, cpu_model, rtol=2e-4):
    """Per parameter: max abs error relative to the parameter-gradient's own max (pre-BN Linear biases have a true
    gradient of exactly 0, so they are compared against the scale of their layer's weight gradient)."""
    cpu = dict(cpu_model.named_parameters())
    for name, p in gpu_model.named_parameters():
        q = cpu[name]
        if q.grad is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        assert p.grad is not None, "no gradient on GPU for %s" % name
        g, r = p.grad.detach().double().cpu(), q.grad.detach().double()
        scale = float(r.abs().max())
        if name.endswith(".bias"):
            w = cpu.get(name[:-4] + "weight")
            if w is not None and w.grad is not None:
                scale = max(scale, float(w.grad.abs().max()))
        err = float((g - r).abs().max())
        assert err <= rtol * scale + 1e-7, "%s: grad max abs err %.3e vs scale %.3e" % (name, err, scale)


def to_dev(x):
    return {k: v.to(DEV) for k, v in x.items()}


def run_pair(make_model, x, train=True, seed=11):
    torch.manual_seed(seed)
    cpu_model = make_model()
    gpu_model = copy.deepcopy(cpu_model).to(DEV)
    cpu_model.train(train)
    gpu_model.train(train)
    y_cpu = cpu_model(x)
    y_gpu = gpu_model(to_dev(x))
    assert y_gpu.is_cuda and y_gpu.shape == y_cpu.shape and y_gpu.dtype == torch.float32
    assert_close(logit(y_gpu), logit(y_cpu), what="logit")
    assert_close(y_gpu, y_cpu, rtol=1e-4, atol=1e-7, what="prob")
    torch.manual_seed(5)
    tgt = torch.randint(0, 2, y_cpu.shape).float()
    torch.nn.BCELoss()(y_cpu, tgt).backward()
    torch.nn.BCELoss()(y_gpu, tgt.to(DEV)).backward()
    assert_grads_close(gpu_model, cpu_model)
    return cpu_model, gpu_model


def criteo_like(B, n_dense=3, n_sparse=6, dim=16, vocab=97, dense_dtype=torch.float32, id_dtype=torch.int64):
    g = torch.Generator().manual_seed(2022)
    x = {"I%d" % i: torch.rand(B, generator=g).to(dense_dtype) for i in range(n_dense)}
    x.update({"C%d" % i: torch.randint(0, vocab + i, (B,), generator=g).to(id_dtype) for i in range(n_sparse)})
    dense = [DenseFeature("I%d" % i) for i in range(n_dense)]
    sparse = [SparseFeature("C%d" % i, vocab_size=vocab + i, embed_dim=dim, initializer=INIT) for i in range(n_sparse)]
    return x, dense, sparse


@pytest.mark.parametrize("B", [1, 7, 257, 1024])
@pytest.mark.parametrize("wiring", ["tutorial", "run_criteo"])
def test_deepfm(B, wiring):
    x, dense, sparse = criteo_like(B)
    deep = dense + sparse if wiring == "tutorial" else dense
    if B == 1:  # BatchNorm1d cannot train on one row (torch raises): eval mode
        run_pair(lambda: DeepFM(deep, sparse, {"dims": [32, 16], "dropout": 0.0, "activation": "relu"}), x, train=False)
    else:
        run_pair(lambda: DeepFM(deep, sparse, {"dims": [32, 16], "dropout": 0.0, "activation": "relu"}), x)


@pytest.mark.parametrize("dim", [4, 8, 12, 16, 32, 64])
def test_deepfm_dims(dim):
    x, dense, sparse = criteo_like(130, dim=dim)
    run_pair(lambda: DeepFM(dense + sparse, sparse, {"dims": [16], "dropout": 0.0, "activation": "relu"}), x)


def test_deepfm_eval_and_dtypes():
    x, dense, sparse = criteo_like(64, dense_dtype=torch.float64, id_dtype=torch.int32)
    run_pair(lambda: DeepFM(dense + sparse, sparse, {"dims": [16, 8], "dropout": 0.3, "activation": "prelu"}), x, train=False)


def test_deepfm_scalar_dim_route():
    # embed_dim 10 is not a multiple of 4: FM/LR run as separate kernels on the gathered tile
    x, dense, sparse = criteo_like(90, dim=10)
    run_pair(lambda: DeepFM(dense + sparse, sparse, {"dims": [16], "dropout": 0.0, "activation": "relu"}), x)


@pytest.mark.parametrize("B", [2, 300])
def test_dcn(B):
    x, dense, sparse = criteo_like(B)
    run_pair(lambda: DCN(dense + sparse, n_cross_layers=3, mlp_params={"dims": [32, 16]}), x)


def test_dcn_wide():
    x, dense, sparse = criteo_like(200, n_dense=13, n_sparse=26, dim=16)
    run_pair(lambda: DCN(dense + sparse, n_cross_layers=3, mlp_params={"dims": [64, 32]}), x)


def test_dcnv2():
    x, dense, sparse = criteo_like(120)
    run_pair(lambda: DCNv2(dense + sparse, n_cross_layers=2, mlp_params={"dims": [32, 16], "dropout": 0.0, "activation": "relu"}), x)


def din_inputs(B, L=12, n_items=60, n_cates=9, n_users=20, dim=8):
    g = torch.Generator().manual_seed(7)
    lens = torch.randint(1, L + 1, (B,), generator=g)
    pos = torch.arange(L).unsqueeze(0)
    hist_i = torch.randint(1, n_items, (B, L), generator=g) * (pos < lens.unsqueeze(1))
    hist_c = torch.randint(1, n_cates, (B, L), generator=g) * (pos < lens.unsqueeze(1))
    x = {
        "target_item_id": torch.randint(1, n_items, (B,), generator=g),
        "target_cate_id": torch.randint(1, n_cates, (B,), generator=g),
        "user_id": torch.randint(1, n_users, (B,), generator=g),
        "hist_item_id": hist_i,
        "hist_cate_id": hist_c,
    }

    def make(softmax=False, dims=(16, 8)):
        feats = [SparseFeature("target_item_id", n_items, dim, initializer=INIT), SparseFeature("target_cate_id", n_cates, dim, initializer=INIT), SparseFeature("user_id", n_users, dim, initializer=INIT)]
        hist = [SequenceFeature("hist_item_id", n_items, dim, pooling="concat", shared_with="target_item_id"), SequenceFeature("hist_cate_id", n_cates, dim, pooling="concat", shared_with="target_cate_id")]
        return DIN(features=feats, history_features=hist, target_features=feats, mlp_params={"dims": list(dims)}, attention_mlp_params={"dims": list(dims), "use_softmax": softmax})

    return x, make


@pytest.mark.parametrize("softmax", [False, True])
def test_din(softmax):
    x, make = din_inputs(96)
    run_pair(lambda: make(softmax), x)


def test_din_eval():
    x, make = din_inputs(33)
    run_pair(lambda: make(False), x, train=False)


# ---- EmbeddingLayer shapes / options ----------------------------------------------------------------------
def test_embedding_layer_layouts_and_pooling():
    torch.manual_seed(3)
    B, L = 50, 9
    feats = [
        SparseFeature("a", 40, 8, initializer=INIT),
        SparseFeature("b", 30, 8, padding_idx=0, initializer=INIT),
        SparseFeature("a2", 40, 8, shared_with="a"),
        SequenceFeature("s_mean", 25, 8, pooling="mean", padding_idx=0, initializer=INIT),
        SequenceFeature("s_sum", 25, 8, pooling="sum", initializer=INIT),
        SequenceFeature("s_shared", 40, 8, pooling="mean", shared_with="a"),
        # the gradient skip follows the OWNING table's padding_idx, the pooling mask the feature's own
        SequenceFeature("s_on_b", 30, 8, pooling="mean", shared_with="b"),  # feature padding_idx None (mask = -1), table "b" skips row 0
        SequenceFeature("s_pad_on_a", 40, 8, pooling="sum", padding_idx=0, shared_with="a"),  # masks id 0, table "a" has no padding row
        DenseFeature("d0"),
        DenseFeature("dvec", embed_dim=3),
    ]
    g = torch.Generator().manual_seed(1)
    x = {
        "a": torch.randint(0, 40, (B,), generator=g),
        "b": torch.randint(0, 30, (B,), generator=g),
        "a2": torch.randint(0, 40, (B,), generator=g),
        "s_mean": torch.randint(0, 25, (B, L), generator=g) * (torch.rand(B, L, generator=g) > 0.4),
        "s_sum": torch.randint(0, 25, (B, L), generator=g),
        "s_shared": torch.randint(0, 40, (B, L), generator=g),
        "s_on_b": torch.randint(0, 30, (B, L), generator=g) * (torch.rand(B, L, generator=g) > 0.3),
        "s_pad_on_a": torch.randint(0, 40, (B, L), generator=g) * (torch.rand(B, L, generator=g) > 0.3),
        "d0": torch.rand(B, generator=g).double(),
        "dvec": torch.rand(B, 3, generator=g),
    }
    x["s_mean"][0] = 0  # a fully padded row: mean pooling divides 0 by 1e-16
    cpu = EmbeddingLayer(feats)
    gpu = copy.deepcopy(cpu).to(DEV)
    for squeeze in (True, False):
        yc = cpu(x, feats, squeeze_dim=squeeze)
        yg = gpu(to_dev(x), feats, squeeze_dim=squeeze)
        assert yg.shape == yc.shape
        assert_close(yg, yc, rtol=1e-6, atol=1e-7, what="EmbeddingLayer squeeze=%s" % squeeze)
        w = torch.randn(yc.shape, generator=g)
        cpu.zero_grad()
        gpu.zero_grad()
        (yc * w).sum().backward()
        (yg * w.to(DEV)).sum().backward()
        assert_grads_close(gpu, cpu, rtol=1e-5)
        # padding_idx row of "b" receives no gradient
        assert float(gpu.embed_dict["b"].weight.grad[0].abs().max()) == 0.0
    # dense only
    yc = cpu(x, feats[-2:], squeeze_dim=True)
    yg = gpu(to_dev(x), feats[-2:], squeeze_dim=True)
    assert_close(yg, yc, rtol=0, atol=0, what="dense only")
    with pytest.raises(ValueError):
        gpu(to_dev(x), feats[-2:], squeeze_dim=False)


def test_embedding_concat_pooling_route():
    torch.manual_seed(4)
    feats = [SequenceFeature("h1", 30, 8, pooling="concat", initializer=INIT), SequenceFeature("h2", 30, 8, pooling="concat", shared_with="h1")]
    x = {"h1": torch.randint(0, 30, (20, 5)), "h2": torch.randint(0, 30, (20, 5))}
    cpu = EmbeddingLayer(feats)
    gpu = copy.deepcopy(cpu).to(DEV)
    yc, yg = cpu(x, feats), gpu(to_dev(x), feats)
    assert yg.shape == yc.shape == (20, 2, 5, 8)
    assert torch.equal(yg.cpu(), yc)
    yc.sum().backward()
    yg.sum().backward()
    assert_grads_close(gpu, cpu, rtol=1e-6)


def test_gathered_indices_bit_exact():
    """Table row r holds the value r in every column: the gathered values ARE the indices the kernel used."""
    V, D, B = 5000, 16, 4096
    feats = [SparseFeature("f%d" % i, V, D) for i in range(26)]
    layer = EmbeddingLayer(feats)
    with torch.no_grad():
        for t in layer.embed_dict.values():
            t.weight.copy_(torch.arange(V, dtype=torch.float32).unsqueeze(1).expand(V, D))
    layer.to(DEV)
    g = torch.Generator().manual_seed(9)
    x = {"f%d" % i: torch.randint(0, V, (B,), generator=g) for i in range(26)}
    out = layer(to_dev(x), feats)  # (B, 26, D)
    want = torch.stack([x["f%d" % i] for i in range(26)], dim=1)
    for d in (0, D - 1):
        assert torch.equal(out[:, :, d].long().cpu(), want)


def test_out_of_range_id_raises_index_error():
    from torch_rechub.b200 import _lib
    feats = [SparseFeature("a", 10, 8)]
    layer = EmbeddingLayer(feats).to(DEV)
    _lib.check_errors()
    out = layer({"a": torch.tensor([1, 10, 3], device=DEV)}, feats)
    assert float(out[1].detach().abs().max()) == 0.0  # the bad row reads as zeros, memory untouched
    with pytest.raises(IndexError):
        _lib.check_errors()
    layer({"a": torch.tensor([-1], device=DEV)}, feats)
    with pytest.raises(IndexError):
        _lib.check_errors()
    _lib.check_errors()  # flag is cleared


def test_stand_alone_layers():
    torch.manual_seed(8)
    x = torch.randn(77, 9, 16)
    for reduce_sum in (True, False):
        xc = x.clone().requires_grad_(True)
        xg = x.clone().to(DEV).requires_grad_(True)
        yc, yg = FM(reduce_sum)(xc), FM(reduce_sum)(xg)
        assert_close(yg, yc, rtol=1e-5, atol=1e-5, what="FM")
        yc.sum().backward()
        yg.sum().backward()
        assert_close(xg.grad, xc.grad, rtol=1e-5, atol=1e-5, what="FM grad")
    # closed form: FM of all-ones (B, 26, 16) = 0.5 * (26^2 - 26) * 16 = 5200
    assert torch.equal(FM()(torch.ones(4, 26, 16, device=DEV)).cpu(), torch.full((4, 1), 5200.0))
    # cross network: w = 0  =>  out = x + sum_l b_l
    cn = CrossNetwork(429, 3).to(DEV)
    with torch.no_grad():
        for lin in cn.w:
            lin.weight.zero_()
        for i, b in enumerate(cn.b):
            b.fill_(0.5 * (i + 1))
    xin = torch.randn(10, 429, device=DEV)
    assert_close(cn(xin), xin + 3.0, rtol=1e-6, atol=1e-6, what="cross closed form")


@pytest.mark.parametrize("act", ["relu", "dice", "prelu", "sigmoid", "leakyrelu"])
@pytest.mark.parametrize("train", [True, False])
def test_mlp_fused_bn_act(act, train):
    torch.manual_seed(21)
    cpu = MLP(40, dims=[96, 33], dropout=0.0, activation=act)
    with torch.no_grad():
        for m in cpu.mlp:
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.5, 0.5)
                m.running_mean.uniform_(-0.2, 0.2)
                m.running_var.uniform_(0.5, 2.0)
    gpu = copy.deepcopy(cpu).to(DEV)
    cpu.train(train)
    gpu.train(train)
    x = torch.randn(300, 40) * 2 + 0.3
    yc, yg = cpu(x), gpu(x.to(DEV))
    assert_close(yg, yc, rtol=2e-5, atol=2e-6, what="MLP %s" % act)
    w = torch.randn(300, 1)
    (yc * w).sum().backward()
    (yg * w.to(DEV)).sum().backward()
    assert_grads_close(gpu, cpu, rtol=2e-4)
    if train:  # running statistics follow torch's update
        for mc, mg in zip(cpu.mlp, gpu.mlp):
            if isinstance(mc, torch.nn.BatchNorm1d):
                assert_close(mg.running_mean, mc.running_mean, rtol=1e-5, atol=1e-6, what="running_mean")
                assert_close(mg.running_var, mc.running_var, rtol=1e-5, atol=1e-6, what="running_var")
                assert int(mg.num_batches_tracked) == int(mc.num_batches_tracked)


def test_dropout_mask_statistics():
    torch.manual_seed(2)
    gpu = MLP(16, output_layer=False, dims=[256], dropout=0.25, activation="sigmoid").to(DEV)
    gpu.train()
    y = gpu(torch.randn(4096, 16, device=DEV))
    dropped = float((y == 0).float().mean())
    assert abs(dropped - 0.25) < 0.01  # sigmoid output is never exactly 0 unless dropped


def test_double_lookup_accumulates_and_sparse_zero():
    """DeepFM with deep ⊇ fm looks a table up once here; the composite route looks it up twice — same dense grad."""
    x, dense, sparse = criteo_like(64, n_sparse=3)
    torch.manual_seed(1)
    m = DeepFM(dense + sparse, sparse, {"dims": [8], "dropout": 0.0, "activation": "relu"}).to(DEV)
    xd = to_dev(x)
    for step in range(3):  # the persistent buffer must be clean at every step
        m.zero_grad()
        m(xd).sum().backward()
        g = m.embedding.embed_dict["C0"].weight.grad
        touched = torch.zeros(g.shape[0], dtype=torch.bool)
        touched[x["C0"]] = True
        assert float(g[~touched.to(DEV)].abs().max()) == 0.0
        if step == 0:
            first = g.clone()
        else:
            assert torch.allclose(g, first, rtol=1e-5, atol=1e-6)  # duplicates accumulate through atomics: order varies


@pytest.mark.parametrize("rows,k,n_extra,sig,bias", [(1, 1, 0, True, True), (5, 17, 1, False, True), (4096, 128, 2, True, True), (777, 300, 2, True, False), (64, 1024, 1, True, True)])
def test_output_head_matches_torch(rows, k, n_extra, sig, bias):
    """rh_head_fwd/bwd (Linear(K,1) + side terms + sigmoid) against the library ops it replaces (layers.py:279-280, deepfm.py:41-43)."""
    from torch_rechub.b200 import ops
    g = torch.Generator().manual_seed(rows + k)
    lin = torch.nn.Linear(k, 1, bias=bias).to(DEV)
    x = torch.randn(rows, k, generator=g).to(DEV).requires_grad_(True)
    extras = [torch.randn(rows, generator=g).to(DEV).requires_grad_(True) for _ in range(n_extra)]
    got = ops.output_head(x, lin, extras, sigmoid=sig)
    assert got is not None and got.shape == (rows,)
    x_r = x.detach().clone().requires_grad_(True)
    lin_r = copy.deepcopy(lin)
    extras_r = [e.detach().clone().requires_grad_(True) for e in extras]
    ref = lin_r(x_r).squeeze(1)
    for e in extras_r:
        ref = ref + e
    ref = torch.sigmoid(ref) if sig else ref
    assert_close(got, ref, rtol=1e-5, atol=1e-6, what="head forward")
    w_out = torch.randn(rows, generator=g).to(DEV)
    (got * w_out).sum().backward()
    (ref * w_out).sum().backward()

    def scale_close(a, b, what):
        scale = float(b.abs().max()) + 1e-12
        assert float((a - b).abs().max()) <= 2e-5 * scale + 1e-7, what

    scale_close(x.grad, x_r.grad, "d_x")
    scale_close(lin.weight.grad, lin_r.weight.grad, "d_w")
    if bias:
        scale_close(lin.bias.grad, lin_r.bias.grad, "d_b")
    for e, e_r in zip(extras, extras_r):
        scale_close(e.grad, e_r.grad, "d_extra")
