"""The fused training-mode BatchNorm + activation + dropout (+ output head) kernels (rh_bn_act_fused_fwd / _bwd: one launch
each way, rows in registers across a grid barrier) against (a) a float64 torch restatement of MLP.forward's
[BatchNorm1d -> activation -> Dropout] (+ Linear(., 1) + side terms + sigmoid) with autograd for the backward, and (b) the
two-kernel route of the same engine (identical dropout masks: the keep decision is a pure function of seed, step and index)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ACTS = {"relu": 1, "dice": 2, "prelu": 3, "sigmoid": 4, "leakyrelu": 5, "none": 0}


def _ref_forward(h, gamma, beta, act, alpha, eps=1e-5, dice_eps=1e-3):
    """float64: training-mode BatchNorm1d (biased variance) + activation (basic/layers.py:283-284, basic/activation.py:15-25)."""
    mu, var = h.mean(0), h.var(0, unbiased=False)
    z = (h - mu) / torch.sqrt(var + eps) * gamma + beta
    if act == "relu":
        return torch.relu(z)
    if act == "dice":
        avg = z.mean(1, keepdim=True)
        v = ((z - avg)**2 + dice_eps).sum(1, keepdim=True)
        ps = torch.sigmoid((z - avg) / torch.sqrt(v))
        return ps * z + (1 - ps) * alpha * z
    if act == "prelu":
        return torch.where(z > 0, z, alpha * z)
    if act == "sigmoid":
        return torch.sigmoid(z)
    if act == "leakyrelu":
        return torch.nn.functional.leaky_relu(z, 0.01)
    return z


def _call_fwd(h, gamma, beta, act, alpha, p_drop, seed, rm, rv, nbt, head=None):
    from torch_rechub.b200 import _lib, ops
    L = _lib.lib()
    rows, cols = h.shape
    stats = torch.empty(2 * cols + 1, device=DEV)
    scratch = ops._bn_fused_scratch(h.device, cols)
    y = None if head is not None else torch.empty(rows, cols, device=DEV)
    out = torch.empty(rows, device=DEV) if head is not None else None
    hw, hb, e0, e1, sig = head if head is not None else (None, None, None, None, 0)
    _lib.check(
        L.rh_bn_act_fused_fwd(h.data_ptr(), h.stride(0), rows, cols, 1e-5, gamma.data_ptr(), beta.data_ptr(), ACTS[act], _lib.ptr(alpha), 1e-3, p_drop, seed, rm.data_ptr(), rv.data_ptr(), nbt.data_ptr(), 0.1, stats.data_ptr(),
                              scratch.data_ptr(), _lib.ptr(y), cols, _lib.ptr(hw), _lib.ptr(hb), _lib.ptr(e0), _lib.ptr(e1), int(sig), _lib.ptr(out), _lib.stream_ptr()), "rh_bn_act_fused_fwd")
    return (out if head is not None else y), stats, scratch


@pytest.mark.parametrize("rows,cols,act", [(4096, 256, "relu"), (4096, 128, "relu"), (4096, 64, "prelu"), (1000, 256, "dice"), (513, 36, "dice"), (300, 512, "sigmoid"), (77, 128, "leakyrelu"), (2, 8, "none"),
                                          (9000, 128, "relu")])
def test_fused_forward_and_backward_against_float64(rows, cols, act):
    from torch_rechub.b200 import _lib
    L = _lib.lib()
    assert L.rh_bn_fused_supported(rows, cols, 0) == 1
    g = torch.Generator().manual_seed(rows + cols)
    h = (torch.randn(rows, cols, generator=g) * 1.7 + 0.4).to(DEV)
    if rows >= 64:
        h[:, 0] += 2000.0  # mean >> std in one column: the shifted sums must not cancel
    gamma, beta = (torch.rand(cols, generator=g) + 0.5).to(DEV), torch.randn(cols, generator=g).to(DEV)
    alpha = torch.tensor([0.25], device=DEV) if act in ("dice", "prelu") else None
    rm0, rv0 = torch.rand(cols, device=DEV), torch.rand(cols, device=DEV) + 0.5
    hd = h.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    ad = alpha.double().requires_grad_(True) if alpha is not None else None
    ref = _ref_forward(hd, gd, bd, act, ad)
    for launch in range(3):  # three times: the two parity buffers of the scratch take turns, arrivals keep counting
        rm, rv, nbt = rm0.clone(), rv0.clone(), torch.tensor(6 + launch, device=DEV)
        y, stats, scratch = _call_fwd(h, gamma, beta, act, alpha, 0.0, 0, rm, rv, nbt)
        torch.cuda.synchronize()
        scale = ref.detach().abs().max().item()
        assert (y.double() - ref.detach()).abs().max().item() <= 2e-5 * scale + 2e-6 * (2000.0 if cols > 0 else 1.0)
        mean_ref, var_ref = hd.detach().mean(0), hd.detach().var(0, unbiased=False)
        assert ((stats[:cols].double() - mean_ref).abs() <= 2e-6 * mean_ref.abs() + 1e-5).all()
        assert ((stats[cols:2 * cols].double() - var_ref).abs() <= 5e-5 * var_ref + 1e-6).all()
        assert int(nbt) == 7 + launch and int(stats[2 * cols:].view(torch.int32)) == 7 + launch
        unbiased = var_ref * (rows / (rows - 1.0))
        assert torch.allclose(rm.double(), 0.9 * rm0.double() + 0.1 * mean_ref, rtol=1e-5, atol=1e-5)
        assert torch.allclose(rv.double(), 0.9 * rv0.double() + 0.1 * unbiased, rtol=1e-4, atol=1e-6)
    # backward
    d_y = torch.randn(rows, cols, generator=g).to(DEV)
    ref.backward(d_y.double())
    from torch_rechub.b200 import ops
    d_h = torch.empty(rows, cols, device=DEV)
    gb = torch.full((3 * cols + 4,), 7.0, device=DEV)  # poisoned: every output slice must be WRITTEN
    _lib.check(
        L.rh_bn_act_fused_bwd(h.data_ptr(), cols, rows, cols, stats.data_ptr(), 1e-5, gamma.data_ptr(), beta.data_ptr(), ACTS[act], _lib.ptr(alpha), 1e-3, 0.0, 0, d_y.data_ptr(), cols, None, None, None, 0,
                              ops._bn_fused_scratch(h.device, cols).data_ptr(), d_h.data_ptr(), cols, gb.data_ptr(), gb[cols:].data_ptr(), gb[3 * cols:].data_ptr() if alpha is not None else None, None, None, None,
                              gb[2 * cols:].data_ptr(), _lib.stream_ptr()), "rh_bn_act_fused_bwd")
    torch.cuda.synchronize()
    s = hd.grad.abs().max().item()
    assert (d_h.double() - hd.grad).abs().max().item() <= 3e-4 * s + 1e-7, (d_h.double() - hd.grad).abs().max().item() / s
    assert (gb[:cols].double() - gd.grad).abs().max().item() <= 3e-4 * gd.grad.abs().max().item() + 1e-5
    assert (gb[cols:2 * cols].double() - bd.grad).abs().max().item() <= 3e-4 * bd.grad.abs().max().item() + 1e-5
    assert float(gb[2 * cols:3 * cols].abs().max()) == 0.0
    if alpha is not None:
        assert abs(float(gb[3 * cols]) - float(ad.grad)) <= 3e-4 * abs(float(ad.grad)) + 1e-4


def test_unsupported_shapes_are_reported():
    from torch_rechub.b200 import _lib
    L = _lib.lib()
    assert L.rh_bn_fused_supported(204800, 256, 0) == 0  # DIN's attention MLP: rows do not fit the register budget
    assert L.rh_bn_fused_supported(4096, 1024, 0) == 0
    assert L.rh_bn_fused_supported(4096, 130, 0) == 0
    assert L.rh_bn_fused_supported(4096, 512, 1) == 0


@pytest.mark.parametrize("act,p_drop,dims", [("relu", 0.2, [256, 128]), ("dice", 0.0, [64, 32]), ("dice", 0.3, [128, 64]), ("prelu", 0.1, [200, 36])])
def test_tower_fused_route_equals_two_kernel_route(act, p_drop, dims):
    """Same MLP, same inputs, same dropout stream: config.fused_bn / fused_bn_head on vs off — outputs, every parameter gradient,
    the input gradient and the BatchNorm running statistics, eager and under CUDA-graph replay."""
    import copy
    from torch_rechub.b200 import config
    from torch_rechub.basic.layers import MLP
    torch.manual_seed(3)
    B, K = 1000, 96
    base = MLP(K, output_layer=True, dims=dims, dropout=p_drop, activation=act).to(DEV).train()
    x = (torch.randn(B, K, device=DEV) * 1.3).requires_grad_(True)
    e0, e1 = torch.randn(B, device=DEV).requires_grad_(True), torch.randn(B, device=DEV).requires_grad_(True)
    w = torch.rand(B, device=DEV) + 0.5
    saved = (config.fused_bn, config.fused_bn_head)
    res = {}
    try:
        for mode in ("fused", "plain"):
            config.fused_bn = config.fused_bn_head = (mode == "fused")
            m = copy.deepcopy(base)
            for mod in m.modules():  # the dropout stream is keyed by the BatchNorm module: give the copies the same key
                if isinstance(mod, torch.nn.BatchNorm1d):
                    mod._rh_salt = 1000 + mod.num_features
            for t in (x, e0, e1):
                t.grad = None
            p = m.forward_head(x, (e0, e1), sigmoid=True)
            (p * w).sum().backward()
            res[mode] = (p.detach().clone(), x.grad.clone(), e0.grad.clone(), e1.grad.clone(), {k: v.grad.clone() for k, v in m.named_parameters()}, {k: v.clone() for k, v in m.named_buffers()})
    finally:
        config.fused_bn, config.fused_bn_head = saved
    a, b = res["fused"], res["plain"]
    assert (a[0] - b[0]).abs().max().item() <= 2e-6
    for i in (1, 2, 3):
        assert (a[i] - b[i]).abs().max().item() <= 2e-5 * b[i].abs().max().item() + 1e-7, i
    for k in b[4]:
        if k.endswith("alpha"):  # ONE scalar = a cancelling sum over rows x cols terms, accumulated in two different orders
            assert (a[4][k] - b[4][k]).abs().max().item() <= 5e-3 * b[4][k].abs().max().item() + 2e-4, k
            continue
        assert (a[4][k] - b[4][k]).abs().max().item() <= 3e-4 * b[4][k].abs().max().item() + 2e-6, k  # column sums: atomics vs a tree
    for k in b[5]:
        assert torch.allclose(a[5][k].float(), b[5][k].float(), rtol=1e-5, atol=1e-6), k


def test_fused_kernels_replay_in_a_cuda_graph():
    from torch_rechub.basic.layers import MLP
    torch.manual_seed(5)
    m = MLP(64, output_layer=True, dims=[128, 64], dropout=0.0, activation="relu").to(DEV).train()
    xs = [torch.randn(512, 64, device=DEV) for _ in range(3)]
    static_x = xs[0].clone().requires_grad_(True)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            m.zero_grad()
            m.forward_head(static_x, (), sigmoid=True).sum().backward()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    m.zero_grad(set_to_none=True)
    static_x.grad = None
    with torch.cuda.graph(g, stream=side):
        out = m.forward_head(static_x, (), sigmoid=True)
        out.sum().backward()
    for x in xs:
        with torch.no_grad():
            static_x.copy_(x)
        g.replay()
        torch.cuda.synchronize()
        got, gx = out.detach().clone(), static_x.grad.clone()
        # eager reference on the same weights
        xe = x.clone().requires_grad_(True)
        import copy
        me = copy.deepcopy(m)
        pe = me.forward_head(xe, (), sigmoid=True)
        pe.sum().backward()
        assert (got - pe.detach()).abs().max().item() <= 2e-6
        assert (gx - xe.grad).abs().max().item() <= 1e-5 * xe.grad.abs().max().item() + 1e-8
