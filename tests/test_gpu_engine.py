"""GPU tests of the training-engine pieces around the kernels: row-wise / dense optimisers against torch.optim,
CUDA-graph replay against eager steps, the packed loader path of CTRTrainer, full-size (26 x 1M x 16, B=4096)
properties through the C ABI."""
import copy

import numpy as np
import pytest
import torch

from torch_rechub.basic.features import DenseFeature, SparseFeature
from torch_rechub.basic.initializers import RandomNormal
from torch_rechub.models.ranking import DeepFM

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
INIT = RandomNormal(0, 0.05)


def small_deepfm(n_sparse=4, vocab=200, dim=16, dropout=0.0):
    dense = [DenseFeature("I%d" % i) for i in range(3)]
    sparse = [SparseFeature("C%d" % i, vocab, dim, initializer=INIT) for i in range(n_sparse)]
    return DeepFM(dense + sparse, sparse, {"dims": [32, 16], "dropout": dropout, "activation": "relu"}), dense, sparse


def batch(B, n_sparse=4, vocab=200, seed=0, hi=None):
    g = torch.Generator().manual_seed(seed)
    x = {"I%d" % i: torch.rand(B, generator=g) for i in range(3)}
    x.update({"C%d" % i: torch.randint(0, hi or vocab, (B,), generator=g) for i in range(n_sparse)})
    return x, torch.randint(0, 2, (B,), generator=g).float()


@pytest.mark.parametrize("opt_name", ["adam", "sgd", "adagrad"])
def test_hybrid_optimizer_matches_torch_on_touched_rows_and_dense_params(opt_name):
    """3 steps on the SAME batch (so every touched row is touched every step): the lazy row-wise update equals the dense
    torch optimiser on those rows (and leaves the others alone); the dense tower follows torch exactly."""
    from torch_rechub.b200 import config
    from torch_rechub.trainers import CTRTrainer
    fn, params = {"adam": (torch.optim.Adam, {"lr": 1e-2, "weight_decay": 1e-3}), "sgd": (torch.optim.SGD, {"lr": 0.1, "weight_decay": 1e-3}), "adagrad": (torch.optim.Adagrad, {"lr": 0.05})}[opt_name]
    torch.manual_seed(0)
    m_ref, _, _ = small_deepfm()
    m_fast = copy.deepcopy(m_ref)
    x, y = batch(128, hi=150)
    xd, yd = {k: v.to(DEV) for k, v in x.items()}, y.to(DEV)
    old = config.rowwise_optimizer
    try:
        config.rowwise_optimizer = False
        t_ref = CTRTrainer(m_ref, optimizer_fn=fn, optimizer_params=params, device=DEV)
        config.rowwise_optimizer = True
        t_fast = CTRTrainer(m_fast, optimizer_fn=fn, optimizer_params=params, device=DEV)
    finally:
        config.rowwise_optimizer = old
    from torch_rechub.b200.optim import HybridOptimizer
    assert isinstance(t_fast.optimizer, HybridOptimizer) and not isinstance(t_ref.optimizer, HybridOptimizer)
    m_ref.train()
    m_fast.train()
    w0 = m_ref.embedding.embed_dict["C0"].weight.detach().clone()
    for _ in range(3):
        l_ref = t_ref._train_step(xd, yd)
        l_fast = t_fast._train_step(xd, yd)
        assert abs(float(l_ref) - float(l_fast)) < 1e-5
    touched = torch.zeros(200, dtype=torch.bool, device=DEV)
    touched[xd["C0"]] = True
    wr, wf = m_ref.embedding.embed_dict["C0"].weight, m_fast.embedding.embed_dict["C0"].weight
    assert torch.allclose(wf[touched], wr[touched], rtol=2e-4, atol=2e-6)
    assert torch.equal(wf[~touched], w0.to(DEV)[~touched])  # lazy semantics: untouched rows do not move
    for (n, p), q in zip(m_fast.named_parameters(), m_ref.parameters()):
        if "embed_dict" not in n and not n.endswith("0.bias") and not n.endswith("4.bias"):  # pre-BN biases: 0-gradient noise
            assert torch.allclose(p, q, rtol=2e-4, atol=2e-6), n
    # the gradient buffer rows were consumed and re-zeroed by the optimiser
    from torch_rechub.b200 import table
    slot = table.find_slot(wf)
    assert float(slot.buffer.abs().max()) == 0.0 and not slot.pending


def test_duplicate_ids_update_once():
    from torch_rechub.b200 import config
    from torch_rechub.trainers import CTRTrainer
    torch.manual_seed(1)
    m, _, _ = small_deepfm(n_sparse=1, vocab=8)
    old = config.rowwise_optimizer
    config.rowwise_optimizer = True
    try:
        t = CTRTrainer(m, optimizer_fn=torch.optim.SGD, optimizer_params={"lr": 1.0}, device=DEV)
    finally:
        config.rowwise_optimizer = old
    x, y = batch(64, n_sparse=1, vocab=8)
    x["C0"][:] = 3  # every sample hits row 3
    xd, yd = {k: v.to(DEV) for k, v in x.items()}, y.to(DEV)
    w = m.embedding.embed_dict["C0"].weight
    w0 = w.detach().clone()
    m.train()
    loss = t._loss(xd, yd)
    m.zero_grad()
    loss.backward()
    g = w.grad.detach().clone()
    t.optimizer.step()
    assert torch.allclose(w.detach()[3], w0[3] - g[3], rtol=1e-5, atol=1e-7)  # ONE SGD step with the summed gradient
    assert torch.equal(w.detach()[[0, 1, 2, 4, 5, 6, 7]], w0[[0, 1, 2, 4, 5, 6, 7]])


def test_cuda_graph_replay_equals_eager_steps():
    from torch_rechub.b200 import config
    from torch_rechub.b200.graph import GraphedStep
    from torch_rechub.trainers import CTRTrainer
    torch.manual_seed(2)
    m_e, _, _ = small_deepfm()
    m_g = copy.deepcopy(m_e)
    old = config.rowwise_optimizer
    config.rowwise_optimizer = True
    try:
        t_e = CTRTrainer(m_e, device=DEV)
        t_g = CTRTrainer(m_g, device=DEV)
    finally:
        config.rowwise_optimizer = old
    m_e.train()
    m_g.train()
    step = GraphedStep(t_g)
    for i in range(8):
        x, y = batch(256, seed=i)
        xd, yd = {k: v.to(DEV) for k, v in x.items()}, y.to(DEV)
        le = float(t_e._train_step(xd, yd))
        lg = float(step(xd, yd))
        assert abs(le - lg) < 1e-5, (i, le, lg)
    assert step.graph is not None
    for (n, p), q in zip(m_g.named_parameters(), m_e.parameters()):
        if not n.endswith("0.bias") and not n.endswith("4.bias"):
            assert torch.allclose(p, q, rtol=1e-4, atol=1e-6), n


def test_dense_optimizer_steps_on_packed_columns_match_the_cpu_route():
    """Default (reference-semantics) optimiser over several steps with PACKED id columns: the sparse re-zeroing of the persistent
    gradient buffers must follow the strided column views, or stale rows would be applied twice."""
    from torch_rechub.b200.data import PackedColumns
    from torch_rechub.trainers import CTRTrainer
    torch.manual_seed(4)
    m_c, dense, sparse = small_deepfm()
    m_g = copy.deepcopy(m_c)
    t_c = CTRTrainer(m_c, optimizer_fn=torch.optim.SGD, optimizer_params={"lr": 0.1}, device="cpu")
    t_g = CTRTrainer(m_g, optimizer_fn=torch.optim.SGD, optimizer_params={"lr": 0.1}, device=DEV)
    m_c.train()
    m_g.train()
    init = {n: p.detach().clone() for n, p in m_c.named_parameters()}
    for i in range(4):
        x, y = batch(128, seed=20 + i)
        ids = torch.stack([x["C%d" % j] for j in range(4)], dim=1)
        nums = torch.stack([x["I%d" % j] for j in range(3)], dim=1)
        packed = PackedColumns(["C%d" % j for j in range(4)], ids.to(DEV), ["I%d" % j for j in range(3)], nums.to(DEV))
        lc = float(t_c._train_step(x, y))
        lg = float(t_g._train_step(packed, y.to(DEV)))
        assert abs(lc - lg) < 1e-5, (i, lc, lg)
    for (n, p), q in zip(m_g.named_parameters(), m_c.parameters()):
        if n.endswith("mlp.0.bias") or n.endswith("mlp.4.bias"):
            continue
        moved = (q - init[n]).abs().max().item()
        assert (p.cpu() - q).abs().max().item() <= 2e-3 * moved + 2e-6, n


def test_trainer_with_packed_loader_and_graph(tmp_path):
    from torch_rechub.b200 import config
    from torch_rechub.b200.data import PackedLoader
    from torch_rechub.trainers import CTRTrainer
    torch.manual_seed(3)
    m, _, _ = small_deepfm(dropout=0.2)
    n = 256 * 6 + 100  # a ragged last batch runs eagerly
    g = np.random.RandomState(0)
    x = {"I%d" % i: g.rand(n) for i in range(3)}
    x.update({"C%d" % i: g.randint(0, 200, n) for i in range(4)})
    yv = g.randint(0, 2, n)
    old = (config.rowwise_optimizer, config.cuda_graph)
    config.rowwise_optimizer, config.cuda_graph = True, True
    try:
        t = CTRTrainer(m, device=DEV, n_epoch=2, model_path=str(tmp_path))
        loader = PackedLoader(x, yv, batch_size=256)
        t.fit(loader, loader)
    finally:
        config.rowwise_optimizer, config.cuda_graph = old
    assert (tmp_path / "model.pth").exists()
    auc = t.evaluate(m, loader)
    assert 0.0 <= auc <= 1.0
    sd = torch.load(tmp_path / "model.pth")
    assert "embedding.embed_dict.C0.weight" in sd and sd["embedding.embed_dict.C0.weight"].shape == (200, 16)


def test_full_size_properties():
    """BASELINE configs[1] sizes (26 tables x 1M x 16, B = 4096) through the fused kernels: closed-form checks."""
    from torch_rechub.b200 import ops
    from torch_rechub.b200.table import FieldTable
    from torch_rechub.basic.layers import EmbeddingLayer
    V, D, F, B = 1_000_000, 16, 26, 4096
    feats = [SparseFeature("C%d" % i, V, D) for i in range(F)]
    for f in feats:
        with torch.device(DEV):
            f.embed = FieldTable(V, D)
    layer = EmbeddingLayer(feats).to(DEV)
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, V, (B, F), generator=g).to(DEV)
    x = {"C%d" % i: ids[:, i] for i in range(F)}
    # (1) all-ones tables: FM = 0.5 * (26^2 - 26) * 16 = 5200 exactly; LR with unit weights = 416 + bias
    with torch.no_grad():
        for t in layer.embed_dict.values():
            t.weight.fill_(1.0)
    plan = layer.build_plan(x, feats)
    for k, r in enumerate(plan.fields):
        r.fm_slot = k
    plan.n_fm, plan.fm_dim, plan.want_fm, plan.want_lr = F, D, True, True
    lw = torch.ones(1, F * D, device=DEV)
    lb = torch.full((1,), 0.5, device=DEV)
    tile, y_fm, y_lr = ops.fused_tile(plan, lw, lb)
    assert torch.equal(y_fm, torch.full((B,), 5200.0, device=DEV))
    assert torch.equal(y_lr, torch.full((B,), 416.5, device=DEV))
    assert tile.shape == (B, F * D) and float(tile.detach().min()) == 1.0 and float(tile.detach().max()) == 1.0
    # (2) row r of table f holds the value r + f/32: the gathered tile reproduces (id + f/32) bit-exactly
    with torch.no_grad():
        base = torch.arange(V, dtype=torch.float32, device=DEV).unsqueeze(1)
        for k, t in enumerate(layer.embed_dict.values()):
            t.weight.copy_((base + k / 32.0).expand(V, D))
    tile = layer(x, feats, squeeze_dim=True)
    want = (ids.float() + torch.arange(F, device=DEV).float() / 32.0).repeat_interleave(D, dim=1)
    assert torch.equal(tile, want)
    # (3) backward: d_tile = 1 everywhere -> every touched row receives exactly (multiplicity) in each column
    tile.sum().backward()
    for k in (0, F - 1):
        gk = layer.embed_dict["C%d" % k].weight.grad
        counts = torch.bincount(ids[:, k], minlength=V).float()
        assert torch.equal(gk[:, 0], counts) and torch.equal(gk[:, D - 1], counts)
    from torch_rechub.b200 import _lib
    _lib.check_errors()


def test_hybrid_optimizer_state_dict_round_trip_resumes_the_run():
    """the row-wise m / v / stamps, the step counter and the tower moments survive state_dict -> load_state_dict:
    a trainer resumed from the checkpoint after 2 steps reproduces steps 3-4 of the uninterrupted run (to fp32 rounding: the fused
    BatchNorm sums columns with atomics; a resume WITHOUT the state restarts Adam's moments and moves every weight by ~lr = 1e-2)."""
    from torch_rechub.b200 import config
    from torch_rechub.trainers import CTRTrainer
    torch.manual_seed(5)
    m_a, _, _ = small_deepfm()
    old = config.rowwise_optimizer
    config.rowwise_optimizer = True
    try:
        t_a = CTRTrainer(m_a, optimizer_params={"lr": 1e-2, "weight_decay": 1e-4}, device=DEV)
        m_a.train()
        batches = [batch(128, seed=10 + i) for i in range(4)]
        dev = [({k: v.to(DEV) for k, v in x.items()}, y.to(DEV)) for x, y in batches]
        for xd, yd in dev[:2]:
            t_a._train_step(xd, yd)
        torch.cuda.synchronize()
        sd_opt = t_a.optimizer.state_dict()
        assert sd_opt["step"] == 2 and sd_opt["rowwise"] and sd_opt["dense"]
        m_b = copy.deepcopy(m_a)  # weights + BatchNorm running statistics after step 2
        t_b = CTRTrainer(m_b, optimizer_params={"lr": 1e-2, "weight_decay": 1e-4}, device=DEV)
        t_b.optimizer.load_state_dict(sd_opt)
        m_b.train()
        for xd, yd in dev[2:]:
            la = t_a._train_step(xd, yd)
            lb = t_b._train_step(xd, yd)
            assert abs(float(la) - float(lb)) <= 1e-5 * abs(float(la))
    finally:
        config.rowwise_optimizer = old
    for (n, p), q in zip(m_a.named_parameters(), m_b.parameters()):
        assert torch.allclose(p, q, rtol=1e-4, atol=2e-5), n
    with pytest.raises(ValueError):
        t_b.optimizer.load_state_dict(torch.optim.Adam([torch.zeros(1, requires_grad=True)]).state_dict())


def test_dense_optimizer_sparse_clean_survives_reused_id_buffers():
    """with the default dense optimiser the rows to re-zero are remembered by VALUE — a caller that overwrites its device id
    buffer in place between steps must not leave stale gradient rows behind; zero_grad(set_to_none=False) cannot grow the list
    without bound."""
    from torch_rechub.b200 import config, table
    assert not config.rowwise_optimizer
    torch.manual_seed(6)
    m, _, _ = small_deepfm(n_sparse=1, vocab=64)
    m.to(DEV).train()
    w = m.embedding.embed_dict["C0"].weight
    x, y = batch(32, n_sparse=1, vocab=64, seed=1, hi=8)  # rows 0..7
    xd, yd = {k: v.to(DEV) for k, v in x.items()}, y.to(DEV)
    torch.nn.BCELoss()(m(xd), yd).backward()
    assert float(w.grad[:8].abs().sum()) > 0
    xd["C0"].add_(32)  # the caller reuses its id buffer: rows 32..39 now
    m.zero_grad(set_to_none=True)
    torch.nn.BCELoss()(m(xd), yd).backward()
    torch.cuda.synchronize()
    assert float(w.grad[:8].abs().sum()) == 0.0, "stale gradient rows: the sparse clean read the overwritten id buffer"
    assert float(w.grad[32:40].abs().sum()) > 0
    slot = table.find_slot(w)
    for _ in range(table._MAX_PENDING + 8):  # zero_grad(set_to_none=False) never reaches clean(): the list must stay bounded
        m.zero_grad(set_to_none=False)
        torch.nn.BCELoss()(m(xd), yd).backward()
    assert len(slot.pending) <= table._MAX_PENDING
    m.zero_grad(set_to_none=True)
    torch.nn.BCELoss()(m(xd), yd).backward()
    ref = w.grad.detach().clone()
    m.zero_grad(set_to_none=True)
    torch.nn.BCELoss()(m(xd), yd).backward()
    assert torch.allclose(w.grad, ref, rtol=1e-5, atol=1e-8)  # after the full clean: one step's gradient again (REDs of duplicate ids: rounding order)
