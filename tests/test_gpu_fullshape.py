"""GPU parity at the EXACT shapes bench.py measures (BASELINE.json configs 2-5), against the numpy oracle:

* DeepFM 13 dense + 26 x 1M x 16, batch 4096, MLP 429-256-128-1 (the headline workload): logits, LR / tower / touched-row
  gradients, that no untouched row received gradient — and then the benchmarked STEP itself: CTRTrainer with the row-wise
  Adam + Adam on the tower, three eager steps + the CUDA-graph capture + replays, against the oracle's Adam on the same rows;
* DCNv2 (CrossNetMix, 3 cross layers, low rank 32, 4 experts) at the same Criteo shape;
* DIN at the Amazon-Electronics shape (batch 4096, L = 50, 100 k items, attention / final MLP [256, 128]);
* DSSM at the MovieLens shape (1M users x 10k items, batch 4096, in-batch hard negatives).

Dropout is 0 in all of them: the reference's dropout stream cannot be reproduced; everything else is the
benchmark's configuration.  Tolerances: logits |d| <= 1e-4 |ref| + 1e-6 (north_star); gradients <= 2e-4 of the parameter's
gradient scale; parameters after k Adam steps: see ``_close_after_adam``."""
import os
import sys

import numpy as np
import pytest
import torch

import _golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import rechub_oracle as orc  # noqa: E402

import bench  # noqa: E402  (repo root: the benchmark's own model / batch builders)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MLP0 = {"dims": [256, 128], "dropout": 0.0, "activation": "relu"}


def _logit(p):
    p = p.astype(np.float64)
    return np.log(p) - np.log1p(-p)


def _criteo_batch(seed, batch=bench.BATCH):
    g = torch.Generator().manual_seed(seed)
    x = {"I%d" % i: torch.rand(batch, generator=g) for i in range(bench.N_DENSE)}
    x.update({"C%d" % i: torch.randint(0, bench.VOCAB, (batch,), generator=g) for i in range(bench.N_SPARSE)})
    return x, torch.randint(0, 2, (batch,), generator=g).float()


def _numpy_state(model):
    return {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}


def _close_up_to_kinks(got, ref, rtol, scale, what, outlier_frac=1e-3, outlier_cap=0.05):
    """|got - ref| <= rtol * scale for all but a handful of entries.  Why a handful may differ: an element whose pre-activation is
    within fp32 rounding of ReLU's kink (|z| ~ 1e-6: about one of the 1M + 0.5M hidden activations of a 4096-batch) takes the other
    branch than in the float64 oracle, and its whole upstream gradient appears in / vanishes from the few entries it feeds (measured:
    35 of the 109 824 entries of the first layer's weight gradient from ONE element; cuBLAS fp32 differs from float64 the same way).
    Those entries are bounded by ``outlier_cap`` of the gradient scale and must stay under ``outlier_frac`` of the tensor."""
    err = np.abs(np.asarray(got, dtype=np.float64) - np.asarray(ref, dtype=np.float64))
    bad = err > rtol * scale
    if err.ndim >= 1 and err.shape[0] > 0:  # per-unit tensors (weight (out, in), BatchNorm / bias (out,)): ONE flipped element of unit u moves
        outlier_frac = max(outlier_frac, 3.0 / err.shape[0])  # everything indexed by u — allow three such units
    assert float(bad.mean()) <= outlier_frac and float(err.max()) <= outlier_cap * scale, (what, float(bad.mean()), float(err.max()), scale)


def _check_dense_grads(model, ref, rtol=2e-4, skip=()):
    for k, prm in model.named_parameters():
        if ".embed_dict." in k or k in skip:
            continue
        r = np.asarray(ref["grads"][k]).reshape(tuple(prm.shape))
        scale = max(np.abs(r).max(), 1e-6)
        if k.endswith(".bias") and k[:-4] + "weight" in ref["grads"]:  # a Linear bias in front of BatchNorm has a true gradient of 0
            scale = max(scale, np.abs(np.asarray(ref["grads"][k[:-4] + "weight"])).max())
        got = prm.grad.detach().cpu().numpy() if prm.grad is not None else np.zeros_like(r)
        _close_up_to_kinks(got, r, rtol, scale, k)


def _check_table_grads(model, ref, rtol=2e-4):
    """CompactGrad (touched rows) vs the engine's dense gradient buffer: equal on the touched rows, zero elsewhere."""
    n = 0
    for k, prm in model.named_parameters():
        if ".embed_dict." not in k:
            continue
        cg = ref["grads"][k]
        assert isinstance(cg, orc.CompactGrad)
        g = prm.grad
        assert g is not None and tuple(g.shape) == cg.shape
        ids = torch.from_numpy(cg.ids).to(g.device)
        got = g.index_select(0, ids).cpu().numpy()
        scale = max(np.abs(cg.rows).max(), 1e-9)
        # a table row's gradient belongs to ONE sample (ids rarely repeat at 1M rows): a ReLU-kink flip of one of that sample's
        # hidden units reaches its rows undiluted by any batch sum — measured 5.4 % of the tensor's largest entry on DCNv2
        # (cross + deep paths), against < 5 % for the batch-summed tower gradients; the fraction of such entries stays bounded
        _close_up_to_kinks(got, cg.rows, rtol, scale, k, outlier_cap=0.15)
        # nothing outside the touched rows: the whole buffer's |sum| equals the touched rows' |sum|
        assert abs(float(g.abs().sum()) - float(np.abs(got).sum())) <= 1e-4 * max(float(np.abs(got).sum()), 1e-9), k
        n += 1
    return n


def test_deepfm_headline_shape_logits_and_gradients_against_oracle():
    model, dense, sparse = bench.build_model(DEV, init_std=0.05, mlp_params=dict(MLP0))
    model.train()
    x, y = _criteo_batch(2022)
    sd = _numpy_state(model)
    p = model({k: v.to(DEV) for k, v in x.items()})
    torch.nn.BCELoss()(p, y.to(DEV)).backward()
    names_d, names_s = [f.name for f in dense], [f.name for f in sparse]
    with orc.compact_table_grads():
        ref = orc.deepfm_forward_backward(sd, {k: v.numpy() for k, v in x.items()}, y.numpy(), names_d, names_s, names_s, 2, train=True)
    got = _logit(p.detach().cpu().numpy())
    # Absolute floor of the logit tolerance: 429-term fp32 dot products of O(1) values + 4096-row BatchNorm statistics leave ANY fp32
    # implementation ~1e-6 away from the float64 oracle near logit 0.  It is measured, not assumed: the reference's own arithmetic (the
    # package's CPU route, bit-identical to the live reference: tests/test_cpu_api.py) on the same weights and batch gives the floor.
    import copy
    cpu_model = copy.deepcopy(model).cpu().train()
    with torch.no_grad():
        p_cpu = cpu_model({k: v for k, v in x.items()})
    ref_noise = float(np.abs(_logit(p_cpu.numpy()) - ref["logit"]).max())
    del cpu_model
    floor = max(1e-6, 3.0 * ref_noise)
    assert np.all(np.abs(got - ref["logit"]) <= 1e-4 * np.abs(ref["logit"]) + floor), (np.abs(got - ref["logit"]).max(), ref_noise)
    assert np.abs(ref["logit"]).max() > 0.05  # a non-trivial forward (tables N(0, 0.05))
    _check_dense_grads(model, ref)
    assert _check_table_grads(model, ref) == bench.N_SPARSE
    # eval mode (running statistics, no batch reduction) at the same shape
    model.eval()
    sd = _numpy_state(model)
    with torch.no_grad():
        pe = model({k: v.to(DEV) for k, v in x.items()})
    ref_e = orc.deepfm_forward_backward(sd, {k: v.numpy() for k, v in x.items()}, y.numpy(), names_d, names_s, names_s, 2, train=False, backward=False)
    got = _logit(pe.cpu().numpy())
    assert np.all(np.abs(got - ref_e["logit"]) <= 1e-4 * np.abs(ref_e["logit"]) + 1e-6)


class _LazyRows(object):
    """Adam moments of the rows an oracle run has touched so far (everything else is zero)."""

    def __init__(self, dim):
        self.dim, self.m, self.v = dim, {}, {}

    def get(self, ids):
        z = np.zeros(self.dim)
        return np.stack([self.m.get(int(i), z) for i in ids]), np.stack([self.v.get(int(i), z) for i in ids])

    def put(self, ids, m, v):
        for j, i in enumerate(ids):
            self.m[int(i)], self.v[int(i)] = m[j], v[j]


def test_deepfm_benchmarked_step_graph_replayed_rowwise_adam_against_oracle():
    """The step bench.py times: CTRTrainer._train_step under GraphedStep (3 eager steps, capture, replays) with the row-wise
    Adam on touched rows + Adam on the tower, lr 1e-3, weight_decay 1e-5 (the trainer's defaults)."""
    from torch_rechub.b200 import config
    from torch_rechub.b200.graph import GraphedStep
    from torch_rechub.trainers import CTRTrainer
    saved = (config.rowwise_optimizer, config.cuda_graph)
    config.rowwise_optimizer, config.cuda_graph = True, True
    try:
        model, dense, sparse = bench.build_model(DEV, init_std=0.05, mlp_params=dict(MLP0))
        trainer = CTRTrainer(model, device=DEV, n_epoch=1)
        model.train()
        step = GraphedStep(trainer)
        assert step.enabled
        sd0 = _numpy_state(model)
        n_steps, lr, wd = 7, 1e-3, 1e-5
        batches = [_criteo_batch(100 + t) for t in range(n_steps)]
        losses = []
        for x, y in batches:
            losses.append(float(step({k: v.to(DEV) for k, v in x.items()}, y.to(DEV)).item()))
        assert step.graph is not None and step.calls == n_steps  # steps 5.. were graph replays
        torch.cuda.synchronize()
        got = _numpy_state(model)
    finally:
        config.rowwise_optimizer, config.cuda_graph = saved

    # ---- the oracle's trainer: same batches, Adam on the tower, Adam on the touched rows ----
    names_d, names_s = [f.name for f in dense], [f.name for f in sparse]
    table_keys = ["embedding.embed_dict.%s.weight" % n for n in names_s]
    dense_keys = [k for k, _ in model.named_parameters() if ".embed_dict." not in k]

    def oracle_run(dtype, tables):
        sd = {k: (v.copy() if (tables or k not in table_keys) else v) for k, v in sd0.items()}
        mom = {k: (np.zeros(sd[k].shape), np.zeros(sd[k].shape)) for k in dense_keys}
        lazy = {k: _LazyRows(bench.DIM) for k in table_keys}
        touched = {k: set() for k in table_keys}
        ls = []
        for t, (x, y) in enumerate(batches, start=1):
            with orc.compact_table_grads():
                ref = orc.deepfm_forward_backward(sd, {k: v.numpy() for k, v in x.items()}, y.numpy(), names_d, names_s, names_s, 2, train=True, dtype=dtype)
            ls.append(float(ref["loss"]))
            for k in dense_keys:
                g = np.asarray(ref["grads"][k], dtype=np.float64).reshape(sd[k].shape)
                w, m, v = orc.adam_update(sd[k].astype(np.float64), g, mom[k][0], mom[k][1], t, lr=lr, weight_decay=wd)
                sd[k], mom[k] = w.astype(np.float32), (m, v)
            for k in table_keys:
                cg = ref["grads"][k]
                m, v = lazy[k].get(cg.ids)
                w, m, v = orc.adam_update(sd[k][cg.ids].astype(np.float64), np.asarray(cg.rows, dtype=np.float64), m, v, t, lr=lr, weight_decay=wd)
                sd[k][cg.ids] = w.astype(np.float32)
                lazy[k].put(cg.ids, m, v)
                touched[k].update(int(i) for i in cg.ids)
        return sd, ls, touched

    sd64, loss64, touched = oracle_run(np.float64, tables=True)
    for t in range(n_steps):
        assert abs(losses[t] - loss64[t]) <= 2e-5 * abs(loss64[t]) + 1e-6, (t, losses[t], loss64[t])
    # Control: the SAME trainer with float32 arithmetic in the oracle.  Adam normalises every update to ~lr, so trajectories that
    # differ by fp32 rounding (and by the ReLU-kink flips it causes, see _close_up_to_kinks) drift apart by a few per cent of the
    # distance travelled within a handful of steps; how far is measured here, not assumed, and the engine must stay within a
    # small multiple of it.
    sd32, _, _ = oracle_run(np.float32, tables=True)
    rms = lambda a: float(np.sqrt(np.mean(np.square(np.asarray(a, dtype=np.float64)))))

    def trimmed(err, drop):
        """rms of ``err`` without its ``drop`` worst units (first axis).  A ReLU-kink flip (see _close_up_to_kinks) changes the
        gradient of everything indexed by ONE hidden unit / ONE sample, Adam turns that into a persistent ~lr-sized detour of those
        entries, and which elements flip differs between any two fp32 implementations — measured: three flipped units of 256 put
        the plain rms of the first layer's weight at 20x the float32 control while every other unit sat at the control's level."""
        e = np.asarray(err, dtype=np.float64).reshape(err.shape[0], -1) if np.ndim(err) >= 1 else np.asarray(err, dtype=np.float64).reshape(1, 1)
        per_unit = np.sqrt(np.mean(np.square(e), axis=1))
        keep = np.sort(per_unit)[:max(1, len(per_unit) - drop)]
        return float(np.sqrt(np.mean(np.square(keep))))

    for k in dense_keys:
        n_units = sd64[k].shape[0] if np.ndim(sd64[k]) >= 1 else 1
        drop = 0 if n_units < 16 else max(3, n_units // 50)
        travel, ctl, mine = rms(sd64[k] - sd0[k]), rms(sd32[k] - sd64[k]), trimmed(got[k] - sd64[k], drop)
        assert mine <= 4.0 * ctl + 5e-3 * travel + 1e-7, (k, mine, ctl, travel)
        assert np.abs(got[k] - sd64[k]).max() <= 2.0 * lr * n_steps + 1e-6, k
    for k in table_keys:
        ids = np.fromiter(touched[k], dtype=np.int64)
        travel, ctl = rms(sd64[k][ids] - sd0[k][ids]), rms(sd32[k][ids] - sd64[k][ids])
        mine = trimmed(got[k][ids] - sd64[k][ids], max(3, len(ids) // 500))
        assert mine <= 4.0 * ctl + 5e-3 * travel + 1e-7, (k, mine, ctl, travel)
        assert np.abs(got[k][ids] - sd64[k][ids]).max() <= 2.0 * lr * n_steps + 1e-6, k
        untouched = np.ones(sd64[k].shape[0], dtype=bool)
        untouched[ids] = False
        assert np.array_equal(got[k][untouched], sd0[k][untouched]), k  # lazy mode: untouched rows do not move
    for k in got:  # BatchNorm running statistics of the last layer moved and stayed finite
        if k.endswith("running_var"):
            assert np.all(np.isfinite(got[k])) and not np.array_equal(got[k], sd0[k])


def test_dcnv2_criteo_shape_against_oracle():
    from torch_rechub.basic.features import DenseFeature, SparseFeature
    from torch_rechub.b200.table import FieldTable
    from torch_rechub.models.ranking import DCNv2
    torch.manual_seed(7)
    dense = [DenseFeature("I%d" % i) for i in range(bench.N_DENSE)]
    sparse = [SparseFeature("C%d" % i, vocab_size=bench.VOCAB, embed_dim=bench.DIM) for i in range(bench.N_SPARSE)]
    for f in sparse:
        with torch.device(DEV):
            t = FieldTable(bench.VOCAB, bench.DIM)
        with torch.no_grad():
            t.weight.normal_(0.0, 0.05)
        f.embed = t
    model = DCNv2(dense + sparse, n_cross_layers=3, mlp_params=dict(MLP0)).to(DEV).train()  # examples/ranking/run_criteo.py:70
    x, y = _criteo_batch(11)
    sd = _numpy_state(model)
    p = model({k: v.to(DEV) for k, v in x.items()})
    torch.nn.BCELoss()(p, y.to(DEV)).backward()
    with orc.compact_table_grads():
        ref = orc.dcnv2_forward_backward(sd, {k: v.numpy() for k, v in x.items()}, y.numpy(), [f.name for f in dense], [f.name for f in sparse], 3, 2)
    got = _logit(p.detach().cpu().numpy())
    assert np.all(np.abs(got - ref["logit"]) <= 1e-4 * np.abs(ref["logit"]) + 1e-6), np.abs(got - ref["logit"]).max()
    _check_dense_grads(model, ref, rtol=3e-4)
    assert _check_table_grads(model, ref, rtol=3e-4) == bench.N_SPARSE


def _din_model_and_batch(batch, seq_len, n_items, n_cates, n_users, dims, seed):
    from torch_rechub.basic.features import SequenceFeature, SparseFeature
    from torch_rechub.basic.initializers import RandomNormal
    from torch_rechub.models.ranking import DIN
    init = RandomNormal(0, 0.05)
    torch.manual_seed(seed)
    feats = [SparseFeature("target_item_id", n_items + 1, 8, initializer=init), SparseFeature("target_cate_id", n_cates + 1, 8, initializer=init), SparseFeature("user_id", n_users + 1, 8, initializer=init)]
    hist = [SequenceFeature("hist_item_id", n_items + 1, 8, pooling="concat", shared_with="target_item_id"), SequenceFeature("hist_cate_id", n_cates + 1, 8, pooling="concat", shared_with="target_cate_id")]
    model = DIN(features=feats, history_features=hist, target_features=feats, mlp_params={"dims": list(dims)}, attention_mlp_params={"dims": list(dims)})
    g = torch.Generator().manual_seed(seed + 1)
    lens = torch.randint(1, seq_len + 1, (batch,), generator=g)
    keep = torch.arange(seq_len).unsqueeze(0) < lens.unsqueeze(1)  # post-padding with 0 (utils/data.py:175-176)
    x = {
        "target_item_id": torch.randint(1, n_items + 1, (batch,), generator=g),
        "target_cate_id": torch.randint(1, n_cates + 1, (batch,), generator=g),
        "user_id": torch.randint(1, n_users + 1, (batch,), generator=g),
        "hist_item_id": torch.randint(1, n_items + 1, (batch, seq_len), generator=g) * keep,
        "hist_cate_id": torch.randint(1, n_cates + 1, (batch, seq_len), generator=g) * keep,
    }
    return model, x, torch.randint(0, 2, (batch,), generator=g).float()


def test_din_amazon_shape_against_oracle():
    """batch 4096, L = 50, 100 k items, 1 k categories, 190 k users, D = 8, attention and final MLP [256, 128]."""
    model, x, y = _din_model_and_batch(4096, 50, 100_000, 1000, 190_000, (256, 128), seed=21)
    model = model.to(DEV).train()
    sd = _numpy_state(model)
    p = model({k: v.to(DEV) for k, v in x.items()})
    torch.nn.BCELoss()(p, y.to(DEV)).backward()
    ref = orc.din_forward_backward(sd, {k: v.numpy() for k, v in x.items()}, y.numpy(), _golden.DIN_FEATURES, _golden.DIN_HISTORY, _golden.DIN_FEATURES, _golden.DIN_SHARED, 2, 2)
    got = _logit(p.detach().cpu().numpy())
    assert np.all(np.abs(got - ref["logit"]) <= 1e-4 * np.abs(ref["logit"]) + 1e-6), np.abs(got - ref["logit"]).max()
    for k, prm in model.named_parameters():
        r = np.asarray(ref["grads"][k]).reshape(tuple(prm.shape))
        scale = max(np.abs(r).max(), 1e-6)
        if k.endswith(".bias") and k[:-4] + "weight" in ref["grads"]:
            scale = max(scale, np.abs(np.asarray(ref["grads"][k[:-4] + "weight"])).max())
        gg = prm.grad.detach().cpu().numpy()
        _close_up_to_kinks(gg, r, 5e-4, scale, k)  # 204 800-row reductions in fp32


def test_dssm_movielens_shape_against_oracle():
    """BASELINE config 5: 1M users x 10k items, batch 4096, in-batch hard negatives (examples/matching/run_ml_dssm.py:52-55 wiring,
    user tower = user_id + mean-pooled history sharing the item table)."""
    import torch_rechub.basic.features as F
    import torch_rechub.models.matching as M
    from torch_rechub.basic.initializers import RandomNormal
    from test_gpu_zz_matching import _inbatch
    torch.manual_seed(41)
    B, L, n_users, n_items, K = 4096, 50, 1_000_000, 10_000, 20
    init = RandomNormal(0, 0.3)
    user = [F.SparseFeature("user_id", n_users, embed_dim=16, initializer=init), F.SequenceFeature("hist_item_id", n_items, embed_dim=16, pooling="mean", shared_with="item_id")]
    item = [F.SparseFeature("item_id", n_items, embed_dim=16, initializer=init)]
    model = M.DSSM(user, item, user_params={"dims": [256, 128, 64]}, item_params={"dims": [256, 128, 64]}).to(DEV).train()
    g = torch.Generator().manual_seed(42)
    lens = torch.randint(1, L + 1, (B,), generator=g)
    x = {"user_id": torch.randint(0, n_users, (B,), generator=g), "item_id": torch.randperm(n_items, generator=g)[:B],  # distinct items: no exact score ties
         "hist_item_id": torch.randint(1, n_items, (B, L), generator=g) * (torch.arange(L).unsqueeze(0) < lens.unsqueeze(1))}
    sd = _numpy_state(model)
    ue, ie, scores, neg, logits, loss = _inbatch(model, {k: v.to(DEV) for k, v in x.items()}, K)
    model.zero_grad()
    loss.backward()
    with orc.compact_table_grads():
        ref = orc.dssm_forward_backward(sd, {k: v.numpy() for k, v in x.items()}, _golden.DSSM_USER, _golden.DSSM_ITEM, 3, 3, K, train=True, neg_idx=neg.cpu().numpy())
    masked = ref["scores"].copy()
    np.fill_diagonal(masked, -np.inf)
    picked = np.take_along_axis(masked, neg.cpu().numpy(), axis=1)
    rest = masked.copy()
    np.put_along_axis(rest, neg.cpu().numpy(), -np.inf, axis=1)
    assert np.all(picked.min(axis=1) >= rest.max(axis=1) - 1e-5)  # the device's pick IS a hard-negative set
    assert np.abs(ue.detach().cpu().numpy() - ref["user_emb"]).max() < 5e-5 and np.abs(ie.detach().cpu().numpy() - ref["item_emb"]).max() < 5e-5
    assert np.abs(logits.detach().cpu().numpy() - ref["logits"]).max() < 1e-4
    assert abs(float(loss) - ref["loss"]) < 1e-4
    for k, p in model.named_parameters():
        want = ref["grads"][k]
        if isinstance(want, orc.CompactGrad):
            ids = torch.from_numpy(want.ids).to(DEV)
            got = p.grad.index_select(0, ids).cpu().numpy()
            scale = max(np.abs(want.rows).max(), 1e-6)
            assert np.abs(got - want.rows).max() <= 1e-3 * scale, (k, np.abs(got - want.rows).max(), scale)
            continue
        want = np.asarray(want).reshape(tuple(p.shape))
        scale = max(np.abs(want).max(), 1e-4)
        if k.endswith(".bias") and k[:-4] + "weight" in ref["grads"]:
            scale = max(scale, np.abs(np.asarray(ref["grads"][k[:-4] + "weight"])).max())
        got = p.grad.detach().cpu().numpy() if p.grad is not None else np.zeros_like(want)
        assert np.abs(got - want).max() <= 1e-3 * scale, (k, np.abs(got - want).max(), scale)
