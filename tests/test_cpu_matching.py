"""Two-tower / in-batch-negative path on CPU: bit-for-bit against the live reference for the same seeds —
sample generation, negative samplers, DSSM forward, MatchTrainer steps — and the reference's own matching tests run unmodified
against this package."""
import copy
import os
import random
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest
import torch

import _live_reference as live

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "torch-rechub_b200")
needs_ref = pytest.mark.skipif(not live.live_reference_available(), reason="live reference only exists in the build container")


def _events(n_users=14, n_items=30, n=160, seed=3):
    g = np.random.RandomState(seed)
    return pd.DataFrame({"user_id": g.randint(0, n_users, n), "item_id": g.randint(0, n_items, n), "cate": g.randint(0, 5, n), "time": np.arange(n)})


def _frames_equal(a, b):
    assert list(a.columns) == list(b.columns) and len(a) == len(b)
    for c in a.columns:
        for u, v in zip(a[c].tolist(), b[c].tolist()):
            assert np.array_equal(np.asarray(u), np.asarray(v)), c


@needs_ref
@pytest.mark.parametrize("mode,neg_ratio,method", [(0, 2, 0), (1, 0, 1), (2, 3, 2)])
def test_sample_generation_matches_reference(mode, neg_ratio, method):
    from torch_rechub.utils import match as mine
    theirs = live.ref_module("utils.match")
    out = []
    for mod in (mine, theirs):
        np.random.seed(11)
        random.seed(12)
        out.append(mod.generate_seq_feature_match(_events(), "user_id", "item_id", "time", item_attribute_cols=["cate"], sample_method=method, mode=mode, neg_ratio=neg_ratio))
    _frames_equal(out[0][0], out[1][0])
    _frames_equal(out[0][1], out[1][1])
    users = pd.DataFrame({"user_id": np.arange(14)})
    items = pd.DataFrame({"item_id": np.arange(30)})
    xa = mine.gen_model_input(out[0][0], users, "user_id", items, "item_id", seq_max_len=6)
    xb = theirs.gen_model_input(out[1][0], users, "user_id", items, "item_id", seq_max_len=6)
    assert list(xa.keys()) == list(xb.keys())
    for k in xa:
        assert np.array_equal(np.asarray(xa[k]), np.asarray(xb[k])), k


@needs_ref
def test_negative_sample_methods_match_reference():
    from torch_rechub.utils import match as mine
    theirs = live.ref_module("utils.match")
    counts = {i: c for i, c in zip(range(20), sorted(np.random.RandomState(1).randint(1, 50, 20).tolist(), reverse=True))}
    for method, ratio in ((0, 64), (1, 64), (2, 64), (3, 15)):  # method 3 draws without replacement: ratio <= number of items
        np.random.seed(4)
        a = mine.negative_sample(counts, ratio, method_id=method)
        np.random.seed(4)
        b = theirs.negative_sample(counts, ratio, method_id=method)
        assert np.array_equal(a, b), method
    with pytest.raises(ValueError):
        mine.negative_sample(counts, 3, method_id=7)


@needs_ref
def test_inbatch_sampler_and_losses_match_reference():
    from torch_rechub.basic.loss_func import BPRLoss
    from torch_rechub.utils import match as mine
    theirs = live.ref_module("utils.match")
    scores = torch.randn(9, 9, generator=torch.Generator().manual_seed(0))
    for ratio in (None, 3, 8, 50):
        a = mine.inbatch_negative_sampling(scores, neg_ratio=ratio, generator=torch.Generator().manual_seed(5))
        b = theirs.inbatch_negative_sampling(scores, neg_ratio=ratio, generator=torch.Generator().manual_seed(5))
        assert torch.equal(a, b)
        assert torch.equal(mine.gather_inbatch_logits(scores, a), theirs.gather_inbatch_logits(scores, b))
    a = mine.inbatch_negative_sampling(scores, neg_ratio=4, hard_negative=True)
    b = theirs.inbatch_negative_sampling(scores, neg_ratio=4, hard_negative=True)
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        mine.inbatch_negative_sampling(torch.zeros(3))
    with pytest.raises(ValueError):
        mine.inbatch_negative_sampling(torch.zeros(1, 1))
    ref_bpr = live.ref_module("basic.loss_func").BPRLoss()
    pos, neg1, negk = torch.randn(7), torch.randn(7), torch.randn(7, 4)
    assert torch.equal(BPRLoss()(pos, neg1), ref_bpr(pos, neg1))
    assert torch.equal(BPRLoss()(pos, negk, in_batch_neg=True), ref_bpr(pos, negk, in_batch_neg=True))


def _two_tower(F, M, n_users=14, n_items=30):
    torch.manual_seed(21)
    user = [F.SparseFeature("user_id", n_users, embed_dim=8), F.SequenceFeature("hist_item_id", n_items, embed_dim=8, pooling="mean", shared_with="item_id")]
    item = [F.SparseFeature("item_id", n_items, embed_dim=8)]
    return M.DSSM(user, item, user_params={"dims": [16, 8]}, item_params={"dims": [16, 8]})


def _batches(n_batches=5, b=12, n_users=14, n_items=30, L=6):
    g = torch.Generator().manual_seed(8)
    out = []
    for _ in range(n_batches):
        x = {"user_id": torch.randint(0, n_users, (b,), generator=g), "item_id": torch.randint(0, n_items, (b,), generator=g), "hist_item_id": torch.randint(0, n_items, (b, L), generator=g)}
        out.append((x, torch.randint(0, 2, (b,), generator=g)))
    return out


@needs_ref
@pytest.mark.parametrize("kw", [dict(mode=0, in_batch_neg=True, in_batch_neg_ratio=4, sampler_seed=2), dict(mode=1, in_batch_neg=True, hard_negative=True), dict(mode=2, in_batch_neg=True, sampler_seed=9),
                                dict(mode=0)])
def test_dssm_and_match_trainer_match_reference(kw):
    import torch_rechub.basic.features as F
    import torch_rechub.models.matching as M
    from torch_rechub.trainers import MatchTrainer
    rF, rM, rT = live.ref_module("basic.features"), live.ref_module("models.matching"), live.ref_module("trainers").MatchTrainer
    mine, theirs = _two_tower(F, M), _two_tower(rF, rM)
    assert list(mine.state_dict().keys()) == list(theirs.state_dict().keys())
    for (k, a), b in zip(mine.state_dict().items(), theirs.state_dict().values()):
        assert torch.equal(a, b), k
    data = _batches()
    mine.eval(), theirs.eval()
    assert torch.equal(mine(data[0][0]), theirs(data[0][0]))
    for mode in ("user", "item"):
        mine.mode = theirs.mode = mode
        assert torch.equal(mine(data[0][0]), theirs(data[0][0]))
    mine.mode = theirs.mode = None
    la = MatchTrainer(mine, n_epoch=1, device="cpu", **kw).train_one_epoch(data, log_interval=2)
    lb = rT(theirs, n_epoch=1, device="cpu", **kw).train_one_epoch(data, log_interval=2)
    assert la == lb
    for (k, a), b in zip(mine.state_dict().items(), theirs.state_dict().values()):
        assert torch.equal(a, b), k


def _variant(F, M, kind, n_users=14, n_items=30):
    torch.manual_seed(33)
    user = [F.SparseFeature("user_id", n_users, embed_dim=8), F.SequenceFeature("hist_item_id", n_items, embed_dim=8, pooling="mean", shared_with="item_id")]
    item = [F.SparseFeature("item_id", n_items, embed_dim=8)]
    if kind == "youtube":
        neg = [F.SequenceFeature("neg_items", n_items, embed_dim=8, pooling="concat", shared_with="item_id")]
        return M.YoutubeDNN(user, item, neg, user_params={"dims": [16, 8]}, temperature=0.5)
    neg = [F.SparseFeature("neg_items", n_items, embed_dim=8, shared_with="item_id")]
    return M.FaceBookDSSM(user, item, neg, user_params={"dims": [16, 8]}, item_params={"dims": [8]})


def _variant_batches(kind, n_batches=4, b=10, n_users=14, n_items=30, L=5, n_neg=3):
    g = torch.Generator().manual_seed(18)
    out = []
    for _ in range(n_batches):
        x = {"user_id": torch.randint(0, n_users, (b,), generator=g), "item_id": torch.randint(0, n_items, (b,), generator=g), "hist_item_id": torch.randint(0, n_items, (b, L), generator=g),
             "neg_items": torch.randint(0, n_items, (b, n_neg) if kind == "youtube" else (b,), generator=g)}
        out.append((x, torch.zeros(b, dtype=torch.long)))
    return out


@needs_ref
@pytest.mark.parametrize("kind,mode", [("youtube", 2), ("facebook", 1)])
def test_listwise_and_pairwise_variants_match_reference(kind, mode):
    """YoutubeDNN (list-wise softmax) and FaceBookDSSM (pair-wise BPR): same weights, outputs, tower modes and MatchTrainer
    epoch as the live reference — the trainer's modes 2 and 1 without in-batch negatives."""
    import torch_rechub.basic.features as F
    import torch_rechub.models.matching as M
    from torch_rechub.trainers import MatchTrainer
    rF, rM, rT = live.ref_module("basic.features"), live.ref_module("models.matching"), live.ref_module("trainers").MatchTrainer
    mine, theirs = _variant(F, M, kind), _variant(rF, rM, kind)
    assert list(mine.state_dict().keys()) == list(theirs.state_dict().keys())
    for (k, a), b in zip(mine.state_dict().items(), theirs.state_dict().values()):
        assert torch.equal(a, b), k
    data = _variant_batches(kind)
    mine.eval(), theirs.eval()
    ya, yb = mine(data[0][0]), theirs(data[0][0])
    if kind == "youtube":
        assert ya.shape == (10, 4) and torch.equal(ya, yb)
    else:
        assert torch.equal(ya[0], yb[0]) and torch.equal(ya[1], yb[1])
    for tower in ("user", "item"):
        mine.mode = theirs.mode = tower
        assert torch.equal(mine(data[0][0]), theirs(data[0][0]))
    mine.mode = theirs.mode = None
    la = MatchTrainer(mine, mode=mode, n_epoch=1, device="cpu").train_one_epoch(data, log_interval=2)
    lb = rT(theirs, mode=mode, n_epoch=1, device="cpu").train_one_epoch(data, log_interval=2)
    assert la == lb
    for (k, a), b in zip(mine.state_dict().items(), theirs.state_dict().values()):
        assert torch.equal(a, b), k


def test_match_trainer_rejects_models_without_towers_and_bad_modes():
    from torch_rechub.trainers import MatchTrainer
    with pytest.raises(ValueError):
        MatchTrainer(torch.nn.Linear(2, 1), in_batch_neg=True)
    with pytest.raises(ValueError):
        MatchTrainer(torch.nn.Linear(2, 1), mode=3)
    import torch_rechub.models.matching as M
    with pytest.raises(NotImplementedError):
        M.MIND()


def test_fit_saves_and_inference_embedding_roundtrip(tmp_path):
    import torch_rechub.basic.features as F
    import torch_rechub.models.matching as M
    from torch_rechub.trainers import MatchTrainer
    model = _two_tower(F, M)
    data = _batches()
    t = MatchTrainer(model, mode=0, n_epoch=1, device="cpu", model_path=str(tmp_path))
    t.fit([(x, y.float()) for x, y in data], val_dataloader=[(x, y) for x, y in data])
    assert (tmp_path / "model.pth").exists()
    users = t.inference_embedding(copy.deepcopy(model), "user", [x for x, _ in data], str(tmp_path))
    items = t.inference_embedding(copy.deepcopy(model), "item", [x for x, _ in data], str(tmp_path))
    assert users.shape == (60, 8) and items.shape == (60, 8)
    assert torch.allclose(users.norm(dim=1), torch.ones(60), atol=1e-5)
    preds = t.predict(model, data)
    assert len(preds) == 60 and all(0.0 <= p <= 1.0 for p in preds)


@needs_ref
def test_reference_matching_tests_pass_against_this_package(tmp_path):
    """The reference's own in-batch sampling tests and the DSSM case of its matching e2e test, byte-for-byte, against THIS package."""
    import shutil
    tdir = tmp_path / "suite" / "tests"
    tdir.mkdir(parents=True)
    for name in ("test_inbatch_sampling.py", "test_e2e_matching.py"):
        shutil.copy(os.path.join(live.REFERENCE_ROOT, "tests", name), tdir / name)
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", str(tdir), "-k", "inbatch or DSSM or YoutubeDNN"]
    res = subprocess.run(cmd, env=dict(os.environ, PYTHONPATH=PKG), cwd=str(tdir), capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert "passed" in res.stdout


def test_package_cpu_route_matches_dssm_golden():
    """Needs no live reference (runs on the GPU box too): this package's DSSM + in-batch hard negatives vs tests/golden/dssm.npz."""
    import _golden
    import torch_rechub.basic.features as F
    import torch_rechub.models.matching as M
    from torch_rechub.utils.match import gather_inbatch_logits, inbatch_negative_sampling
    rec = _golden.load("dssm")
    model = _golden.build_dssm(rec, F, M).train()
    x = {k: torch.from_numpy(v) for k, v in rec["x"].items()}
    ue, ie = model.user_tower(x), model.item_tower(x)
    scores = ue @ ie.t()
    neg = inbatch_negative_sampling(scores, neg_ratio=rec["meta"]["neg_ratio"], hard_negative=True)
    logits = gather_inbatch_logits(scores, neg)
    loss = torch.nn.CrossEntropyLoss()(logits, torch.zeros(logits.size(0), dtype=torch.long))
    model.zero_grad()
    loss.backward()
    out = {"user_emb": ue.detach().numpy(), "item_emb": ie.detach().numpy(), "prob": torch.sigmoid((ue * ie).sum(1)).detach().numpy(), "logits": logits.detach().numpy(), "loss": float(loss)}
    _golden.check_dssm_against_golden(rec, out, {k: p.grad.numpy() for k, p in model.named_parameters()}, emb_tol=1e-7, grad_rtol=1e-6)
    assert np.array_equal(neg.numpy(), rec["neg_idx"])


def test_batched_sampler_properties_on_cpu_tensors():
    """The CUDA route's batched draw, exercised here on CPU tensors: shape, range, no positive, no repeats, seed behaviour,
    uniform coverage, and agreement of the batched hard-negative pick with the row-wise one."""
    from torch_rechub.utils.match import _sample_batched, _sample_rowwise
    n, k = 101, 13
    scores = torch.randn(n, n, generator=torch.Generator().manual_seed(1))
    diag = torch.arange(n).unsqueeze(1)
    a = _sample_batched(scores, k, False, torch.Generator().manual_seed(3))
    assert a.shape == (n, k) and a.dtype == torch.long
    assert int(a.min()) >= 0 and int(a.max()) < n and not bool((a == diag).any())
    assert all(len(set(row)) == k for row in a.tolist())
    assert torch.equal(a, _sample_batched(scores, k, False, torch.Generator().manual_seed(3)))
    assert not torch.equal(a, _sample_batched(scores, k, False, torch.Generator().manual_seed(4)))
    full = _sample_batched(scores, n - 1, False, None)
    others = torch.arange(n).expand(n, n)[diag != torch.arange(n)].view(n, n - 1)
    assert torch.equal(full.sort(dim=1).values, others)
    counts = torch.bincount(_sample_batched(torch.zeros(64, 64), 8, False, torch.Generator().manual_seed(0)).flatten(), minlength=64)
    assert int(counts.min()) > 0 and int(counts.max()) < 30  # 512 uniform draws over 64 columns: mean 8
    assert torch.equal(_sample_batched(scores, k, True, None), _sample_rowwise(scores, k, True, None))
