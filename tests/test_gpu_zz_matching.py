"""GPU parity of the two-tower path: DSSM towers through the fused gather (+ mean-pooled history sharing
the item table) and the tower kernels, MatchTrainer's in-batch branch with the batched CUDA sampler — against the golden
vectors recorded from the live reference and against the numpy oracle on fresh inputs.  (Sorted last on purpose: it was
written after the last GPU session of its round.)"""
import os
import sys

import numpy as np
import pytest
import torch

import _golden

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import rechub_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _inbatch(model, x, k):
    from torch_rechub.utils.match import gather_inbatch_logits, inbatch_negative_sampling
    ue, ie = model.user_tower(x), model.item_tower(x)
    scores = ue @ ie.t()
    neg = inbatch_negative_sampling(scores, neg_ratio=k, hard_negative=True)
    logits = gather_inbatch_logits(scores, neg)
    loss = torch.nn.CrossEntropyLoss()(logits, torch.zeros(logits.size(0), dtype=torch.long, device=logits.device))
    return ue, ie, scores, neg, logits, loss


def _as_out(ue, ie, logits, loss):
    return {"user_emb": ue.detach().cpu().numpy(), "item_emb": ie.detach().cpu().numpy(), "prob": torch.sigmoid((ue * ie).sum(1)).detach().cpu().numpy(),
            "logits": logits.detach().cpu().numpy(), "loss": float(loss)}


def test_dssm_cuda_matches_reference_golden():
    import torch_rechub.basic.features as F
    import torch_rechub.models.matching as M
    rec = _golden.load("dssm")
    model = _golden.build_dssm(rec, F, M).to(DEV).train()
    x = {k: torch.from_numpy(v).to(DEV) for k, v in rec["x"].items()}
    ue, ie, scores, neg, logits, loss = _inbatch(model, x, rec["meta"]["neg_ratio"])
    model.zero_grad()
    loss.backward()
    grads = {k: (p.grad.detach().cpu().numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)) for k, p in model.named_parameters()}
    _golden.check_dssm_against_golden(rec, _as_out(ue, ie, logits, loss), grads, emb_tol=2e-5, grad_rtol=2e-4)
    assert not bool((neg == torch.arange(neg.size(0), device=DEV).unsqueeze(1)).any())


def test_dssm_cuda_against_numpy_oracle_on_a_larger_batch():
    import torch_rechub.basic.features as F
    import torch_rechub.models.matching as M
    from torch_rechub.basic.initializers import RandomNormal
    torch.manual_seed(31)
    B, L, n_users, n_items, K = 512, 10, 3000, 5000, 7
    init = RandomNormal(0, 0.3)
    user = [F.SparseFeature("user_id", n_users, embed_dim=16, initializer=init), F.SequenceFeature("hist_item_id", n_items, embed_dim=16, pooling="mean", shared_with="item_id")]
    item = [F.SparseFeature("item_id", n_items, embed_dim=16, initializer=init), F.SparseFeature("cate_id", 40, embed_dim=16, initializer=init)]
    model = M.DSSM(user, item, user_params={"dims": [64, 32]}, item_params={"dims": [64, 32]}).to(DEV).train()  # both towers: 32 -> 64 -> 32 (tensor-core GEMMs)
    g = torch.Generator().manual_seed(4)
    lens = torch.randint(1, L + 1, (B,), generator=g)
    x = {"user_id": torch.randint(0, n_users, (B,), generator=g), "item_id": torch.randperm(n_items, generator=g)[:B],  # distinct items: no exact score ties
         "cate_id": torch.randint(0, 40, (B,), generator=g),
         "hist_item_id": torch.randint(1, n_items, (B, L), generator=g) * (torch.arange(L).unsqueeze(0) < lens.unsqueeze(1))}
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    item_spec = _golden.DSSM_ITEM + [("sparse", "cate_id", "cate_id")]
    ue, ie, scores, neg, logits, loss = _inbatch(model, {k: v.to(DEV) for k, v in x.items()}, K)
    model.zero_grad()
    loss.backward()
    # the oracle scores the SAME columns the device picked (two scores within fp32 rounding of each other may rank either way) ...
    ref = orc.dssm_forward_backward(sd, {k: v.numpy() for k, v in x.items()}, _golden.DSSM_USER, item_spec, 2, 2, K, train=True, neg_idx=neg.cpu().numpy())
    # ... after checking that the pick IS a hard-negative set: nothing left out beats anything picked by more than rounding
    masked = ref["scores"].copy()
    np.fill_diagonal(masked, -np.inf)
    picked = np.take_along_axis(masked, neg.cpu().numpy(), axis=1)
    rest = masked.copy()
    np.put_along_axis(rest, neg.cpu().numpy(), -np.inf, axis=1)
    assert np.all(picked.min(axis=1) >= rest.max(axis=1) - 1e-5)
    assert np.abs(ue.detach().cpu().numpy() - ref["user_emb"]).max() < 5e-5 and np.abs(ie.detach().cpu().numpy() - ref["item_emb"]).max() < 5e-5
    assert np.abs(logits.detach().cpu().numpy() - ref["logits"]).max() < 1e-4
    assert abs(float(loss) - ref["loss"]) < 1e-4
    for k, p in model.named_parameters():
        want = np.asarray(ref["grads"][k]).reshape(tuple(p.shape))
        scale = max(np.abs(want).max(), 1e-4)
        if k.endswith(".bias") and k[:-4] + "weight" in ref["grads"]:
            scale = max(scale, np.abs(ref["grads"][k[:-4] + "weight"]).max())
        got = p.grad.detach().cpu().numpy() if p.grad is not None else np.zeros_like(want)
        assert np.abs(got - want).max() <= 1e-3 * scale, (k, np.abs(got - want).max(), scale)


def test_cuda_inbatch_sampler_properties():
    from torch_rechub.utils.match import inbatch_negative_sampling
    n, k = 257, 19
    scores = torch.randn(n, n, device=DEV)
    diag = torch.arange(n, device=DEV).unsqueeze(1)
    gen = torch.Generator(device=DEV)
    gen.manual_seed(3)
    a = inbatch_negative_sampling(scores, neg_ratio=k, generator=gen)
    assert a.shape == (n, k) and a.dtype == torch.long and a.device.type == "cuda"
    assert int(a.min()) >= 0 and int(a.max()) < n and not bool((a == diag).any())
    assert all(len(set(row)) == k for row in a.cpu().tolist())  # without replacement
    gen.manual_seed(3)
    assert torch.equal(a, inbatch_negative_sampling(scores, neg_ratio=k, generator=gen))
    gen.manual_seed(4)
    assert not torch.equal(a, inbatch_negative_sampling(scores, neg_ratio=k, generator=gen))
    full = inbatch_negative_sampling(scores)  # default: every other column
    assert full.shape == (n, n - 1) and torch.equal(full.sort(dim=1).values, torch.arange(n, device=DEV).expand(n, n)[diag != torch.arange(n, device=DEV)].view(n, n - 1))
    counts = torch.bincount(inbatch_negative_sampling(torch.zeros(64, 64, device=DEV), neg_ratio=16).flatten(), minlength=64)
    assert int(counts.min()) > 0  # 63 rows each take 16 of a column's 63 chances: P(some column never drawn) ~ 64 * (47/63)^63 ~ 6e-7
    hard = inbatch_negative_sampling(scores, neg_ratio=k, hard_negative=True)
    assert torch.equal(hard.cpu(), inbatch_negative_sampling(scores.cpu(), neg_ratio=k, hard_negative=True))


def test_match_trainer_inbatch_epoch_on_cuda():
    import torch_rechub.basic.features as F
    import torch_rechub.models.matching as M
    from torch_rechub.trainers import MatchTrainer
    torch.manual_seed(5)
    user = [F.SparseFeature("user_id", 50, embed_dim=8), F.SequenceFeature("hist_item_id", 80, embed_dim=8, pooling="mean", shared_with="item_id")]
    item = [F.SparseFeature("item_id", 80, embed_dim=8)]
    model = M.DSSM(user, item, user_params={"dims": [16]}, item_params={"dims": [16]})
    g = torch.Generator().manual_seed(6)
    data = [({"user_id": torch.randint(0, 50, (32,), generator=g), "item_id": torch.randint(0, 80, (32,), generator=g), "hist_item_id": torch.randint(0, 80, (32, 5), generator=g)}, torch.zeros(32))
            for _ in range(6)]
    before = {k: v.detach().clone() for k, v in model.state_dict().items()}
    t = MatchTrainer(model, mode=0, in_batch_neg=True, in_batch_neg_ratio=5, sampler_seed=1, n_epoch=1, device=DEV)
    loss = t.train_one_epoch(data, log_interval=2)
    assert np.isfinite(loss) and 0.0 < loss < 10.0
    after = model.state_dict()
    assert any(not torch.equal(before[k].to(DEV), after[k]) for k in before if k.endswith("weight"))
    emb = model.user_tower({k: v.to(DEV) for k, v in data[0][0].items()})
    assert torch.allclose(emb.norm(dim=1), torch.ones(32, device=DEV), atol=1e-5)


@pytest.mark.parametrize("hard,K", [(False, 20), (True, 20), (False, None), (True, 5)])
def test_fused_inbatch_loss_equals_the_composition(hard, K):
    """MatchTrainer's in-batch branch on the engine's kernels (one sampling launch, cross entropy + backward from the tower outputs,
    no dense (B, B) score gradient) against the reference-order composition (matmul -> gather_inbatch_logits -> CrossEntropyLoss)
    evaluated on the SAME picks: loss and every parameter gradient."""
    import copy
    import torch_rechub.basic.features as F
    import torch_rechub.models.matching as M
    from torch_rechub.basic.initializers import RandomNormal
    from torch_rechub.trainers import MatchTrainer
    from torch_rechub.utils.match import gather_inbatch_logits
    torch.manual_seed(11)
    B, L, n_users, n_items = 512, 8, 2000, 3000
    init = RandomNormal(0, 0.3)
    user = [F.SparseFeature("user_id", n_users, embed_dim=16, initializer=init), F.SequenceFeature("hist_item_id", n_items, embed_dim=16, pooling="mean", shared_with="item_id")]
    item = [F.SparseFeature("item_id", n_items, embed_dim=16, initializer=init)]
    model = M.DSSM(user, item, user_params={"dims": [64, 32]}, item_params={"dims": [64, 32]})
    ref = copy.deepcopy(model).to(DEV).train()
    g = torch.Generator().manual_seed(4)
    x = {"user_id": torch.randint(0, n_users, (B,), generator=g).to(DEV), "item_id": torch.randperm(n_items, generator=g)[:B].to(DEV), "hist_item_id": torch.randint(0, n_items, (B, L), generator=g).to(DEV)}
    t = MatchTrainer(model, mode=0, in_batch_neg=True, in_batch_neg_ratio=K, hard_negative=hard, sampler_seed=7, n_epoch=1, device=DEV)
    t.model.train()
    loss = t._inbatch_loss(x)
    picks = t.last_picks
    kk = K if K is not None else B - 1
    assert picks.shape == (B, kk) and picks.dtype == torch.int64
    diag = torch.arange(B, device=DEV).unsqueeze(1)
    assert int(picks.min()) >= 0 and int(picks.max()) < B and not bool((picks == diag).any())
    assert all(len(set(r)) == kk for r in picks[:64].cpu().tolist())  # without replacement
    t.model.zero_grad()
    loss.backward()
    ue, ie = ref.user_tower(x), ref.item_tower(x)
    scores = ue @ ie.t()
    if hard:  # the picks are a hard-negative set of the reference scores (ties within rounding may order either way)
        masked = scores.detach().clone()
        masked.fill_diagonal_(float("-inf"))
        picked = torch.gather(masked, 1, picks)
        rest = masked.clone()
        rest.scatter_(1, picks, float("-inf"))
        assert bool((picked.min(dim=1).values >= rest.max(dim=1).values - 1e-5).all())
    logits = gather_inbatch_logits(scores, picks)
    want = torch.nn.CrossEntropyLoss()(logits, torch.zeros(B, dtype=torch.long, device=DEV))
    assert abs(float(loss) - float(want)) <= 2e-5 * abs(float(want)) + 1e-6
    want.backward()
    for (n, p), q in zip(t.model.named_parameters(), ref.parameters()):
        if q.grad is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, n
            continue
        scale = max(float(q.grad.abs().max()), 1e-6)
        if n.endswith(".bias"):
            scale = max(scale, 1e-4)
        assert float((p.grad - q.grad).abs().max()) <= 1e-3 * scale, (n, float((p.grad - q.grad).abs().max()), scale)
    # the sampler replays under the same seed and moves under another
    t2 = MatchTrainer(copy.deepcopy(model), mode=0, in_batch_neg=True, in_batch_neg_ratio=K, hard_negative=hard, sampler_seed=7, n_epoch=1, device=DEV)
    t2.model.train()
    t2._inbatch_loss(x)
    if not hard and K is not None:
        t3 = MatchTrainer(copy.deepcopy(model), mode=0, in_batch_neg=True, in_batch_neg_ratio=K, hard_negative=hard, sampler_seed=8, n_epoch=1, device=DEV)
        t3.model.train()
        t3._inbatch_loss(x)
        assert not torch.equal(t3.last_picks, picks)
        counts = torch.bincount(picks.flatten(), minlength=B)
        assert int(counts.min()) > 0 and int(counts.max()) < 4 * kk  # every column is drawn, none dominates
