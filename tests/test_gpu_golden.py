"""GPU parity against the committed golden vectors (recorded from the live reference) and against the numpy oracle on
fresh seeded inputs — the -m gpu tests proper: CUDA kernels through the Python API / C ABI vs the checker."""
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


def _logit(p):
    p = p.astype(np.float64)
    return np.log(p) - np.log1p(-p)


@pytest.mark.parametrize("name", _golden.NAMES)
def test_cuda_path_matches_reference_golden(name):
    import torch_rechub.basic.features as F
    import torch_rechub.models.ranking as M
    rec = _golden.load(name)
    model = _golden.build_model(name, rec, F, M).to(DEV)
    x, y = _golden.torch_inputs(rec, DEV)
    model.eval()
    with torch.no_grad():
        pe = model(x).cpu().numpy()
    assert np.all(np.abs(_logit(pe) - _logit(rec["eval_prob"])) <= 1e-4 * np.abs(_logit(rec["eval_prob"])) + 1e-6)
    model.train()
    p = model(x)
    got = _logit(p.detach().cpu().numpy())
    ref = rec["train_logit"]
    assert np.all(np.abs(got - ref) <= 1e-4 * np.abs(ref) + 1e-6), np.abs(got - ref).max()
    torch.nn.BCELoss()(p, y).backward()
    for k, prm in model.named_parameters():
        g = rec["grad"][k]
        scale = max(np.abs(g).max(), 1e-3)
        if k.endswith(".bias") and k[:-4] + "weight" in rec["grad"]:
            scale = max(scale, np.abs(rec["grad"][k[:-4] + "weight"]).max())
        got_g = prm.grad.detach().cpu().numpy() if prm.grad is not None else np.zeros_like(g)
        assert np.abs(got_g - g).max() <= 2e-4 * scale, (k, np.abs(got_g - g).max(), scale)


@pytest.mark.parametrize("B,F,D,V", [(1, 3, 16, 50), (63, 26, 16, 1000), (513, 7, 8, 333), (2048, 26, 16, 100000)])
def test_deepfm_against_numpy_oracle(B, F, D, V):
    from torch_rechub.basic.features import DenseFeature, SparseFeature
    from torch_rechub.basic.initializers import RandomNormal
    from torch_rechub.models.ranking import DeepFM
    torch.manual_seed(B + F)
    dense = [DenseFeature("I%d" % i) for i in range(4)]
    sparse = [SparseFeature("C%d" % i, V, D, initializer=RandomNormal(0, 0.05)) for i in range(F)]
    model = DeepFM(dense + sparse, sparse, {"dims": [64, 32], "dropout": 0.0, "activation": "relu"}).to(DEV)
    g = torch.Generator().manual_seed(B)
    x = {"I%d" % i: torch.rand(B, generator=g) for i in range(4)}
    x.update({"C%d" % i: torch.randint(0, V, (B,), generator=g) for i in range(F)})
    y = torch.randint(0, 2, (B,), generator=g).float()
    train = B > 1
    model.train(train)
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    p = model({k: v.to(DEV) for k, v in x.items()})
    torch.nn.BCELoss()(p, y.to(DEV)).backward()
    ref = orc.deepfm_forward_backward(sd, {k: v.numpy() for k, v in x.items()}, y.numpy(), [f.name for f in dense], [f.name for f in sparse], [f.name for f in sparse], 2, train=train)
    got = _logit(p.detach().cpu().numpy())
    assert np.all(np.abs(got - ref["logit"]) <= 1e-4 * np.abs(ref["logit"]) + 1e-6), np.abs(got - ref["logit"]).max()
    for k, prm in model.named_parameters():
        r = np.asarray(ref["grads"][k]).reshape(prm.shape)
        scale = max(np.abs(r).max(), 1e-4)
        if k.endswith(".bias") and k[:-4] + "weight" in ref["grads"]:
            scale = max(scale, np.abs(ref["grads"][k[:-4] + "weight"]).max())
        assert np.abs(prm.grad.detach().cpu().numpy() - r).max() <= 2e-4 * scale, k


def test_dcn_and_din_against_numpy_oracle():
    from torch_rechub.basic.features import DenseFeature, SequenceFeature, SparseFeature
    from torch_rechub.basic.initializers import RandomNormal
    from torch_rechub.models.ranking import DCN, DIN
    init = RandomNormal(0, 0.05)
    torch.manual_seed(9)
    B = 300
    dense = [DenseFeature("I%d" % i) for i in range(13)]
    sparse = [SparseFeature("C%d" % i, 500, 16, initializer=init) for i in range(26)]
    model = DCN(dense + sparse, n_cross_layers=3, mlp_params={"dims": [64, 32]}).to(DEV).train()
    g = torch.Generator().manual_seed(1)
    x = {"I%d" % i: torch.rand(B, generator=g) for i in range(13)}
    x.update({"C%d" % i: torch.randint(0, 500, (B,), generator=g) for i in range(26)})
    y = torch.randint(0, 2, (B,), generator=g).float()
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    p = model({k: v.to(DEV) for k, v in x.items()})
    torch.nn.BCELoss()(p, y.to(DEV)).backward()
    ref = orc.dcn_forward_backward(sd, {k: v.numpy() for k, v in x.items()}, y.numpy(), [f.name for f in dense], [f.name for f in sparse], 3, 2)
    got = _logit(p.detach().cpu().numpy())
    assert np.all(np.abs(got - ref["logit"]) <= 1e-4 * np.abs(ref["logit"]) + 1e-6)
    for k in ("cn.w.0.weight", "cn.b.2", "embedding.embed_dict.C5.weight", "linear.fc.weight"):
        r = np.asarray(ref["grads"][k])
        assert np.abs(dict(model.named_parameters())[k].grad.cpu().numpy().reshape(r.shape) - r).max() <= 2e-4 * max(np.abs(r).max(), 1e-4), k

    # DIN, Amazon-Electronics shape scaled down (L = 50, D = 8)
    torch.manual_seed(10)
    B, L = 128, 50
    feats = [SparseFeature("target_item_id", 2000, 8, initializer=init), SparseFeature("target_cate_id", 60, 8, initializer=init), SparseFeature("user_id", 300, 8, initializer=init)]
    hist = [SequenceFeature("hist_item_id", 2000, 8, pooling="concat", shared_with="target_item_id"), SequenceFeature("hist_cate_id", 60, 8, pooling="concat", shared_with="target_cate_id")]
    din = DIN(features=feats, history_features=hist, target_features=feats, mlp_params={"dims": [64, 32]}, attention_mlp_params={"dims": [64, 32]}).to(DEV).train()
    lens = torch.randint(1, L + 1, (B,), generator=g)
    pos = torch.arange(L).unsqueeze(0)
    xd = {
        "target_item_id": torch.randint(1, 2000, (B,), generator=g),
        "target_cate_id": torch.randint(1, 60, (B,), generator=g),
        "user_id": torch.randint(1, 300, (B,), generator=g),
        "hist_item_id": torch.randint(1, 2000, (B, L), generator=g) * (pos < lens.unsqueeze(1)),
        "hist_cate_id": torch.randint(1, 60, (B, L), generator=g) * (pos < lens.unsqueeze(1)),
    }
    y = torch.randint(0, 2, (B,), generator=g).float()
    sd = {k: v.detach().cpu().numpy() for k, v in din.state_dict().items()}
    p = din({k: v.to(DEV) for k, v in xd.items()})
    torch.nn.BCELoss()(p, y.to(DEV)).backward()
    ref = orc.din_forward_backward(sd, {k: v.numpy() for k, v in xd.items()}, y.numpy(), _golden.DIN_FEATURES, _golden.DIN_HISTORY, _golden.DIN_FEATURES, _golden.DIN_SHARED, 2, 2)
    got = _logit(p.detach().cpu().numpy())
    assert np.all(np.abs(got - ref["logit"]) <= 1e-4 * np.abs(ref["logit"]) + 1e-6), np.abs(got - ref["logit"]).max()
    for k in ("embedding.embed_dict.target_item_id.weight", "embedding.embed_dict.target_cate_id.weight", "attention_layers.0.attention.mlp.0.weight", "attention_layers.1.attention.mlp.6.alpha", "mlp.mlp.4.weight"):
        r = np.asarray(ref["grads"][k])
        gg = dict(din.named_parameters())[k].grad.cpu().numpy().reshape(r.shape)
        assert np.abs(gg - r).max() <= 3e-4 * max(np.abs(r).max(), 1e-4), (k, np.abs(gg - r).max(), np.abs(r).max())
