"""Generate golden vectors from the LIVE, UNMODIFIED reference (build container only: needs /root/reference).

    python tests/golden/make_golden.py

Writes tests/golden/{deepfm_tutorial,deepfm_runcriteo,dcn,dcnv2,din,din_softmax}.npz (and dssm.npz, see dump_dssm): the reference model's
state_dict, the input batch, the labels, its pre-sigmoid logits / probabilities in train mode (dropout 0, BatchNorm
batch statistics) and in eval mode, and every parameter gradient of BCELoss(mean) in train mode — including the
dense (vocab, dim) table gradients the reference materialises.  Tables are initialised N(0, 0.05) instead of the
default N(0, 1e-4) so that logits are numerically non-trivial.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _live_reference as L  # noqa: E402


def dump(name, model, x, y, extra):
    model.train()
    p = model(x)
    loss = torch.nn.BCELoss()(p, y)
    model.zero_grad()
    loss.backward()
    rec = {"y": y.numpy(), "train_prob": p.detach().numpy(), "train_logit": torch.logit(p.detach().double()).numpy(), "loss": np.array(loss.item())}
    for k, v in x.items():
        rec["x." + k] = v.numpy()
    for k, v in model.state_dict().items():
        rec["sd." + k] = v.numpy()
    for k, prm in model.named_parameters():
        rec["grad." + k] = prm.grad.numpy()
    # eval AFTER recording the state_dict (train forward already updated the running stats that eval uses)
    model.eval()
    with torch.no_grad():
        pe = model(x)
    rec["eval_prob"] = pe.numpy()
    for k, v in model.state_dict().items():
        if "running_" in k or "num_batches" in k:
            rec["sd_after." + k] = v.numpy()
    for k, v in extra.items():
        rec["meta." + k] = np.array(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print(name, "loss", loss.item(), "logit range", float(rec["train_logit"].min()), float(rec["train_logit"].max()))


def main():
    L.load_reference()
    F = L.ref_module("basic.features")
    I = L.ref_module("basic.initializers")
    M = L.ref_module("models.ranking")
    init = I.RandomNormal(0, 0.05)
    B = 48
    g = torch.Generator().manual_seed(2022)

    def criteo(n_dense, n_sparse, dim, vocab):
        x = {"I%d" % i: torch.rand(B, generator=g) for i in range(n_dense)}
        x.update({"C%d" % i: torch.randint(0, vocab, (B,), generator=g) for i in range(n_sparse)})
        dense = [F.DenseFeature("I%d" % i) for i in range(n_dense)]
        sparse = [F.SparseFeature("C%d" % i, vocab_size=vocab, embed_dim=dim, initializer=init) for i in range(n_sparse)]
        return x, dense, sparse

    y = torch.randint(0, 2, (B,), generator=g).float()

    torch.manual_seed(1)
    x, dense, sparse = criteo(3, 5, 8, 23)
    dump("deepfm_tutorial", M.DeepFM(dense + sparse, sparse, {"dims": [16, 8], "dropout": 0.0, "activation": "relu"}), x, y, {"n_dense": 3, "n_sparse": 5, "n_hidden": 2})
    torch.manual_seed(2)
    x, dense, sparse = criteo(3, 5, 8, 23)
    dump("deepfm_runcriteo", M.DeepFM(dense, sparse, {"dims": [16, 8], "dropout": 0.0, "activation": "relu"}), x, y, {"n_dense": 3, "n_sparse": 5, "n_hidden": 2})
    torch.manual_seed(3)
    x, dense, sparse = criteo(3, 5, 8, 23)
    dump("dcn", M.DCN(dense + sparse, n_cross_layers=3, mlp_params={"dims": [16, 8]}), x, y, {"n_dense": 3, "n_sparse": 5, "n_hidden": 2, "n_cross": 3})
    torch.manual_seed(4)
    x, dense, sparse = criteo(3, 5, 8, 23)
    dump("dcnv2", M.DCNv2(dense + sparse, n_cross_layers=2, mlp_params={"dims": [16, 8], "dropout": 0.0, "activation": "relu"}, low_rank=4), x, y, {"n_dense": 3, "n_sparse": 5, "n_hidden": 2, "n_cross": 2})

    Lh = 6
    lens = torch.randint(1, Lh + 1, (B,), generator=g)
    pos = torch.arange(Lh).unsqueeze(0)
    xd = {
        "target_item_id": torch.randint(1, 30, (B,), generator=g),
        "target_cate_id": torch.randint(1, 7, (B,), generator=g),
        "user_id": torch.randint(1, 11, (B,), generator=g),
        "hist_item_id": torch.randint(1, 30, (B, Lh), generator=g) * (pos < lens.unsqueeze(1)),
        "hist_cate_id": torch.randint(1, 7, (B, Lh), generator=g) * (pos < lens.unsqueeze(1)),
    }
    for name, softmax, seed in (("din", False, 5), ("din_softmax", True, 6)):
        torch.manual_seed(seed)
        feats = [F.SparseFeature("target_item_id", 30, 8, initializer=init), F.SparseFeature("target_cate_id", 7, 8, initializer=init), F.SparseFeature("user_id", 11, 8, initializer=init)]
        hist = [F.SequenceFeature("hist_item_id", 30, 8, pooling="concat", shared_with="target_item_id"), F.SequenceFeature("hist_cate_id", 7, 8, pooling="concat", shared_with="target_cate_id")]
        dump(name, M.DIN(features=feats, history_features=hist, target_features=feats, mlp_params={"dims": [16, 8]}, attention_mlp_params={"dims": [16, 8], "use_softmax": softmax}), xd, y,
             {"n_hidden": 2, "n_att_hidden": 2, "use_softmax": int(softmax)})


def dump_dssm():
    """dssm.npz: the reference DSSM (mean-pooled history sharing the item table) + MatchTrainer's in-batch branch with HARD
    negatives (deterministic) and cross entropy: tower embeddings, point-wise probabilities, scores, sampled columns, logits,
    loss and every parameter gradient, train mode (BatchNorm batch statistics)."""
    F = L.ref_module("basic.features")
    I = L.ref_module("basic.initializers")
    M = L.ref_module("models.matching")
    U = L.ref_module("utils.match")
    init = I.RandomNormal(0, 0.3)
    B, Lh, n_users, n_items, K = 40, 6, 17, 29, 5
    g = torch.Generator().manual_seed(77)
    lens = torch.randint(1, Lh + 1, (B,), generator=g)
    x = {"user_id": torch.randint(0, n_users, (B,), generator=g), "item_id": torch.randint(0, n_items, (B,), generator=g),
         "hist_item_id": torch.randint(1, n_items, (B, Lh), generator=g) * (torch.arange(Lh).unsqueeze(0) < lens.unsqueeze(1))}
    torch.manual_seed(7)
    user = [F.SparseFeature("user_id", n_users, embed_dim=8, initializer=init), F.SequenceFeature("hist_item_id", n_items, embed_dim=8, pooling="mean", shared_with="item_id")]
    item = [F.SparseFeature("item_id", n_items, embed_dim=8, initializer=init)]
    model = M.DSSM(user, item, user_params={"dims": [16, 8], "activation": "relu"}, item_params={"dims": [16, 8], "activation": "relu"})
    model.train()
    rec = {}
    for k, v in x.items():
        rec["x." + k] = v.numpy()
    for k, v in model.state_dict().items():
        rec["sd." + k] = v.numpy()
    ue, ie = model.user_tower(x), model.item_tower(x)
    scores = ue @ ie.t()
    neg = U.inbatch_negative_sampling(scores, neg_ratio=K, hard_negative=True)
    logits = U.gather_inbatch_logits(scores, neg)
    loss = torch.nn.CrossEntropyLoss()(logits, torch.zeros(B, dtype=torch.long))
    model.zero_grad()
    loss.backward()
    rec.update({"user_emb": ue.detach().numpy(), "item_emb": ie.detach().numpy(), "prob": torch.sigmoid((ue * ie).sum(1)).detach().numpy(), "scores": scores.detach().numpy(),
                "neg_idx": neg.numpy(), "logits": logits.detach().numpy(), "loss": np.array(loss.item())})
    for k, prm in model.named_parameters():
        rec["grad." + k] = prm.grad.numpy()
    for k, v in {"neg_ratio": K, "n_users": n_users, "n_items": n_items, "n_hidden": 2}.items():
        rec["meta." + k] = np.array(v)
    np.savez_compressed(os.path.join(HERE, "dssm.npz"), **rec)
    gaps = np.sort(scores.detach().numpy() - 10 * np.eye(B), axis=1)
    print("dssm loss", loss.item(), "min gap between ranked scores", float(np.diff(gaps, axis=1)[:, -K - 1:].min()))


if __name__ == "__main__":
    main()
    dump_dssm()
