"""Helpers to load tests/golden/*.npz (written by tests/golden/make_golden.py from the live reference)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["deepfm_tutorial", "deepfm_runcriteo", "dcn", "dcnv2", "din", "din_softmax"]


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    rec = {"sd": {}, "x": {}, "grad": {}, "meta": {}, "sd_after": {}}
    for k in z.files:
        head, _, tail = k.partition(".")
        if head in rec and tail:
            rec[head][tail] = z[k]
        else:
            rec[k] = z[k]
    rec["meta"] = {k: int(v) for k, v in rec["meta"].items()}
    return rec


def names_of(rec):
    dense = sorted([k for k in rec["x"] if k.startswith("I")], key=lambda s: int(s[1:]))
    sparse = sorted([k for k in rec["x"] if k.startswith("C")], key=lambda s: int(s[1:]))
    return dense, sparse


DIN_FEATURES = ["target_item_id", "target_cate_id", "user_id"]
DIN_HISTORY = ["hist_item_id", "hist_cate_id"]
DIN_SHARED = {"hist_item_id": "target_item_id", "hist_cate_id": "target_cate_id"}


def oracle_run(orc, name, rec, train=True, backward=True):
    m = rec["meta"]
    dense, sparse = names_of(rec)
    if name == "deepfm_tutorial":
        return orc.deepfm_forward_backward(rec["sd"], rec["x"], rec["y"], dense, sparse, sparse, m["n_hidden"], train=train, backward=backward)
    if name == "deepfm_runcriteo":
        return orc.deepfm_forward_backward(rec["sd"], rec["x"], rec["y"], dense, [], sparse, m["n_hidden"], train=train, backward=backward)
    if name == "dcn":
        return orc.dcn_forward_backward(rec["sd"], rec["x"], rec["y"], dense, sparse, m["n_cross"], m["n_hidden"], train=train, backward=backward)
    if name == "dcnv2":
        return orc.dcnv2_forward_backward(rec["sd"], rec["x"], rec["y"], dense, sparse, m["n_cross"], m["n_hidden"], train=train, backward=backward)
    return orc.din_forward_backward(rec["sd"], rec["x"], rec["y"], DIN_FEATURES, DIN_HISTORY, DIN_FEATURES, DIN_SHARED, m["n_att_hidden"], m["n_hidden"], use_softmax=bool(m["use_softmax"]), train=train,
                                    backward=backward)


def build_model(name, rec, features_mod, models_mod, initializer=None):
    """Our (or the reference's) model with the golden state_dict loaded."""
    import torch
    F, M = features_mod, models_mod
    sd = rec["sd"]
    dense_n, sparse_n = names_of(rec)
    if name.startswith("din"):
        feats = [F.SparseFeature(n, sd["embedding.embed_dict.%s.weight" % n].shape[0], 8) for n in DIN_FEATURES]
        hist = [F.SequenceFeature(n, sd["embedding.embed_dict.%s.weight" % DIN_SHARED[n]].shape[0], 8, pooling="concat", shared_with=DIN_SHARED[n]) for n in DIN_HISTORY]
        model = M.DIN(features=feats, history_features=hist, target_features=feats, mlp_params={"dims": [16, 8]}, attention_mlp_params={"dims": [16, 8], "use_softmax": bool(rec["meta"]["use_softmax"])})
    else:
        dense = [F.DenseFeature(n) for n in dense_n]
        sparse = [F.SparseFeature(n, sd["embedding.embed_dict.%s.weight" % n].shape[0], sd["embedding.embed_dict.%s.weight" % n].shape[1]) for n in sparse_n]
        mlp = {"dims": [16, 8], "dropout": 0.0, "activation": "relu"}
        if name == "deepfm_tutorial":
            model = M.DeepFM(dense + sparse, sparse, mlp)
        elif name == "deepfm_runcriteo":
            model = M.DeepFM(dense, sparse, mlp)
        elif name == "dcn":
            model = M.DCN(dense + sparse, n_cross_layers=3, mlp_params={"dims": [16, 8]})
        else:
            model = M.DCNv2(dense + sparse, n_cross_layers=2, mlp_params=mlp, low_rank=4)
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    return model


def torch_inputs(rec, device="cpu"):
    import torch
    return {k: torch.from_numpy(np.array(v)).to(device) for k, v in rec["x"].items()}, torch.from_numpy(np.array(rec["y"])).to(device)


# ---- two-tower golden (dssm.npz) ------------------------------------------------------------------------------------
DSSM_USER = [("sparse", "user_id", "user_id"), ("seq", "hist_item_id", "item_id", "mean")]
DSSM_ITEM = [("sparse", "item_id", "item_id")]


def dssm_oracle_run(orc, rec, backward=True):
    m = rec["meta"]
    return orc.dssm_forward_backward(rec["sd"], rec["x"], DSSM_USER, DSSM_ITEM, m["n_hidden"], m["n_hidden"], m["neg_ratio"], train=True, backward=backward)


def build_dssm(rec, F, M):
    """This package's (or the reference's) DSSM with the golden state_dict loaded."""
    import torch
    m = rec["meta"]
    user = [F.SparseFeature("user_id", m["n_users"], embed_dim=8), F.SequenceFeature("hist_item_id", m["n_items"], embed_dim=8, pooling="mean", shared_with="item_id")]
    item = [F.SparseFeature("item_id", m["n_items"], embed_dim=8)]
    model = M.DSSM(user, item, user_params={"dims": [16, 8], "activation": "relu"}, item_params={"dims": [16, 8], "activation": "relu"})
    model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in rec["sd"].items()})
    return model


def check_dssm_against_golden(rec, out, grads, emb_tol, grad_rtol):
    """``out``: dict with user_emb / item_emb / prob / logits / loss (numpy); ``grads``: name -> numpy.  The sampled columns are
    compared through the logits they select: duplicate items in a batch tie exactly, and any of the tied columns is a valid pick."""
    for k in ("user_emb", "item_emb", "prob"):
        assert np.abs(out[k] - rec[k]).max() <= emb_tol, (k, np.abs(out[k] - rec[k]).max())
    assert np.abs(out["logits"] - rec["logits"]).max() <= 2 * emb_tol
    assert abs(float(out["loss"]) - float(rec["loss"])) <= 4 * emb_tol
    assert set(grads) == set(rec["grad"]), set(grads) ^ set(rec["grad"])
    for k, ref in rec["grad"].items():
        got = np.asarray(grads[k]).reshape(ref.shape)
        scale = max(np.abs(ref).max(), 1e-3)
        if k.endswith(".bias") and k[:-4] + "weight" in rec["grad"]:  # a Linear bias in front of BatchNorm has a true gradient of 0
            scale = max(scale, np.abs(rec["grad"][k[:-4] + "weight"]).max())
        assert np.abs(got - ref).max() <= grad_rtol * scale, (k, np.abs(got - ref).max(), scale)
