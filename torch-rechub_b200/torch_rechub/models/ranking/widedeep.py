"""Wide & Deep (mirror of reference ``torch_rechub/models/ranking/widedeep.py``) — a two-call user of the fused
``EmbeddingLayer`` front end, kept because the reference's own ranking e2e test drives it (tests/test_e2e_ranking.py:53)."""
import torch

from ...basic.layers import LR, MLP, EmbeddingLayer


class WideDeep(torch.nn.Module):
    """``sigmoid(LR(e_wide) + MLP(e_deep))``.

    Args:
        wide_features (list): features of the linear (wide) part.
        deep_features (list): features of the MLP (deep) part.
        mlp_params (dict): ``{"dims": list, "activation": str, "dropout": float, "output_layer": bool}``.
    """

    def __init__(self, wide_features, deep_features, mlp_params):
        super().__init__()
        width = lambda feas: sum(fea.embed_dim for fea in feas)
        self.wide_features, self.deep_features = wide_features, deep_features
        self.wide_dims, self.deep_dims = width(wide_features), width(deep_features)
        # creation order linear -> embedding -> mlp: the reference's RNG consumption (widedeep.py:27-29)
        self.linear = LR(self.wide_dims)
        self.embedding = EmbeddingLayer(wide_features + deep_features)
        self.mlp = MLP(self.deep_dims, **mlp_params)

    def forward(self, x):
        wide, deep = (self.embedding(x, feas, squeeze_dim=True) for feas in (self.wide_features, self.deep_features))
        if deep.is_cuda:
            from ...b200 import config
            if config.fused_head_all:  # the wide term rides along as a per-sample extra of the deep tower's fused head
                p = self.mlp.forward_head(deep, (self.linear(wide).squeeze(1),), sigmoid=True)
                if p is not None:
                    return p
        return torch.sigmoid((self.linear(wide) + self.mlp(deep)).squeeze(1))
