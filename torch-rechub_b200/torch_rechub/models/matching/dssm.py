"""DSSM two-tower retrieval model (mirror of reference ``torch_rechub/models/matching/dssm.py:16-72``).

Each tower is the same hot path as the ranking models: the multi-field gather (with mean-pooled history sequences) into one
flattened tile — ONE fused launch per row width on CUDA (``rh_fields_fwd`` + ``rh_seq_pool_fwd``) — followed by the tower
MLP (tensor-core GEMMs + fused BatchNorm/activation passes).  Both towers share one ``EmbeddingLayer`` so that
``shared_with`` features (the history sequence re-using the item table) resolve to the same table.
"""
import torch
import torch.nn.functional as F

from ...basic.layers import MLP, EmbeddingLayer


class DSSM(torch.nn.Module):
    """Deep Structured Semantic Model.

    Args:
        user_features (list): features of the user tower.
        item_features (list): features of the item tower.
        user_params (dict): user-tower MLP params ``{"dims": list, "activation": str, "dropout": float}``.
        item_params (dict): item-tower MLP params.
        temperature (float): kept for API compatibility; the reference does not apply it (``dssm.py:51``).

    ``mode``: ``None`` -> ``forward`` returns ``sigmoid(<u, v>)`` per sample; ``"user"`` / ``"item"`` -> the L2-normalised
    embedding of that tower only (inference, ``MatchTrainer.inference_embedding``).
    """

    def __init__(self, user_features, item_features, user_params, item_params, temperature=1.0):
        super().__init__()
        self.user_features = user_features
        self.item_features = item_features
        self.temperature = temperature
        self.user_dims = sum(fea.embed_dim for fea in user_features)
        self.item_dims = sum(fea.embed_dim for fea in item_features)
        self.embedding = EmbeddingLayer(user_features + item_features)
        self.user_mlp = MLP(self.user_dims, output_layer=False, **user_params)
        self.item_mlp = MLP(self.item_dims, output_layer=False, **item_params)
        self.mode = None

    def _tower(self, x, features, mlp):
        tile = self.embedding(x, features, squeeze_dim=True)  # (B, sum of embed dims)
        return F.normalize(mlp(tile), p=2, dim=1)

    def user_tower(self, x):
        return None if self.mode == "item" else self._tower(x, self.user_features, self.user_mlp)

    def item_tower(self, x):
        return None if self.mode == "user" else self._tower(x, self.item_features, self.item_mlp)

    def forward(self, x):
        user_embedding = self.user_tower(x)
        item_embedding = self.item_tower(x)
        if self.mode == "user":
            return user_embedding
        if self.mode == "item":
            return item_embedding
        return torch.sigmoid((user_embedding * item_embedding).sum(dim=1))
