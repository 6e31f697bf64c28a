"""FaceBookDSSM (mirror of reference ``torch_rechub/models/matching/dssm_facebook.py:15-82``): DSSM trained pair-wise — each
sample carries a positive and a sampled negative item that go through the SAME item tower; ``forward`` returns
``(pos_score, neg_score)`` for ``MatchTrainer``'s BPR criterion."""
import torch
import torch.nn.functional as F

from ...basic.layers import MLP, EmbeddingLayer


class FaceBookDSSM(torch.nn.Module):
    """Args:
        user_features (list): features of the user tower.
        pos_item_features (list): features of the positive item.
        neg_item_features (list): features of the sampled negative item (same tower, usually tables shared with the positive's).
        user_params / item_params (dict): tower MLP params.
        temperature (float): kept for API compatibility (unused by the reference too).

    ``mode``: ``None`` -> ``(pos_score, neg_score)`` cosine scores; ``"user"`` -> unit-norm user embeddings; ``"item"`` -> the
    item tower's output for the positive features, NOT normalised (the reference's behaviour, ``dssm_facebook.py:74-75``).
    """

    def __init__(self, user_features, pos_item_features, neg_item_features, user_params, item_params, temperature=1.0):
        super().__init__()
        self.user_features, self.pos_item_features, self.neg_item_features = user_features, pos_item_features, neg_item_features
        self.temperature = temperature
        self.user_dims = sum(fea.embed_dim for fea in user_features)
        self.item_dims = sum(fea.embed_dim for fea in pos_item_features)
        self.embedding = EmbeddingLayer(user_features + pos_item_features + neg_item_features)
        self.user_mlp = MLP(self.user_dims, output_layer=False, **user_params)
        self.item_mlp = MLP(self.item_dims, output_layer=False, **item_params)
        self.mode = None

    def _item(self, x, features):
        return self.item_mlp(self.embedding(x, features, squeeze_dim=True))

    def user_tower(self, x):
        if self.mode == "item":
            return None
        return F.normalize(self.user_mlp(self.embedding(x, self.user_features, squeeze_dim=True)), p=2, dim=1)

    def item_tower(self, x):
        if self.mode == "user":
            return None, None
        positive = self._item(x, self.pos_item_features)
        if self.mode == "item":
            return positive, None
        negative = self._item(x, self.neg_item_features)
        return F.normalize(positive, p=2, dim=1), F.normalize(negative, p=2, dim=1)

    def forward(self, x):
        user = self.user_tower(x)
        positive, negative = self.item_tower(x)
        if self.mode == "user":
            return user
        if self.mode == "item":
            return positive
        return (user * positive).sum(dim=1), (user * negative).sum(dim=1)
