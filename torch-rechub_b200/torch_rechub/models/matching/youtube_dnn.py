"""YoutubeDNN retrieval model (mirror of reference ``torch_rechub/models/matching/youtube_dnn.py:15-76``).

A user tower against RAW item embeddings (no item MLP), trained list-wise: every sample carries one positive item and
``n_neg`` sampled negatives (a ``pooling="concat"`` sequence feature sharing the item table), and the model returns the
``(B, 1 + n_neg)`` cosine logits ``MatchTrainer`` feeds to a softmax cross entropy with the positive in column 0.
The towers ride the engine's fused gather on CUDA like every other ``EmbeddingLayer`` consumer.
"""
import torch
import torch.nn.functional as F

from ...basic.layers import MLP, EmbeddingLayer


class YoutubeDNN(torch.nn.Module):
    """Args:
        user_features (list): features of the user tower.
        item_features (list): the item id feature (its table IS the item representation).
        neg_item_feature (list): the negative item ids, a concat-pooled sequence feature sharing the item table.
        user_params (dict): user-tower MLP params; its last dim must equal the item ``embed_dim``.
        temperature (float): logits are divided by it.

    ``mode``: ``None`` -> logits ``(B, 1 + n_neg)``; ``"user"`` / ``"item"`` -> that side's unit-norm ``(B, D)`` embeddings.
    """

    def __init__(self, user_features, item_features, neg_item_feature, user_params, temperature=1.0):
        super().__init__()
        self.user_features, self.item_features, self.neg_item_feature = user_features, item_features, neg_item_feature
        self.temperature = temperature
        self.user_dims = sum(fea.embed_dim for fea in user_features)
        self.embedding = EmbeddingLayer(user_features + item_features)
        self.user_mlp = MLP(self.user_dims, output_layer=False, **user_params)
        self.mode = None

    def user_tower(self, x):
        if self.mode == "item":
            return None
        tile = self.embedding(x, self.user_features, squeeze_dim=True)
        unit = F.normalize(self.user_mlp(tile).unsqueeze(1), p=2, dim=2)  # (B, 1, D)
        return unit.squeeze(1) if self.mode == "user" else unit

    def item_tower(self, x):
        if self.mode == "user":
            return None
        positive = F.normalize(self.embedding(x, self.item_features, squeeze_dim=False), p=2, dim=2)  # (B, 1, D)
        if self.mode == "item":
            return positive.squeeze(1)
        negatives = F.normalize(self.embedding(x, self.neg_item_feature, squeeze_dim=False).squeeze(1), p=2, dim=2)  # (B, n_neg, D)
        return torch.cat((positive, negatives), dim=1)

    def forward(self, x):
        user, items = self.user_tower(x), self.item_tower(x)
        if self.mode == "user":
            return user
        if self.mode == "item":
            return items
        return (user * items).sum(dim=2) / self.temperature
