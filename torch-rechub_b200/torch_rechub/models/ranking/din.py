"""Deep Interest Network (mirror of reference ``torch_rechub/models/ranking/din.py:16-93``)."""
import torch
import torch.nn as nn

from ...basic.features import SequenceFeature, SparseFeature
from ...basic.layers import MLP, EmbeddingLayer


class DIN(nn.Module):
    """Target attention of every behaviour sequence against its paired target feature, then an MLP.

    Args:
        features (list): profile / context features (fed to the final MLP only).
        history_features (list): behaviour sequences (``pooling="concat"``), one ActivationUnit each.
        target_features (list): target features; ``history_features[i]`` attends to ``target_features[i]``.
        mlp_params (dict): final tower (activation forced to ``"dice"``, din.py:36).
        attention_mlp_params (dict): ``{"dims": list, "activation": str, "use_softmax": bool}``.

    As in the reference NO padding mask enters the attention (ConcatPooling ignores it, layers.py:204-205):
    padded positions attend with whatever row id 0 holds.
    """

    def __init__(self, features, history_features, target_features, mlp_params, attention_mlp_params):
        super().__init__()
        self.features = features
        self.history_features = history_features
        self.target_features = target_features
        self.num_history_features = len(history_features)
        self.all_dims = sum([fea.embed_dim for fea in features + history_features + target_features])
        self.embedding = EmbeddingLayer(features + history_features + target_features)
        self.attention_layers = nn.ModuleList([ActivationUnit(fea.embed_dim, **attention_mlp_params) for fea in self.history_features])
        self.mlp = MLP(self.all_dims, activation="dice", **mlp_params)

    def _fusable(self, x):
        n = self.num_history_features
        if n == 0 or n > len(self.target_features):
            return False
        for i, h in enumerate(self.history_features):
            t = self.target_features[i]
            if not (isinstance(h, SequenceFeature) and h.pooling == "concat" and isinstance(t, SparseFeature)):
                return False
            if x[h.name].dim() != 2 or x[t.name].dim() != 1 or h.embed_dim != t.embed_dim:
                return False
            if self.embedding.table_of(h).weight.shape[1] != h.embed_dim or self.embedding.table_of(t).weight.shape[1] != h.embed_dim:
                return False
        return all(isinstance(f, SparseFeature) for f in list(self.target_features) + list(self.features))

    def _forward_fused(self, x):
        from ...b200 import ops
        n = self.num_history_features
        pooled, targets = [], []
        for i in range(n):
            h, t = self.history_features[i], self.target_features[i]
            # gather of the sequence + [t, h, t-h, t*h] in one kernel (din.py:42,44,80-81)
            att_in, hist, tgt = ops.din_attention_input(self.embedding.table_of(h), self.embedding.table_of(t), x[h.name], x[t.name])
            unit = self.attention_layers[i]
            att_w = unit.attention(att_in).view(-1, hist.shape[1])
            pooled.append(ops.din_weighted_sum(att_w, hist, unit.use_softmax))
            targets.append(tgt)
        rest = list(self.target_features[n:]) + list(self.features)
        parts = pooled + targets
        if rest:
            parts.append(self.embedding(x, rest, squeeze_dim=True))
        return torch.cat(parts, dim=1)

    def forward(self, x):
        if self.embedding._on_cuda(x, self.features + self.history_features + self.target_features) and self._fusable(x):
            mlp_in = self._forward_fused(x)
        else:
            embed_x_features = self.embedding(x, self.features)  # (B, n_features, D)
            embed_x_history = self.embedding(x, self.history_features)  # (B, n_hist, L, D)
            embed_x_target = self.embedding(x, self.target_features)  # (B, n_target, D)
            attention_pooling = []
            for i in range(self.num_history_features):
                attention_seq = self.attention_layers[i](embed_x_history[:, i, :, :], embed_x_target[:, i, :])
                attention_pooling.append(attention_seq.unsqueeze(1))
            attention_pooling = torch.cat(attention_pooling, dim=1)  # (B, n_hist, D)
            mlp_in = torch.cat([attention_pooling.flatten(start_dim=1), embed_x_target.flatten(start_dim=1), embed_x_features.flatten(start_dim=1)], dim=1)
        if mlp_in.is_cuda:
            from ...b200 import config
            if config.fused_head_all:
                p = self.mlp.forward_head(mlp_in, (), sigmoid=True)
                if p is not None:
                    return p
        y = self.mlp(mlp_in)
        return torch.sigmoid(y.squeeze(1))


class ActivationUnit(nn.Module):
    """DIN's target attention for one behaviour sequence (reference ``din.py:58-93``).

    ``w[b,l] = MLP([t, h_l, t-h_l, t*h_l])`` (optionally softmaxed over l), output ``sum_l w[b,l] h_l``.

    Shape: history ``(B, L, D)``, target ``(B, D)`` -> ``(B, D)``.
    """

    def __init__(self, emb_dim, dims=None, activation="dice", use_softmax=False):
        super(ActivationUnit, self).__init__()
        if dims is None:
            dims = [36]
        self.emb_dim = emb_dim
        self.use_softmax = use_softmax
        self.attention = MLP(4 * self.emb_dim, dims=dims, activation=activation)

    def forward(self, history, target):
        seq_length = history.size(1)
        target = target.unsqueeze(1).expand(-1, seq_length, -1)
        att_input = torch.cat([target, history, target - history, target * history], dim=-1)  # (B, L, 4D)
        att_weight = self.attention(att_input.view(-1, 4 * self.emb_dim)).view(-1, seq_length)
        if history.is_cuda and history.dtype == torch.float32:
            from ...b200 import ops
            return ops.din_weighted_sum(att_weight, history, self.use_softmax)
        if self.use_softmax:
            att_weight = att_weight.softmax(dim=-1)
        return (att_weight.unsqueeze(-1) * history).sum(dim=1)
