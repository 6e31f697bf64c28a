"""Losses on the CTR path (mirror of reference ``torch_rechub/basic/loss_func.py:6-68``).

``RegularizationLoss`` is on the CTR hot path (``trainers/ctr_trainer.py:94`` calls it every step); ``BPRLoss`` is the
pair-wise criterion of ``MatchTrainer`` (reference ``loss_func.py:95-107``).  The generative-model losses of
the reference file are out of scope.
"""
import torch
import torch.nn as nn

_NORM_TYPES = (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d, nn.LayerNorm, nn.GroupNorm, nn.InstanceNorm1d, nn.InstanceNorm2d, nn.InstanceNorm3d)
_TABLE_TYPES = (nn.Embedding, nn.EmbeddingBag)


class RegularizationLoss(nn.Module):
    """L1/L2 penalties, with separate coefficients for embedding tables and for dense weights.

    Parameters of normalisation layers are skipped; tables are recognised by module type
    (``isinstance(nn.Embedding/EmbeddingBag)``, reference ``loss_func.py:45-49``).  With all four
    coefficients at 0 the result is the Python float ``0.0`` (no tensor work), as in the reference.
    """

    def __init__(self, embedding_l1=0.0, embedding_l2=0.0, dense_l1=0.0, dense_l2=0.0):
        super(RegularizationLoss, self).__init__()
        self.embedding_l1 = embedding_l1
        self.embedding_l2 = embedding_l2
        self.dense_l1 = dense_l1
        self.dense_l2 = dense_l2

    def forward(self, model):
        if not (self.embedding_l1 > 0 or self.embedding_l2 > 0 or self.dense_l1 > 0 or self.dense_l2 > 0):
            return 0.0
        skip, table_ids = set(), set()
        for module in model.modules():
            if isinstance(module, _NORM_TYPES):
                skip.update(id(p) for p in module.parameters())
            elif isinstance(module, _TABLE_TYPES):
                table_ids.update(id(p) for p in module.parameters())
        total = 0.0
        for p in model.parameters():
            if not p.requires_grad or id(p) in skip:
                continue
            l1, l2 = (self.embedding_l1, self.embedding_l2) if id(p) in table_ids else (self.dense_l1, self.dense_l2)
            if l1 > 0:
                total = total + l1 * torch.sum(torch.abs(p))
            if l2 > 0:
                total = total + l2 * torch.sum(p**2)
        return total


class BPRLoss(nn.Module):
    """Bayesian personalised ranking: ``mean(-log sigmoid(pos - neg))`` (reference ``loss_func.py:95-107``).

    ``neg_score`` may hold one negative per sample ``(B,)`` or several ``(B, K)``; ``in_batch_neg`` is accepted for call
    compatibility with ``MatchTrainer`` and changes nothing.
    """

    def forward(self, pos_score, neg_score, in_batch_neg=False):
        pos = pos_score.reshape(-1)
        margin = pos - neg_score if neg_score.dim() == 1 else pos.unsqueeze(1) - neg_score
        return -torch.log(torch.sigmoid(margin)).mean()
