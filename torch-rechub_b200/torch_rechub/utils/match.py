"""Matching-path helpers (mirror of the parts of reference ``torch_rechub/utils/match.py`` that the two-tower training path
uses): sample generation for point/pair/list-wise training, the in-batch negative sampler and the logit
gather that ``MatchTrainer`` calls every step (reference ``match.py:104-161``).

Same results as the reference for the same seeds on CPU (same RNG calls in the same order).  On CUDA the sampler is one
batched draw instead of the reference's Python loop with one ``randperm`` per row (``match.py:136-145``): a B x B score
matrix at B = 4096 would otherwise cost 4096 host round trips per step.

The approximate-nearest-neighbour wrappers of the reference file (Annoy / Faiss / Milvus) are serving utilities outside this
engine's scope; their names raise with a pointer to upstream.
"""
import random
from collections import Counter

import numpy as np
import pandas as pd
import torch
import tqdm

from .data import df_to_dict, pad_sequences


def gen_model_input(df, user_profile, user_col, item_profile, item_col, seq_max_len, padding='pre', truncating='pre'):
    """Join the user and item profiles onto ``df`` (left joins: sample order is kept) and pad every ``hist_*`` / ``tag_*`` list
    column to ``seq_max_len``; returns the ``dict`` of arrays the models take (reference ``match.py:32-59``)."""
    df = pd.merge(df, user_profile, on=user_col, how='left')
    df = pd.merge(df, item_profile, on=item_col, how='left')
    for prefix in ("hist_", "tag_"):
        for col in df.columns.to_list():
            if col.startswith(prefix):
                df[col] = pad_sequences(df[col], maxlen=seq_max_len, value=0, padding=padding, truncating=truncating).tolist()
    return df_to_dict(df)


def negative_sample(items_cnt_order, ratio, method_id=0):
    """``ratio`` negative item ids drawn from the item-count table (keys sorted by count, descending).

    method 0: uniform; 1: ``count^0.75`` (word2vec); 2: ``log(count + 1) + 1e-6``; 3: Tencent RALM rank weights, without
    replacement (reference ``match.py:62-101``).  One ``np.random.choice`` call, as in the reference.
    """
    items = list(items_cnt_order.keys())
    if method_id == 0:
        return np.random.choice(items, size=ratio, replace=True)
    if method_id == 1:
        weights = np.array([c**0.75 for c in items_cnt_order.values()])
    elif method_id == 2:
        weights = np.array([np.log(c + 1) + 1e-6 for c in items_cnt_order.values()])
    elif method_id == 3:
        n = len(items_cnt_order)
        weights = np.array([(np.log(k + 2) - np.log(k + 1)) / np.log(n + 1) for k in items_cnt_order.values()])
    else:
        raise ValueError("method id should in (0,1,2,3)")
    return np.random.choice(items, size=ratio, replace=method_id != 3, p=weights / weights.sum())


def generate_seq_feature_match(data, user_col, item_col, time_col, item_attribute_cols=None, sample_method=0, mode=0, neg_ratio=0, min_item=0):
    """Leave-last-out sequence samples with negatives for matching (reference ``match.py:163-248``).

    Per user (events sorted by time) every prefix ``items[:i]`` predicts ``items[i]``; the last event goes to the test set.
    mode 0 (point-wise): a ``label`` column, each positive followed by ``neg_ratio`` negatives; mode 1 (pair-wise): one
    ``neg_items`` id per row; mode 2 (list-wise): ``neg_ratio`` ids per row.  Returns ``(df_train, df_test)``.
    """
    attrs = list(item_attribute_cols or [])
    if mode == 2:
        assert neg_ratio > 0, 'neg_ratio must be greater than 0 when list-wise learning'
    elif mode == 1:
        neg_ratio = 1
    if mode not in (0, 1, 2):
        raise ValueError("mode should in (0,1,2)")
    print("preprocess data")
    data.sort_values(time_col, inplace=True)
    counts = dict(sorted(Counter(data[item_col].tolist()).items(), key=lambda kv: kv[1], reverse=True))
    negatives = negative_sample(counts, ratio=data.shape[0] * neg_ratio, method_id=sample_method)
    cursor = 0
    last_col = "label" if mode == 0 else "neg_items"
    train_rows, test_rows, n_cold = [], [], 0
    for uid, events in tqdm.tqdm(data.groupby(user_col), desc='generate sequence features'):
        items = events[item_col].tolist()
        if len(items) < min_item:
            n_cold += 1
            continue
        attr_lists = [events[c].tolist() for c in attrs]
        for i in range(1, len(items)):
            prefix = items[:i]
            head = [uid, items[i], prefix, len(prefix)] + [a[:i] for a in attr_lists]
            if i == len(items) - 1:  # the user's last event: held out (its last column is only a placeholder for modes 1, 2)
                test_rows.append(head + [1])
            elif mode == 0:
                train_rows.append(head + [1])
                for _ in range(neg_ratio):
                    neg_row = list(head)
                    neg_row[1] = negatives[cursor]
                    cursor += 1
                    train_rows.append(neg_row + [0])
            elif mode == 1:
                train_rows.append(head + [negatives[cursor]])
                cursor += 1
            else:
                train_rows.append(head + [negatives[cursor:cursor + neg_ratio]])
                cursor += neg_ratio
    random.shuffle(train_rows)
    random.shuffle(test_rows)
    print("n_train: %d, n_test: %d" % (len(train_rows), len(test_rows)))
    print("%d cold start user dropped " % n_cold)
    columns = [user_col, item_col, "hist_" + item_col, "histlen_" + item_col] + ["hist_" + c for c in attrs] + [last_col]
    return pd.DataFrame(train_rows, columns=columns), pd.DataFrame(test_rows, columns=columns)


def inbatch_negative_sampling(scores, neg_ratio=None, hard_negative=False, generator=None):
    """Indices ``(B, K)`` of in-batch negatives for every row of a ``(B, B)`` score matrix; the diagonal (the positive) is never
    drawn (reference ``match.py:104-147``).

    ``K = neg_ratio`` clipped to ``B - 1`` (also when omitted or non-positive).  ``hard_negative``: the K best-scoring
    off-diagonal columns per row; otherwise K columns uniformly without replacement.  CPU draws replay the reference's RNG
    stream (one ``randperm(B - 1)`` per row); CUDA draws come from one ``(B, B)`` uniform key matrix.
    """
    if scores.dim() != 2:
        raise ValueError(f"inbatch_negative_sampling expects 2D scores, got shape {tuple(scores.shape)}")
    n = scores.size(0)
    if n <= 1:
        raise ValueError("In-batch negative sampling requires batch_size > 1")
    k = neg_ratio if (neg_ratio is not None and 0 < neg_ratio <= n - 1) else n - 1
    if scores.device.type == "cuda":
        return _sample_batched(scores, k, hard_negative, generator)
    return _sample_rowwise(scores, k, hard_negative, generator)


def _sample_batched(scores, k, hard_negative, generator):
    """All rows at once (the CUDA route): top-k of the diagonal-masked scores, or the k smallest of one (B, B) matrix of uniform
    keys whose diagonal is pushed to the end — a uniform draw without replacement from the other B - 1 columns."""
    n = scores.size(0)
    if hard_negative:
        masked = scores.detach().clone()
        masked.fill_diagonal_(float("-inf"))
        return torch.topk(masked, k=k, dim=1).indices
    keys = torch.rand((n, n), device=scores.device, generator=generator)
    keys.fill_diagonal_(2.0)  # sorts last: never among the first k <= B - 1
    return torch.argsort(keys, dim=1)[:, :k]


def _sample_rowwise(scores, k, hard_negative, generator):
    """Row by row (the CPU route): replays the reference's RNG stream — one ``randperm(B - 1)`` per row — and its top-k tie order."""
    n = scores.size(0)
    out = torch.empty((n, k), dtype=torch.long, device=scores.device)
    if hard_negative:
        masked = scores.detach().clone()
        masked.fill_diagonal_(float("-inf"))
        for i in range(n):
            out[i] = torch.topk(masked[i], k=k).indices
        return out
    for i in range(n):
        pick = torch.randperm(n - 1, device=scores.device, generator=generator)[:k]  # positions in the row with its own column removed
        out[i] = pick + (pick >= i).long()
    return out


def gather_inbatch_logits(scores, neg_indices):
    """``(B, 1 + K)``: column 0 the positive ``scores[i, i]``, then ``scores[i, neg_indices[i, j]]`` (reference ``match.py:150-161``)."""
    return torch.cat([torch.diagonal(scores).unsqueeze(1), torch.gather(scores, 1, neg_indices.to(scores.device))], dim=1)


def _serving_only(name):

    class _ServingOnly(object):
        __doc__ = "%s (vector retrieval for serving) is outside the B200 hot-path engine; use upstream torch-rechub for it." % name

        def __init__(self, *args, **kwargs):
            raise NotImplementedError(self.__doc__)

    _ServingOnly.__name__ = _ServingOnly.__qualname__ = name
    return _ServingOnly


Annoy = _serving_only("Annoy")
Milvus = _serving_only("Milvus")
Faiss = _serving_only("Faiss")
