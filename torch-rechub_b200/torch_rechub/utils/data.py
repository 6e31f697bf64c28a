"""Input helpers the ranking examples import (mirror of reference ``torch_rechub/utils/data.py:14-289``).

Host-side code; only the pieces ``examples/ranking/run_criteo.py`` and ``run_amazon_electronics.py`` use
(``DataGenerator``, ``TorchDataset``, ``df_to_dict``, ``generate_seq_feature``, ``pad_sequences``,
``get_auto_embedding_dim``) plus ``MatchDataGenerator`` for the two-tower path.  Generative dataset helpers are
out of scope.
"""
import random

import numpy as np
import pandas as pd
import torch
import tqdm
from sklearn.metrics import mean_squared_error, roc_auc_score
from sklearn.preprocessing import LabelEncoder
from torch.utils.data import DataLoader, Dataset, random_split


class TorchDataset(Dataset):
    """``(x, y)`` with ``x`` a mapping column -> array-like; sample ``i`` is ``({k: x[k][i]}, y[i])`` (reference ``:14-25``)."""

    def __init__(self, x, y):
        super().__init__()
        self.x = x
        self.y = y

    def __getitem__(self, index):
        return {k: v[index] for k, v in self.x.items()}, self.y[index]

    def __len__(self):
        return len(self.y)


class PredictDataset(Dataset):
    """Label-free variant (reference ``:28-38``)."""

    def __init__(self, x):
        super().__init__()
        self.x = x

    def __getitem__(self, index):
        return {k: v[index] for k, v in self.x.items()}

    def __len__(self):
        return len(self.x[next(iter(self.x.keys()))])


class MatchDataGenerator(object):
    """Train / test-user / all-item DataLoaders of the matching examples (reference ``:41-58``): a labelled ``TorchDataset`` when
    ``y`` is given, a label-free one for pair-wise training; the test and item loaders keep their order (ground-truth alignment)."""

    def __init__(self, x, y=[]):
        super().__init__()
        self.dataset = TorchDataset(x, y) if len(y) != 0 else PredictDataset(x)

    def generate_dataloader(self, x_test_user, x_all_item, batch_size, num_workers=8):
        train = DataLoader(self.dataset, batch_size=batch_size, shuffle=True, num_workers=num_workers)
        ordered = lambda cols: DataLoader(PredictDataset(cols), batch_size=batch_size, shuffle=False, num_workers=num_workers)
        return train, ordered(x_test_user), ordered(x_all_item)


class DataGenerator(object):
    """Builds train/val/test DataLoaders from a DataFrame or a dict of arrays (reference ``:61-83``)."""

    def __init__(self, x, y):
        super().__init__()
        self.dataset = TorchDataset(x, y)
        self.length = len(self.dataset)

    def generate_dataloader(self, x_val=None, y_val=None, x_test=None, y_test=None, split_ratio=None, batch_size=16, num_workers=0):
        if split_ratio is not None:
            n_train = int(self.length * split_ratio[0])
            n_val = int(self.length * split_ratio[1])
            n_test = self.length - n_train - n_val
            print("the samples of train : val : test are  %d : %d : %d" % (n_train, n_val, n_test))
            train_set, val_set, test_set = random_split(self.dataset, (n_train, n_val, n_test))
        else:
            train_set = self.dataset
            val_set = TorchDataset(x_val, y_val)
            test_set = TorchDataset(x_test, y_test)
        mk = lambda ds, shuffle: DataLoader(ds, batch_size=batch_size, shuffle=shuffle, num_workers=num_workers)
        return mk(train_set, True), mk(val_set, False), mk(test_set, False)


def get_auto_embedding_dim(num_classes):
    """``floor(6 * num_classes ** 0.25)`` — the DCN rule of thumb (reference ``:86-101``)."""
    return int(np.floor(6 * np.power(num_classes, 0.25)))


def get_loss_func(task_type="classification"):
    if task_type == "classification":
        return torch.nn.BCELoss()
    if task_type == "regression":
        return torch.nn.MSELoss()
    raise ValueError("task_type must be classification or regression")


def get_metric_func(task_type="classification"):
    if task_type == "classification":
        return roc_auc_score
    if task_type == "regression":
        return mean_squared_error
    raise ValueError("task_type must be classification or regression")


def neg_sample(click_hist, item_size):
    """A random item id in ``[1, item_size]`` the user has not clicked (reference ``:239-243``)."""
    while True:
        cand = random.randint(1, item_size)
        if cand not in click_hist:
            return cand


def generate_seq_feature(data, user_col, item_col, time_col, item_attribute_cols=[], min_item=0, shuffle=True, max_len=50):
    """Sliding-window behaviour sequences with one sampled negative per positive (reference ``:122-216``).

    Every column is label-encoded to ``1..n`` (0 is the padding id); per user, for step ``i`` the history is
    the first ``i`` items post-padded with 0 to ``max_len``; the last step goes to test, the one before to
    validation, the rest to train.  Returns three DataFrames with columns
    ``label, target_item_id, <user_col>, hist_item_id[, hist_<attr>, target_<attr> ...]``.
    """
    for col in data:
        data[col] = LabelEncoder().fit_transform(data[col]) + 1  # 0 is reserved for padding
    data = data.astype('int32')
    n_items = data[item_col].max()
    attr_of = {col: data[[item_col, col]].set_index([item_col])[col].to_dict() for col in item_attribute_cols}

    buckets = {"train": [], "val": [], "test": []}
    data.sort_values(time_col, inplace=True)
    for uid, hist in tqdm.tqdm(data.groupby(user_col), desc='generate sequence features'):
        pos_list = hist[item_col].tolist()
        n_pos = len(pos_list)
        if n_pos < min_item:
            continue
        neg_list = [neg_sample(pos_list, n_items) for _ in range(n_pos)]
        attr_hist = {col: hist[col].tolist() for col in item_attribute_cols}
        for i in range(1, min(n_pos, max_len)):
            padded = pos_list[:i] + [0] * (max_len - i)
            rows = {1: [1, pos_list[i], uid, padded], 0: [0, neg_list[i], uid, padded]}
            for col in item_attribute_cols:
                h_attr = attr_hist[col][:i] + [0] * (max_len - i)
                rows[1] += [h_attr, attr_of[col][pos_list[i]]]
                rows[0] += [h_attr, attr_of[col][neg_list[i]]]
            where = "test" if i == n_pos - 1 else ("val" if i == n_pos - 2 else "train")
            buckets[where].append(rows[1])
            buckets[where].append(rows[0])

    columns = ['label', 'target_item_id', user_col, 'hist_item_id']
    for col in item_attribute_cols:
        columns += ['hist_' + col, 'target_' + col]
    if shuffle:
        for part in ("train", "val", "test"):
            random.shuffle(buckets[part])
    return tuple(pd.DataFrame(buckets[part], columns=columns) for part in ("train", "val", "test"))


def df_to_dict(data):
    """DataFrame -> ``{column: np.ndarray}`` (list-valued cells become 2-D arrays; reference ``:219-236``)."""
    as_lists = data.to_dict('list')
    return {key: np.array(as_lists[key]) for key in data.keys()}


def pad_sequences(sequences, maxlen=None, dtype='int32', padding='pre', truncating='pre', value=0.):
    """Keras-style padding of a list of lists to ``(n, maxlen)`` (reference ``:245-289``)."""
    assert padding in ["pre", "post"], "Invalid padding={}.".format(padding)
    assert truncating in ["pre", "post"], "Invalid truncating={}.".format(truncating)
    if maxlen is None:
        maxlen = max(len(s) for s in sequences)
    out = np.full((len(sequences), maxlen), value, dtype=dtype)
    for row, seq in enumerate(sequences):
        if len(seq) == 0:
            continue
        kept = np.asarray(seq[-maxlen:] if truncating == 'pre' else seq[:maxlen], dtype=dtype)
        if padding == 'pre':
            out[row, -len(kept):] = kept
        else:
            out[row, :len(kept)] = kept
    return out
