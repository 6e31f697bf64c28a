"""Parquet streaming dataset (mirror of reference ``torch_rechub/data/dataset.py:17-121``).

``ParquetIterableDataset(file_paths, columns=None, batch_size=1024)`` yields ``dict[column -> tensor]`` batch by batch without
loading a file whole; under a multi-worker ``DataLoader`` every worker scans its own contiguous share of the files.  Use it with
``DataLoader(ds, batch_size=None)`` as in the reference.  ``packed(id_names, num_names, ...)`` is this package's addition: the same
stream as ``PackedColumns`` batches (ids in ONE int64 block, numerics in one fp32 block) — what the engine's input path wants
(SURVEY.md §8 f1/f4): one H2D copy per block instead of one per column."""
import pyarrow.dataset as pads
from torch.utils.data import IterableDataset, get_worker_info

from .convert import pa_array_to_tensor

_DEFAULT_BATCH_SIZE = 1024


class ParquetIterableDataset(IterableDataset):
    """Stream Parquet data as dicts of tensors.

    Args:
        file_paths (list): paths of the Parquet files (positional only, as in the reference).
        columns (list, optional): columns to read; ``None`` reads all.
        batch_size (int): rows per yielded batch (the last batch of a file may be shorter).
    """

    def __init__(self, file_paths, /, columns=None, batch_size=_DEFAULT_BATCH_SIZE):
        self._file_paths = tuple(str(p) for p in file_paths)
        self._columns = None if columns is None else tuple(columns)
        self._batch_size = batch_size

    def _get_partition(self):
        """This worker's contiguous share of the files (all of them outside a worker process)."""
        info = get_worker_info()
        if info is None:
            return self._file_paths
        n = len(self._file_paths)
        share = (n + info.num_workers - 1) // info.num_workers
        return self._file_paths[info.id * share:min(n, (info.id + 1) * share)]

    def _record_batches(self):
        files = self._get_partition()
        if not files:
            return
        scanner = pads.dataset(files, format="parquet").scanner(columns=None if self._columns is None else list(self._columns), batch_size=self._batch_size)
        yield from scanner.to_batches()

    def __iter__(self):
        for rb in self._record_batches():
            yield {name: pa_array_to_tensor(col) for name, col in zip(rb.column_names, rb.columns)}

    def packed(self, id_names, num_names, seq_names=(), label=None, pin_memory=False):
        """Iterate the same stream as ``(PackedColumns, y)`` batches: integer id columns stacked into one int64 ``(B, n_id)`` block
        (no float round trip), numeric columns into one fp32 ``(B, n_num)`` block, rectangular list columns into ``(B, n_seq, L)``;
        ``label`` names the target column (``y`` is ``None`` without it)."""
        import numpy as np
        import torch

        from ..b200.data import PackedColumns
        pin = (lambda t: t.pin_memory()) if (pin_memory and torch.cuda.is_available()) else (lambda t: t)
        col = lambda rb, n: rb.column(rb.schema.get_field_index(n))
        for rb in self._record_batches():
            ids = pin(torch.from_numpy(np.stack([col(rb, n).to_numpy(zero_copy_only=False).astype(np.int64) for n in id_names], axis=1))) if id_names else None
            nums = pin(torch.from_numpy(np.stack([col(rb, n).to_numpy(zero_copy_only=False).astype(np.float32) for n in num_names], axis=1))) if num_names else None
            seqs = pin(torch.stack([pa_array_to_tensor(col(rb, n)).long() for n in seq_names], dim=1)) if seq_names else None
            y = pin(pa_array_to_tensor(col(rb, label))) if label is not None else None
            yield PackedColumns(list(id_names), ids, list(num_names), nums, list(seq_names), seqs), y
