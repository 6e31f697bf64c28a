"""Columnar batches: the input path the engine wants.

The reference's ``DataLoader(TorchDataset)`` indexes every sample of a batch in Python and collates 39 columns
(93 ms per 4096-batch measured) and the trainer then issues one H2D copy per column
(``trainers/ctr_trainer.py:84``).  ``PackedColumns`` keeps a batch as at most three tensors — ids ``(B, n_id)``
int64, numeric ``(B, n_num)`` fp32, sequences ``(B, n_seq, L)`` int64 — and still IS the ``dict[str, Tensor]`` the
models take (each value is a column view, so ``x[name]`` works unchanged and the fused gather reads the packed ids
with a stride).  ``PackedLoader`` slices pre-packed (optionally pinned) arrays; ``DataGenerator`` users opt in with
``packed_loader(...)``.
"""
import numpy as np
import torch


class PackedColumns(dict):
    """dict name -> column view over a few packed tensors.  ``.to(device)`` moves the packed tensors (<= 3 copies)."""

    def __init__(self, id_names, ids, num_names, nums, seq_names=(), seqs=None):
        super().__init__()
        self.id_names, self.num_names, self.seq_names = list(id_names), list(num_names), list(seq_names)
        self.ids, self.nums, self.seqs = ids, nums, seqs
        for j, n in enumerate(self.id_names):
            self[n] = ids[:, j]
        for j, n in enumerate(self.num_names):
            self[n] = nums[:, j]
        for j, n in enumerate(self.seq_names):
            self[n] = seqs[:, j]

    def to(self, device, non_blocking=True):
        mv = lambda t: None if t is None else t.to(device, non_blocking=non_blocking)
        return PackedColumns(self.id_names, mv(self.ids), self.num_names, mv(self.nums), self.seq_names, mv(self.seqs))

    def clone(self):
        cl = lambda t: None if t is None else t.clone()
        return PackedColumns(self.id_names, cl(self.ids), self.num_names, cl(self.nums), self.seq_names, cl(self.seqs))

    def copy_into(self, other, non_blocking=True):
        """Copy this batch into the (static, same-shaped) buffers of ``other``."""
        for a, b in ((self.ids, other.ids), (self.nums, other.nums), (self.seqs, other.seqs)):
            if a is not None:
                b.copy_(a, non_blocking=non_blocking)

    def h2d_bytes(self):
        return sum(t.numel() * t.element_size() for t in (self.ids, self.nums, self.seqs) if t is not None)


class PackedLoader(object):
    """Iterates ``(PackedColumns, y)`` batches over pre-packed host arrays (no per-sample Python work)."""

    def __init__(self, x, y, batch_size, id_names=None, num_names=None, seq_names=(), shuffle=False, pin_memory=True, drop_last=False, id_dtype=torch.int64):
        """``id_dtype=torch.int32`` halves the id bytes per batch (H2D copy and kernel reads; SURVEY §8d: 1776 instead of
        1880 B/sample on the fused forward) when every vocabulary fits 31 bits; the kernels read either width."""
        cols = {k: (v.values if hasattr(v, "values") else np.asarray(v)) for k, v in x.items()}
        if id_names is None or num_names is None:
            id_names = [k for k, v in cols.items() if v.ndim == 1 and np.issubdtype(v.dtype, np.integer)]
            num_names = [k for k, v in cols.items() if v.ndim == 1 and not np.issubdtype(v.dtype, np.integer)]
            seq_names = [k for k, v in cols.items() if v.ndim == 2]
        self.id_names, self.num_names, self.seq_names = list(id_names), list(num_names), list(seq_names)
        n = len(y)
        pin = (lambda t: t.pin_memory()) if (pin_memory and torch.cuda.is_available()) else (lambda t: t)
        if id_dtype not in (torch.int64, torch.int32):
            raise ValueError("id_dtype must be torch.int64 or torch.int32")
        np_ids = np.int64 if id_dtype == torch.int64 else np.int32
        if self.id_names and id_dtype == torch.int32 and max((int(np.max(cols[k])) for k in self.id_names if len(cols[k])), default=0) >= 2**31:
            raise ValueError("an id does not fit int32; keep id_dtype=torch.int64")
        self.ids = pin(torch.from_numpy(np.stack([cols[k].astype(np_ids) for k in self.id_names], axis=1))) if self.id_names else None
        self.nums = pin(torch.from_numpy(np.stack([cols[k].astype(np.float32) for k in self.num_names], axis=1))) if self.num_names else None
        self.seqs = pin(torch.from_numpy(np.stack([cols[k].astype(np_ids) for k in self.seq_names], axis=1))) if self.seq_names else None
        self.y = pin(torch.as_tensor(np.asarray(y)).float())
        self.n, self.batch_size, self.shuffle, self.drop_last = n, batch_size, shuffle, drop_last

    def __len__(self):
        return self.n // self.batch_size if self.drop_last else (self.n + self.batch_size - 1) // self.batch_size

    _CHUNK_BATCHES = 64  # shuffled batches gathered into one pinned chunk buffer at a time (two buffers, used alternately)

    def __iter__(self):
        mk = lambda sl: (PackedColumns(self.id_names, sl(self.ids), self.num_names, sl(self.nums), self.seq_names, sl(self.seqs)), sl(self.y))
        if not self.shuffle:
            for b in range(len(self)):
                lo, hi = b * self.batch_size, min((b + 1) * self.batch_size, self.n)
                yield mk(lambda t: None if t is None else t[lo:hi])
            return
        # Shuffled epochs.  Fancy-indexing a pinned array (`t[idx]`) lands in fresh PAGEABLE memory, and the trainer's non-blocking H2D
        # copy of it degrades to a staged, synchronous one.  The permuted rows are therefore gathered chunk-wise (64 batches) into one
        # of two PINNED chunk buffers and the batches are contiguous views of it — the same property the unshuffled path has.  A
        # buffer is refilled 64 batches after its last batch was handed out; copies out of it are long queued by then, and a device
        # synchronisation before the refill (once per 64 steps) makes sure they have completed whatever stream they run on.
        order = torch.randperm(self.n)
        nb, bs, R = len(self), self.batch_size, self._CHUNK_BATCHES
        srcs = [self.ids, self.nums, self.seqs, self.y]
        if getattr(self, "_chunks", None) is None:
            pinned = bool(self.y.is_pinned()) if hasattr(self.y, "is_pinned") else False
            rows = min(R * bs, self.n)
            self._chunks = [[None if t is None else torch.empty((rows,) + tuple(t.shape[1:]), dtype=t.dtype, pin_memory=pinned) for t in srcs] for _ in range(2)]
            self._chunk_used = [False, False]
        for c, b0 in enumerate(range(0, nb, R)):
            which = c & 1
            bufs = self._chunks[which]
            if self._chunk_used[which] and torch.cuda.is_available():
                torch.cuda.synchronize()  # H2D copies out of this buffer (handed out >= 64 batches ago) are done
            self._chunk_used[which] = True
            lo_c, hi_c = b0 * bs, min((b0 + R) * bs, self.n if not self.drop_last else nb * bs)
            idx = order[lo_c:hi_c]
            for t, buf in zip(srcs, bufs):
                if t is not None:
                    torch.index_select(t, 0, idx, out=buf[:idx.numel()])
            for b in range(b0, min(b0 + R, nb)):
                lo, hi = b * bs - lo_c, min((b + 1) * bs, hi_c) - lo_c
                ids_b, nums_b, seqs_b, y_b = (None if buf is None else buf[lo:hi] for buf in bufs)
                yield PackedColumns(self.id_names, ids_b, self.num_names, nums_b, self.seq_names, seqs_b), y_b
