"""Embedding tables and their persistent gradient buffers.

``FieldTable`` is the ``nn.Embedding`` every feature descriptor creates (reference
``basic/initializers.py:16-21``).  It stays an ``nn.Embedding`` (state_dict key ``<name>.weight``,
``isinstance`` checks of ``basic/loss_func.py:47``) and adds the engine behaviour for CUDA weights:

* lookups run through ``rh_rows_gather`` / ``rh_fields_fwd``;
* the backward pass scatter-adds row gradients into ONE persistent dense ``(vocab, dim)`` buffer that is
  handed to autograd as ``weight.grad`` — the same dense gradient the reference gets from
  ``aten::embedding_dense_backward``, without allocating and zero-filling
  ``vocab x dim`` floats per lookup per step (66 % of the reference's CPU step).  The buffer is cleaned
  SPARSELY: only the rows dirtied by the previous step are re-zeroed.
"""
import weakref

import torch
import torch.nn as nn

from . import _lib, config


class GradSlot(object):
    """Engine-side gradient state of one table weight."""
    __slots__ = ("buffer", "pending", "all_dirty", "__weakref__")

    def __init__(self):
        self.buffer = None  # (vocab, dim) fp32, zero outside the rows listed in `pending`
        self.pending = []  # [(ids tensor (any shape, int64/int32), is_snapshot)] rows dirtied since last clean
        self.all_dirty = False  # a foreign dense gradient was accumulated into the buffer


_slots = {}  # id(weight Parameter) -> (weakref to the Parameter, GradSlot); tensors cannot key a WeakKeyDictionary (== is elementwise)


def find_slot(weight):
    ent = _slots.get(id(weight))
    if ent is not None and ent[0]() is weight:
        return ent[1]
    return None


def slot_of(weight):
    s = find_slot(weight)
    if s is None:
        s = GradSlot()
        key = id(weight)
        _slots[key] = (weakref.ref(weight, lambda _r, _k=key: _slots.pop(_k, None)), s)
        # A dense gradient reaching this weight through autograd (e.g. an L2 penalty on the table) is added IN PLACE to
        # whatever .grad is — possibly our buffer — and may touch every row: remember to clean everything next time.
        # (A tensor hook receives None when only the engine's kernels contributed, the summed gradient otherwise.)
        def _mark(grad, _slot_ref=weakref.ref(s)):
            sl = _slot_ref()
            if grad is not None and sl is not None:
                sl.all_dirty = True

        if weight.requires_grad:
            weight.register_hook(_mark)
    return s


def _ensure_buffer(weight, slot):
    w = weight.detach()
    if slot.buffer is None or slot.buffer.shape != w.shape or slot.buffer.device != w.device:
        slot.buffer = torch.zeros_like(w, memory_format=torch.contiguous_format)
        slot.pending = []
        slot.all_dirty = False
    return slot.buffer


def clean(weight, slot=None):
    """Re-zero the rows dirtied since the last clean (sparse ``zero_grad``)."""
    slot = slot or slot_of(weight)
    if slot.buffer is None:
        return
    if slot.all_dirty:
        slot.buffer.zero_()
    else:
        L = _lib.lib()
        vocab, dim = slot.buffer.shape
        for ids, _ in slot.pending:
            if not ids.is_contiguous():
                ids = ids.contiguous()  # a column view of a packed (B, F) id block: rh_rows_zero takes a dense id list
            _lib.check(L.rh_rows_zero(slot.buffer.data_ptr(), vocab, dim, ids.data_ptr(), int(ids.dtype == torch.int32), ids.numel(), _lib.stream_ptr()), "rh_rows_zero")
    slot.pending = []
    slot.all_dirty = False


def grad_target(weight):
    """Dense gradient tensor the backward kernels must scatter into, attached as ``weight.grad``.

    Returns ``(tensor, slot_or_None)``; ``slot`` is None when ``weight.grad`` is a foreign dense tensor
    (then the kernels add straight into it and nothing is tracked).
    """
    if not weight.requires_grad:
        return None, None
    slot = slot_of(weight)
    g = weight.grad
    if g is not None and (slot.buffer is None or g.data_ptr() != slot.buffer.data_ptr()):
        if g.is_sparse or g.shape != weight.shape or not g.is_contiguous():
            raise RuntimeError("table weight has an incompatible .grad (sparse or strided); call zero_grad() first")
        return g, None
    buf = _ensure_buffer(weight, slot)
    if g is None:
        if slot.pending or slot.all_dirty:
            clean(weight, slot)
        weight.grad = buf
    return buf, slot


_MAX_PENDING = 64  # id lists remembered for the sparse re-zero before it degrades to one full zero_()


def note_dirty(slot, ids):
    """Record that the rows ``ids`` of the slot's buffer now hold gradient.

    The id list is SNAPSHOT unless the row-wise optimiser owns the step (it consumes and re-zeroes the rows inside the same step,
    ``mark_clean``): with a dense ``torch.optim`` optimiser the list is read again at the NEXT ``zero_grad`` — by then a caller that
    reuses its device id buffers (``PackedLoader`` staging, a captured step's static inputs) has overwritten it, and the sparse
    clean would zero the wrong rows and leave stale gradients behind.  ``zero_grad(set_to_none=False)`` never reaches ``clean``;
    the list is bounded: past ``_MAX_PENDING`` entries the slot is marked all-dirty (one full ``zero_()`` at the next clean).
    """
    if slot is None or slot.all_dirty:
        return
    if not config.rowwise_optimizer and len(slot.pending) >= _MAX_PENDING:  # (the row-wise optimiser drains the list every step)
        slot.pending = []
        slot.all_dirty = True
        return
    if config.static_inputs or not config.rowwise_optimizer:
        slot.pending.append((ids.detach().clone(), True))
    else:
        slot.pending.append((ids.detach(), False))


def mark_clean(weight):
    """Called by the row-wise optimiser: it consumed AND re-zeroed every pending row."""
    slot = find_slot(weight)
    if slot is not None:
        slot.pending = []


class FieldTable(nn.Embedding):
    """``nn.Embedding`` whose CUDA lookups / gradients go through the sm_100a engine.

    CPU weights behave exactly like ``nn.Embedding`` (dense ``weight.grad`` from autograd).
    Options of ``nn.Embedding`` that the reference never sets (``max_norm``, ``scale_grad_by_freq``,
    ``sparse``) are honoured on CPU and rejected on CUDA.
    """

    def forward(self, ids):
        if not self.weight.is_cuda:
            return super().forward(ids)
        if self.max_norm is not None or self.scale_grad_by_freq or self.sparse:
            raise NotImplementedError("FieldTable on CUDA supports the reference's nn.Embedding configuration only " "(max_norm=None, scale_grad_by_freq=False, sparse=False)")
        from . import ops
        return ops.table_lookup(self, ids)
