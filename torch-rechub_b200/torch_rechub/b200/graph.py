"""CUDA-graph replay of CTRTrainer's training step.

At batch 4096 one DeepFM step is a few tens of microseconds of kernels: launch latency and Python dominate an
eager loop by an order of magnitude.  ``GraphedStep`` captures ``trainer._train_step`` (zero_grad, forward,
BCE, backward with the scatter-add, optimiser) once and replays it per batch; the batch is copied into static
input buffers first (one async H2D/D2D copy per column).

Host batches are software-pipelined: batch t+1 travels over PCIe on a copy stream into staging buffers while step t
computes; the step's own stream only pays three small D2D copies.  ``LaggedReader`` completes the pipeline on the
result side: the loss (and the out-of-range-id flag) of step t are read while step t+1 is already queued, so the
trainer's per-step ``loss.item()`` no longer drains the GPU.

Constraints: static shapes (a ragged last batch runs eagerly), the hybrid row-wise optimiser (its step
re-zeroes the gradient rows it consumed, so no per-step host bookkeeping is left), single process.
"""
import collections

import torch

from . import optim as _optim
from .data import PackedColumns

_WARMUP_STEPS = 3


class LaggedReader(object):
    """Device scalars -> host one step late.  ``push(loss)`` queues an async D2H of the loss and of the engine's error flag
    into pinned slots; ``pop()`` returns the oldest loss as a float (raising the reference's ``IndexError`` if that step
    saw an out-of-range id).  With one step of lag the host never waits for the step it has just launched."""

    def __init__(self, device, depth=4):
        self.device = torch.device(device)
        self.loss = torch.zeros(depth, dtype=torch.float32).pin_memory()
        self.flag = torch.zeros(depth, dtype=torch.int32).pin_memory()
        self.events = [torch.cuda.Event() for _ in range(depth)]
        self.depth, self.head = depth, 0
        self.queue = collections.deque()

    def pending(self):
        return len(self.queue)

    def reset(self):
        """Forget results of an interrupted loop."""
        self.queue.clear()

    def push(self, loss):
        from . import _lib
        if len(self.queue) == self.depth:
            raise RuntimeError("LaggedReader: %d results outstanding; pop() before pushing more" % self.depth)
        i = self.head
        self.head = (self.head + 1) % self.depth
        self.loss[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        self.flag[i:i + 1].copy_(_lib.err_flag(self.device), non_blocking=True)
        self.events[i].record()
        self.queue.append(i)

    def pop(self):
        from . import _lib
        i = self.queue.popleft()
        self.events[i].synchronize()
        v = int(self.flag[i])
        if v != 0:
            _lib.err_flag(self.device).zero_()
            self.queue.clear()
            if v == 0x7ffffff0:  # RH_ERRFLAG_SYNC_TIMEOUT
                raise RuntimeError("a cross-GPU hand-over of the sharded exchange timed out: a peer rank never published its step (crashed or diverged rank?)")
            raise IndexError("index out of range in self (embedding lookup, field #%d of the launch)" % (v - 1))
        return float(self.loss[i])


class GraphedStep(object):

    def __init__(self, trainer):
        self.trainer = trainer
        self.graph = None
        self.static_x = None
        self.static_y = None
        self.loss = None
        self.staging_x = self.staging_y = None  # landing buffers of the copy stream (host batches only)
        self.copy_stream = None
        self.ev_h2d = self.ev_consumed = None
        self._pf_cols = None  # (id column name, table weight) pairs of the next-batch L2 prefetch
        self.calls = 0
        self.stream = None  # warm-up AND capture run on this side stream (autograd binds AccumulateGrad nodes to the stream
        # they were first used on; a default-stream node inside a capture invalidates it)
        self.enabled = isinstance(trainer.optimizer, _optim.HybridOptimizer)
        if not self.enabled:
            print("[rechub-b200] cuda_graph needs config.rowwise_optimizer (dense optimisers run eagerly)")

    def _signature(self, x_dict, y):
        if isinstance(x_dict, PackedColumns):  # three packed tensors describe all the columns: keep the per-step host cost flat
            return tuple((None if t is None else (tuple(t.shape), t.dtype)) for t in (x_dict.ids, x_dict.nums, x_dict.seqs)) + (tuple(y.shape), y.dtype)
        return tuple((k, tuple(v.shape), v.dtype) for k, v in x_dict.items()) + (tuple(y.shape), y.dtype)

    def _capture(self, x_dict, y):
        dev = self.trainer.device
        if isinstance(x_dict, PackedColumns):
            self.static_x = x_dict.to(dev).clone() if not x_dict.ids.is_cuda else x_dict.clone()
        else:
            self.static_x = {k: v.to(dev).clone() for k, v in x_dict.items()}
        self.static_y = y.to(dev).float().clone()
        self.sig = self._signature(x_dict, y)
        self.graph = torch.cuda.CUDAGraph()
        torch.cuda.current_stream().synchronize()
        with torch.cuda.graph(self.graph, stream=self._side_stream()):
            self.loss = self.trainer._train_step(self.static_x, self.static_y).detach()

    @staticmethod
    def _copy_batch(src_x, src_y, dst_x, dst_y):
        if isinstance(src_x, PackedColumns) and isinstance(dst_x, PackedColumns):
            pairs = [(a, b) for a, b in ((src_x.ids, dst_x.ids), (src_x.nums, dst_x.nums), (src_x.seqs, dst_x.seqs), (src_y, dst_y)) if a is not None]
            if all(a.is_cuda and b.is_cuda and a.device == b.device and a.dtype == b.dtype and a.shape == b.shape and a.is_contiguous() and b.is_contiguous() for a, b in pairs):
                import ctypes
                from . import _lib
                n = len(pairs)
                dst = (ctypes.c_void_p * n)(*[b.data_ptr() for _, b in pairs])
                src = (ctypes.c_void_p * n)(*[a.data_ptr() for a, _ in pairs])
                nbytes = (ctypes.c_int64 * n)(*[a.numel() * a.element_size() for a, _ in pairs])
                _lib.check(_lib.lib().rh_copy_segments(n, dst, src, nbytes, _lib.stream_ptr()), "rh_copy_segments")  # one launch instead of <= 4 copies
                return
            src_x.copy_into(dst_x)
        else:
            for k, v in src_x.items():
                dst_x[k].copy_(v, non_blocking=True)
        dst_y.copy_(src_y, non_blocking=True)

    def load_inputs(self, x_dict, y):
        """Batch -> static buffers.  Device batches are copied directly.  Host batches go over PCIe on the copy stream into
        staging buffers — overlapping the step that is still running — and reach the static buffers by D2D copies."""
        from . import config
        probe = x_dict.ids if isinstance(x_dict, PackedColumns) and x_dict.ids is not None else next(iter(x_dict.values()))
        if probe.is_cuda or not config.pipelined_inputs:
            self._copy_batch(x_dict, y, self.static_x, self.static_y)
            return
        if self.staging_x is None:
            self.staging_x = self.static_x.clone() if isinstance(self.static_x, PackedColumns) else {k: v.clone() for k, v in self.static_x.items()}
            self.staging_y = self.static_y.clone()
            self.copy_stream = torch.cuda.Stream(device=self.trainer.device)
            self.ev_h2d, self.ev_consumed = torch.cuda.Event(), torch.cuda.Event()
            self.ev_consumed.record()
        cur, cs = torch.cuda.current_stream(), self.copy_stream
        cs.wait_event(self.ev_consumed)  # the previous batch has left the staging buffers
        with torch.cuda.stream(cs):
            self._copy_batch(x_dict, y, self.staging_x, self.staging_y)
            self.ev_h2d.record(cs)
            if config.next_batch_prefetch:
                self._prefetch_rows(self.staging_x)  # on the copy stream, behind the H2D copy, while the previous step computes
        cur.wait_event(self.ev_h2d)
        self._copy_batch(self.staging_x, self.staging_y, self.static_x, self.static_y)
        self.ev_consumed.record(cur)

    def _prefetch_rows(self, staged):
        """``rh_fields_prefetch`` for every 1-D sparse id column of the staged batch: table rows, their gradient rows and their
        interleaved Adam records into L2, one step before the kernels that touch them."""
        import ctypes
        from . import _lib, table as _table
        from ..basic.features import SparseFeature
        from ..basic.layers import EmbeddingLayer
        if self._pf_cols is None:
            cols, seen = [], set()
            for mod in self.trainer.model.modules():
                if isinstance(mod, EmbeddingLayer):
                    for fea in mod.features:
                        if isinstance(fea, SparseFeature) and fea.name in staged and fea.name not in seen:
                            w = mod.table_of(fea).weight
                            if w.is_cuda and w.dim() == 2 and w.numel() > 0 and w.shape[1] % 4 == 0:
                                seen.add(fea.name)
                                cols.append((fea.name, w))
            self._pf_cols = cols
        by_dim = {}
        for name, w in self._pf_cols:
            ids = staged[name]
            if ids.dim() == 1 and ids.dtype in (torch.int64, torch.int32):
                by_dim.setdefault(w.shape[1], []).append((ids, w))
        rw = getattr(self.trainer.optimizer, "rowwise", None)
        L, st = _lib.lib(), _lib.stream_ptr()
        for dim, items in by_dim.items():
            for i in range(0, len(items), _lib.RH_MAX_FIELDS):
                chunk = items[i:i + _lib.RH_MAX_FIELDS]
                arr = (_lib.RhField * len(chunk))()
                states = (ctypes.c_void_p * len(chunk))()
                stride = 0
                for j, (ids, w) in enumerate(chunk):
                    slot = _table.find_slot(w)
                    a = arr[j]
                    a.table = w.data_ptr()
                    a.table_grad = slot.buffer.data_ptr() if slot is not None and slot.buffer is not None else None
                    a.ids = ids.data_ptr()
                    a.id_stride = ids.stride(0) if ids.shape[0] > 1 else 1
                    a.ids_are_i32 = int(ids.dtype == torch.int32)
                    a.vocab = w.shape[0]
                    a.padding_idx, a.tile_col, a.fm_slot = -1, -1, -1
                    rec = rw.state.get(id(w)) if rw is not None else None
                    if rec is not None and rec.get("stride") == 2 * dim:  # interleaved (m | v) records
                        states[j] = rec["m"].data_ptr()
                        stride = 2 * dim
                _lib.check(L.rh_fields_prefetch(arr, len(chunk), dim, chunk[0][0].shape[0], states if stride else None, stride, st), "rh_fields_prefetch")

    def _side_stream(self):
        if self.stream is None:
            self.stream = torch.cuda.Stream(device=self.trainer.device)
        return self.stream

    def _eager(self, x_dict, y):
        dev = self.trainer.device
        x_dict = x_dict.to(dev) if isinstance(x_dict, PackedColumns) else {k: v.to(dev) for k, v in x_dict.items()}
        y = y.to(dev).float()
        if not self.enabled:
            return self.trainer._train_step(x_dict, y).detach()
        cur, side = torch.cuda.current_stream(), self._side_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            loss = self.trainer._train_step(x_dict, y).detach()
        cur.wait_stream(side)
        return loss

    def __call__(self, x_dict, y):
        self.calls += 1
        if not self.enabled or self.calls <= _WARMUP_STEPS:
            return self._eager(x_dict, y)  # eager warm-up on real batches (also builds optimiser state)
        if self.graph is None:
            self._capture(x_dict, y)  # records only; the replay below executes this batch exactly once
        elif self._signature(x_dict, y) != self.sig:
            return self._eager(x_dict, y)  # ragged batch
        else:
            self.load_inputs(x_dict, y)
        opt = self.trainer.optimizer
        lr = opt.dense.param_groups[0]["lr"]
        opt.rowwise.set_lr(float(lr))  # outside the graph: a scheduler may have moved it
        self.graph.replay()
        return self.loss
