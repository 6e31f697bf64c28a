"""CUDA-graph replay of CTRTrainer's training step.

At batch 4096 one DeepFM step is a few tens of microseconds of kernels: launch latency and Python dominate an
eager loop by an order of magnitude.  ``GraphedStep`` captures ``trainer._train_step`` (zero_grad, forward,
BCE, backward with the scatter-add, optimiser) once and replays it per batch; the batch is copied into static
input buffers first (one async H2D/D2D copy per column).

Constraints: static shapes (a ragged last batch runs eagerly), the hybrid row-wise optimiser (its step
re-zeroes the gradient rows it consumed, so no per-step host bookkeeping is left), single process.
"""
import torch

from . import optim as _optim
from .data import PackedColumns

_WARMUP_STEPS = 3


class GraphedStep(object):

    def __init__(self, trainer):
        self.trainer = trainer
        self.graph = None
        self.static_x = None
        self.static_y = None
        self.loss = None
        self.calls = 0
        self.stream = None  # warm-up AND capture run on this side stream (autograd binds AccumulateGrad nodes to the stream
        # they were first used on; a default-stream node inside a capture invalidates it)
        self.enabled = isinstance(trainer.optimizer, _optim.HybridOptimizer)
        if not self.enabled:
            print("[rechub-b200] cuda_graph needs config.rowwise_optimizer (dense optimisers run eagerly)")

    def _signature(self, x_dict, y):
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

    def load_inputs(self, x_dict, y):
        """Batch -> static buffers (straight from pinned host memory when the batch is still on the host)."""
        if isinstance(x_dict, PackedColumns) and isinstance(self.static_x, PackedColumns):
            x_dict.copy_into(self.static_x)
        else:
            for k, v in x_dict.items():
                self.static_x[k].copy_(v, non_blocking=True)
        self.static_y.copy_(y, non_blocking=True)

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
