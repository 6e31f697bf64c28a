"""Row-wise (lazy) optimisers for the embedding tables + the hybrid optimiser CTRTrainer uses in fast mode.

The reference hands ALL parameters, tables included, to a dense torch optimiser
(``trainers/ctr_trainer.py:60-61``): Adam over 26 x (1M x 16) floats moves ~11.6 GB per step whatever the
batch touched (SURVEY.md §7 hard part 2).  ``RowwiseOptimizer`` applies the same update rule to the rows the
batch touched, in ONE launch for all tables (``rh_fields_rowwise_update``), consuming and re-zeroing the dense
gradient buffer rows on the way (so no separate zero_grad work is left).

Semantic deviation (deliberate, opt-in): untouched rows do not decay / do not move by momentum — the
"lazy"/"sparse" Adam family (cf. ``torch.optim.SparseAdam``).  Gradients are identical to the reference's;
parity tests assert logits and ``weight.grad``, and the exact dense optimiser remains the default.
"""
import ctypes

import torch

from . import _lib, config, table as _table
from ._lib import RhField, check, stream_ptr

_KINDS = {torch.optim.SGD: 0, torch.optim.Adam: 1, torch.optim.Adagrad: 2}


def _ptrs(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


class RowwiseOptimizer(object):
    """Updates table rows touched since the last step.  ``params``: table weight Parameters (2-D, fp32, CUDA)."""

    def __init__(self, params, kind, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = list(params)
        self.kind = kind
        self.lr = float(lr)
        self.betas = (float(betas[0]), float(betas[1]))
        self.eps = float(eps)
        self.weight_decay = float(weight_decay)
        self.state = {}  # id(param) -> dict(m, v, stamp)
        dev = self.params[0].device
        self._step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self._lr_dev = torch.full((1,), self.lr, dtype=torch.float32, device=dev)
        self._lr_host = self.lr
        self._bc_dev = torch.ones(2, dtype=torch.float32, device=dev)  # Adam bias corrections, refreshed by rh_opt_advance

    # -- state ---------------------------------------------------------------------------------------
    def _state(self, p):
        st = self.state.get(id(p))
        if st is None:
            st = {"stamp": torch.zeros(p.shape[0], dtype=torch.int32, device=p.device)}
            V, D = p.shape
            if self.kind == 1 and D % 4 == 0:  # Adam: m and v interleaved, one (m | v) record of 2*D floats per row
                mv = torch.zeros((V, 2 * D), dtype=torch.float32, device=p.device)
                st["m"], st["v"], st["stride"] = mv[:, :D], mv[:, D:], 2 * D
            else:
                if self.kind in (1, 2):
                    st["m"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                if self.kind == 1:
                    st["v"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["stride"] = D
            self.state[id(p)] = st
        return st

    def set_lr(self, lr):
        lr = float(lr)
        if lr != self._lr_host:
            self._lr_host = lr
            self._lr_dev.fill_(lr)  # a device scalar: captured graphs see the new value on replay

    # -- step ----------------------------------------------------------------------------------------
    def step(self):
        self.advance()
        self.apply()

    def advance(self):
        """Step counter += 1 and the Adam bias corrections, on the device (shared with the dense optimiser).  A step that called
        ``advance_early`` only joins the side stream here."""
        early = getattr(self, "_early", None)
        if early is not None:
            self._early = None
            torch.cuda.current_stream(early.device).wait_stream(early)
            return
        check(_lib.lib().rh_opt_advance(self._step_dev.data_ptr(), self._bc_dev.data_ptr(), self.betas[0], self.betas[1], stream_ptr()), "rh_opt_advance")

    def advance_early(self):
        """Called by the trainers at the START of a step: the counter / bias-correction launch depends on nothing the step computes
        (only on the previous step's optimiser kernels having read the old values), so it runs on the side stream next to the
        forward instead of between the scatter and the row-wise update — one node less on the step's critical path (captured
        graphs keep the fork).  ``advance`` joins."""
        from . import ops
        dev = self.params[0].device
        if dev.type != "cuda" or not config.early_opt_advance or getattr(self, "_early", None) is not None:
            return
        cur, aux = torch.cuda.current_stream(dev), ops._aux_stream(dev)
        aux.wait_stream(cur)  # behind the previous step's rh_fields_rowwise_update / rh_dense_update (they read the old values)
        with torch.cuda.stream(aux):
            check(_lib.lib().rh_opt_advance(self._step_dev.data_ptr(), self._bc_dev.data_ptr(), self.betas[0], self.betas[1], stream_ptr()), "rh_opt_advance")
        self._early = aux

    def apply(self):
        """Update (and re-zero the gradient of) every row touched since the last step."""
        L = _lib.lib()
        st = stream_ptr()
        # one (table, id-list) entry per lookup recorded by the backward kernels
        entries = []
        for p in self.params:
            slot = _table.find_slot(p)
            if slot is None or slot.buffer is None or not slot.pending:
                continue
            if slot.all_dirty or p.grad is None or p.grad.data_ptr() != slot.buffer.data_ptr():
                raise RuntimeError("RowwiseOptimizer needs the engine's sparse-tracked gradient buffer; a dense gradient "
                                   "(e.g. an L1/L2 penalty on embedding tables) reached this table — use the dense optimiser "
                                   "(torch_rechub.b200.config.rowwise_optimizer = False) for that configuration")
            for ids, _ in slot.pending:
                entries.append((p, slot, ids))
        # group by (dim, batch) so that one launch covers all tables of a batch
        groups = {}
        for p, slot, ids in entries:
            if p.shape[1] % 4 == 0 and ids.dim() == 1:
                groups.setdefault((p.shape[1], ids.shape[0]), []).append((p, slot, ids))
            else:
                self._single(L, p, slot, ids, st)
        for (dim, batch), items in groups.items():
            for i in range(0, len(items), _lib.RH_MAX_FIELDS):
                chunk = items[i:i + _lib.RH_MAX_FIELDS]
                arr = (RhField * len(chunk))()
                tables, s1, s2, stamps = [], [], [], []
                for j, (p, slot, ids) in enumerate(chunk):
                    stt = self._state(p)
                    a = arr[j]
                    a.table = p.data_ptr()
                    a.table_grad = slot.buffer.data_ptr()
                    a.ids = ids.data_ptr()
                    a.id_stride = ids.stride(0) if ids.shape[0] > 1 else 1
                    a.ids_are_i32 = int(ids.dtype == torch.int32)
                    a.vocab = p.shape[0]
                    a.padding_idx = -1
                    a.tile_col = -1
                    a.fm_slot = -1
                    tables.append(p)
                    s1.append(stt.get("m"))
                    s2.append(stt.get("v"))
                    stamps.append(stt["stamp"])
                check(
                    L.rh_fields_rowwise_update(arr, len(chunk), dim, batch, _ptrs(tables), _ptrs(s1) if self.kind != 0 else None, _ptrs(s2) if self.kind == 1 else None, _ptrs(stamps), self.kind,
                                               self._step_dev.data_ptr(), self._lr_dev.data_ptr(), self._bc_dev.data_ptr(), self._state(chunk[0][0])["stride"], self.betas[0], self.betas[1], self.eps, self.weight_decay, st), "rh_fields_rowwise_update")
        for p in self.params:
            _table.mark_clean(p)

    def _single(self, L, p, slot, ids, st):
        stt = self._state(p)
        idc = ids.contiguous()
        m, v = stt.get("m"), stt.get("v")
        check(
            L.rh_rowwise_update(p.data_ptr(), slot.buffer.data_ptr(), None if m is None else m.data_ptr(), None if v is None else v.data_ptr(), stt["stamp"].data_ptr(), p.shape[0], p.shape[1], idc.data_ptr(),
                                int(idc.dtype == torch.int32), idc.numel(), self.kind, self._step_dev.data_ptr(), self._lr_dev.data_ptr(), self._bc_dev.data_ptr(), stt["stride"], self.betas[0], self.betas[1], self.eps, self.weight_decay, st),
            "rh_rowwise_update")

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                _table.clean(p)


class DenseOptimizer(object):
    """SGD / Adam / Adagrad over the small dense tensors in ONE launch (``rh_dense_update``), sharing the step counter,
    learning-rate scalar and Adam bias corrections of the row-wise optimiser (all on the device: graph-replay safe)."""

    def __init__(self, params, rowwise):
        self.params = [p for p in params]
        self.rw = rowwise
        self.state = {}

    def prepare(self):
        """Allocate the moment buffers of every parameter that has a gradient (on the caller's stream)."""
        rw = self.rw
        ps = [p for p in self.params if p.grad is not None]
        for p in ps:
            if id(p) not in self.state:
                self.state[id(p)] = (torch.zeros_like(p, memory_format=torch.contiguous_format) if rw.kind != 0 else None,
                                     torch.zeros_like(p, memory_format=torch.contiguous_format) if rw.kind == 1 else None)
        return ps

    def step(self):
        L = _lib.lib()
        rw = self.rw
        ps = self.prepare()
        if not ps:
            return
        grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in ps]
        n = len(ps)
        numel = (ctypes.c_int64 * n)(*[p.numel() for p in ps])
        s1 = _ptrs([self.state[id(p)][0] for p in ps]) if rw.kind != 0 else None
        s2 = _ptrs([self.state[id(p)][1] for p in ps]) if rw.kind == 1 else None
        check(
            L.rh_dense_update(n, _ptrs(ps), _ptrs(grads), s1, s2, numel, rw.kind, rw._lr_dev.data_ptr(), rw._bc_dev.data_ptr(), rw.betas[0], rw.betas[1], rw.eps, rw.weight_decay, stream_ptr()),
            "rh_dense_update")


class HybridOptimizer(object):
    """Row-wise optimiser for the tables + the user's torch optimiser for everything else.

    Quacks like a torch optimiser where CTRTrainer needs it (``step``, ``zero_grad``, ``param_groups``,
    ``state_dict``); schedulers attach to ``scheduler_target`` (the dense torch optimiser) and the tables follow
    its learning rate.
    """

    def __init__(self, rowwise, dense, dense_params):
        self.rowwise = rowwise
        self.dense = dense  # a torch optimiser that only carries param_groups / lr for schedulers; never stepped
        self.dense_engine = DenseOptimizer(dense_params, rowwise)
        self.scheduler_target = dense

    @classmethod
    def build(cls, model, optimizer_fn, optimizer_params, exclude=()):
        """None when the configuration has no row-wise counterpart (then the trainer stays on the dense optimiser).
        ``exclude``: ids of table weights that must stay on the dense side (tables replicated across ranks: their gradients are
        all-reduced as dense tensors, so every rank has to apply the same dense update)."""
        from .table import FieldTable
        kind = _KINDS.get(optimizer_fn)
        if kind is None:
            return None
        allowed = {"lr", "weight_decay", "betas", "eps"} if kind == 1 else ({"lr", "weight_decay"} if kind == 0 else {"lr", "weight_decay", "eps"})
        if not set(optimizer_params).issubset(allowed):
            return None  # momentum / amsgrad / lr_decay ...: keep exact dense semantics
        tables = []
        for mod in model.modules():
            if isinstance(mod, FieldTable) and mod.weight.is_cuda and mod.weight.requires_grad and mod.weight.dtype == torch.float32 and id(mod.weight) not in exclude:
                tables.append(mod.weight)
        if not tables:
            return None
        tset = {id(p) for p in tables}
        others = [p for p in model.parameters() if id(p) not in tset and p.numel() > 0 and p.requires_grad]
        if not others:
            return None
        dense = optimizer_fn(others, **dict(optimizer_params))
        defaults = {0: dict(lr=1e-3), 1: dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8), 2: dict(lr=1e-2, eps=1e-10)}[kind]
        kw = dict(defaults)
        kw.update({k: v for k, v in optimizer_params.items()})
        rw = RowwiseOptimizer(tables, kind, **kw)
        return cls(rw, dense, others)

    @property
    def param_groups(self):
        return self.dense.param_groups

    def step(self):
        lr = self.dense.param_groups[0]["lr"]
        self.rowwise.set_lr(float(lr) if not torch.is_tensor(lr) else float(lr.item()))
        self.rowwise.advance()  # the shared step counter / bias corrections first
        dev = self.dense_engine.params[0].device if self.dense_engine.params else None
        if config.concurrent_optimizers and dev is not None and dev.type == "cuda":
            # the touched-row update (random 64-B accesses, latency-bound) and the tower update (a few 100 k contiguous floats)
            # share nothing but the step counter: run them side by side
            from . import ops
            cur, aux = torch.cuda.current_stream(), ops._aux_stream(dev)
            self.dense_engine.prepare()
            aux.wait_stream(cur)
            with torch.cuda.stream(aux):
                self.dense_engine.step()
            self.rowwise.apply()
            cur.wait_stream(aux)
        else:
            self.rowwise.apply()
            self.dense_engine.step()

    def zero_grad(self, set_to_none=True):
        self.rowwise.zero_grad(set_to_none)
        self.dense.zero_grad(set_to_none)

    def state_dict(self):
        """Everything a resume needs (the wrapped torch optimiser is never stepped, its own state is empty): the step counter, the
        tables' row-wise moments + stamps and the tower's moments — keyed by POSITION in ``rowwise.params`` / ``dense_engine.params``
        (the order ``build`` derives from ``model.modules()`` / ``model.parameters()``), as torch optimisers key by index."""
        rw, de = self.rowwise, self.dense_engine
        rows = {}
        for i, p in enumerate(rw.params):
            st = rw.state.get(id(p))
            if st is not None:
                rows[i] = {k: st[k].detach().clone() for k in ("m", "v", "stamp") if st.get(k) is not None}
        dense = {}
        for i, p in enumerate(de.params):
            st = de.state.get(id(p))
            if st is not None:
                dense[i] = tuple(None if t is None else t.detach().clone() for t in st)
        return {"format": "rechub_b200.hybrid/1", "kind": rw.kind, "step": int(rw._step_dev.item()), "bias_corr": rw._bc_dev.detach().clone(), "rowwise": rows, "dense": dense,
                "param_groups": self.dense.state_dict()["param_groups"]}

    def load_state_dict(self, sd):
        if sd.get("format") != "rechub_b200.hybrid/1":
            raise ValueError("not a HybridOptimizer state_dict (a dense torch optimiser's checkpoint cannot seed the row-wise state)")
        rw, de = self.rowwise, self.dense_engine
        if sd["kind"] != rw.kind or max(list(sd["rowwise"]) + [-1]) >= len(rw.params) or max(list(sd["dense"]) + [-1]) >= len(de.params):
            raise ValueError("HybridOptimizer.load_state_dict: the checkpoint was written for another optimiser kind / parameter list")
        rw._step_dev.fill_(int(sd["step"]))
        rw._bc_dev.copy_(sd["bias_corr"])
        for i, ent in sd["rowwise"].items():
            st = rw._state(rw.params[i])
            for k, t in ent.items():
                if st.get(k) is None or st[k].shape != t.shape:
                    raise ValueError("HybridOptimizer.load_state_dict: row-wise state %r of table %d does not fit" % (k, i))
                st[k].copy_(t)
        for i, ent in sd["dense"].items():
            p = de.params[i]
            cur = de.state.get(id(p))
            if cur is None:
                cur = (torch.zeros_like(p, memory_format=torch.contiguous_format) if rw.kind != 0 else None, torch.zeros_like(p, memory_format=torch.contiguous_format) if rw.kind == 1 else None)
                de.state[id(p)] = cur
            for dst, src in zip(cur, ent):
                if dst is not None and src is not None:
                    dst.copy_(src)
        groups = sd.get("param_groups")
        if groups:
            for g, saved in zip(self.dense.param_groups, groups):
                g.update({k: v for k, v in saved.items() if k != "params"})
