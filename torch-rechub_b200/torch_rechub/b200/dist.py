"""Multi-GPU: one process per GPU, embedding tables sharded BY FEATURE FIELD, dense tower data-parallel.

Replaces the reference's single-process ``nn.DataParallel`` (``trainers/ctr_trainer.py:53-55``), which re-broadcasts every
table on every forward (1.66 GB/step at Criteo shape).  Here:

  forward   ids all-to-all  ->  owner-local fused gather over the GLOBAL batch (its fields only)
            ->  rows all-to-all back to the sample's rank  ->  tile in the model's column order
  backward  tile-gradient all-to-all to the owners -> owner-local scatter-add into its tables;
            ONE all-reduce of the flattened dense-parameter gradients (+ the loss scalar riding in the same bucket)

Semantics kept from DataParallel: every rank's BatchNorm uses its own sub-batch statistics, and the loss is the mean
over the GLOBAL batch (each rank back-propagates ``local_mean / world``; dense gradients are summed).
Collectives are ``torch.distributed`` (NCCL over NVLink on GPUs; gloo on CPU for the tests) — all static-shaped, so the
whole step stays CUDA-graph capturable.
"""
import os

import torch
import torch.distributed as dist

from ..basic.features import DenseFeature, SequenceFeature, SparseFeature


def field_owners(names, world):
    """Owner rank of every table name: round-robin in list order (26 fields / 8 ranks -> 4,4,3,3,3,3,3,3)."""
    return {n: i % world for i, n in enumerate(names)}


class _AllToAllRows(torch.autograd.Function):
    """out[s] = what rank s sent me; equal splits.  x: (world, rows, width) -> (world, rows, width).  Backward = the same exchange.
    ``anchor`` is a requires-grad scalar that makes autograd run this backward on EVERY rank — also on a rank that owns no
    trainable table (its ``x`` carries no gradient), which would otherwise skip the collective the other ranks are waiting in."""

    @staticmethod
    def forward(ctx, x, group, anchor):
        ctx.group = group
        x = x.contiguous()
        out = torch.empty_like(x)
        dist.all_to_all_single(out, x, group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        out = torch.empty_like(g)
        dist.all_to_all_single(out, g, group=ctx.group)
        return out, None, None


class P2PExchange(object):
    """Peer-mapped (symmetric-memory) buffers of one sharded front-end plan + the cross-GPU barrier that orders them.

    ids   (fmax, W, b) int64   on owner r: [k, s, :] = ids of rank s's samples for r's k-th field (written by rank s;
                               sample-major (W, b, fmax) when config.p2p_field_major_ids is off)
    rows  (W, b, fmax*dim) f32 on rank s : block r = rows of owner r's fields for s's samples     (written by owner r's gather kernel)
    drows (W, b, fmax*dim) f32 on owner r: block s = gradients of those rows from rank s          (RED by rank s's backward kernel;
                               absent with direct gradients: the REDs then target the owner's table gradient buffers)
    """

    def __init__(self, group, device, W, b, fmax, dim, with_drows=True):
        import torch.distributed._symmetric_memory as symm
        from . import config
        self.field_major = bool(config.p2p_field_major_ids)
        self.ids = symm.empty(W * b * fmax, dtype=torch.int64, device=device)
        self.rows = symm.empty(W * b * fmax * dim, dtype=torch.float32, device=device)
        self.h_ids = symm.rendezvous(self.ids, group)
        self.h_rows = symm.rendezvous(self.rows, group)
        self.ids_ptrs = [int(p) for p in self.h_ids.buffer_ptrs]
        self.rows_ptrs = [int(p) for p in self.h_rows.buffer_ptrs]
        self.drows = self.drows_ptrs = None
        if with_drows:
            self.drows = symm.empty(W * b * fmax * dim, dtype=torch.float32, device=device)
            self.h_drows = symm.rendezvous(self.drows, group)
            self.drows_ptrs = [int(p) for p in self.h_drows.buffer_ptrs]
            self.drows.zero_()
        self.ids_local = torch.zeros_like(self.ids)
        self.own = None
        if config.p2p_own_barrier and W <= 8:
            import ctypes
            flags = symm.empty(8, dtype=torch.int32, device=device)
            flags.zero_()
            h_flags = symm.rendezvous(flags, group)
            self.own = {"flags": flags, "handle": h_flags, "ptrs": (ctypes.c_void_p * W)(*[int(p) for p in h_flags.buffer_ptrs]), "epoch": torch.zeros(1, dtype=torch.int32, device=device),
                        "rank": dist.get_rank(group), "world": W}
            torch.cuda.synchronize(device)
            dist.barrier(group=group)  # nobody publishes into flags that are still being zeroed
        # hand-overs folded into the launches (config.p2p_fused_sync): per-phase flags in peer memory + the exchange's step counter
        self.fs = None
        if config.p2p_fused_sync and W <= 8:
            import ctypes
            fl = symm.empty(16, dtype=torch.int32, device=device)  # [0:8] id phase, [8:16] row phase: flags[s] = last step rank s published
            fl.zero_()
            h_fl = symm.rendezvous(fl, group)
            base = [int(p) for p in h_fl.buffer_ptrs]
            self.fs = {"flags": fl, "handle": h_fl, "ids_ptrs": (ctypes.c_void_p * W)(*base), "rows_ptrs": (ctypes.c_void_p * W)(*[p + 32 for p in base]),
                       "step": torch.zeros(1, dtype=torch.int32, device=device), "tickets": torch.zeros(2, dtype=torch.int32, device=device), "rank": dist.get_rank(group)}
            torch.cuda.synchronize(device)
            dist.barrier(group=group)

    def barrier(self):
        if self.own is not None:
            from . import _lib
            o = self.own
            _lib.check(_lib.lib().rh_peer_barrier(o["ptrs"], o["flags"].data_ptr(), o["rank"], o["world"], o["epoch"].data_ptr(), _lib.stream_ptr()), "rh_peer_barrier")
        else:
            self.h_rows.barrier(channel=0)


def _snapshot(front, ids):
    """The owner's id views alias the exchange's id buffer, which the NEXT forward overwrites before a dense optimiser's lazy
    sparse clean reads it: record a copy then.  The row-wise optimiser consumes (and re-zeroes) the rows inside the same step."""
    return ids.clone() if front.lazy_clean else ids


class _ShardedP2P(torch.autograd.Function):
    """The sharded front end as ONE autograd node, all exchanges done by the engine's own kernels over NVLink peer memory:

      forward   rh_ids_scatter (ids -> owners)  | barrier | rh_fields_fwd_p2p (owner gathers, rows land in the samples' GPUs)
                | barrier | rh_fields_fwd on the received rows (unpack + dense columns + FM + LR)
      backward  direct gradients (default): rh_fields_bwd REDs every row gradient over NVLink into the OWNER's gradient buffer
                at the row id | barrier (config.p2p_defer_barrier: issued by DistEngine.train_step after the all-reduce launch)
                staged gradients: rh_fields_bwd REDs into the owner's staging buffer | barrier | rh_fields_bwd on the owner
                (scatter-add into its tables) | re-zero the staging buffer
    """

    @staticmethod
    def forward(ctx, front, p, x, fm_features, lr_w, lr_b, anchor, *owned_weights):
        from . import _lib, ops, table as _table
        L = _lib.lib()
        ex, W, me = p["ex"], front.world, front.rank
        sparse, dense, fmax, dim = p["sparse"], p["dense"], p["fmax"], p["dim"]
        b = p["batch"]
        dev = front.device
        st = ops.stream_ptr()
        err = _lib.err_flag(dev).data_ptr()
        import ctypes
        # F1: my samples' ids -> the owners
        cols = (ops.RhField * len(p["slots"]))()  # slots are sorted by (owner, slot)
        col_dest = (ctypes.c_int32 * len(p["slots"]))()
        my_ids = {}
        for n, (r, k, f) in enumerate(p["slots"]):
            ids = my_ids[f.name] = ops._as_ids(x[f.name])
            cols[n].ids = ids.data_ptr()
            cols[n].id_stride = ids.stride(0) if ids.shape[0] > 1 else 1
            cols[n].ids_are_i32 = int(ids.dtype == torch.int32)
            col_dest[n] = r
        direct = front.direct if any(ctx.needs_input_grad) else None  # peer-mapped gradient buffers of every table (None: staged route / no backward)
        if direct is not None:
            # my own tables' buffers must be attached and clean BEFORE any peer's backward REDs into them: do it ahead of the
            # forward barriers (the staged route does this in its owner-side backward pass)
            for f in p["by_owner"][me]:
                _table.grad_target(front.layer.table_of(f).weight)
        fs = ex.fs
        if ex.field_major:  # [slot, source rank, sample]
            bases = (ctypes.c_void_p * W)(*[ex.ids_ptrs[r] + me * b * 8 for r in range(W)])
            scatter_args = (cols, len(p["slots"]), col_dest, b, bases, W, fmax, W * b, 1)
        else:  # [source rank, sample, slot]
            bases = (ctypes.c_void_p * W)(*[ex.ids_ptrs[r] + me * b * fmax * 8 for r in range(W)])
            scatter_args = (cols, len(p["slots"]), col_dest, b, bases, W, fmax, 1, fmax)
        mine = p["by_owner"][me]
        if fs is not None:
            # F1 + hand-over: the last CTA publishes "rank me's ids of this step have landed" to every owner
            ops.check(L.rh_ids_scatter_signal(*scatter_args, fs["ids_ptrs"], me, W, fs["step"].data_ptr(), fs["tickets"].data_ptr(), st), "rh_ids_scatter_signal")
        else:
            ops.check(L.rh_ids_scatter(*scatter_args, st), "rh_ids_scatter")
            ex.barrier()
            ex.ids_local.copy_(ex.ids)
        # F3: owner-side gather over the global batch, rows stored straight into the destination GPUs' tiles.  The owner keeps a
        # LOCAL copy of the ids: peers may refill ex.ids for the next step while this rank's backward / optimiser still need them
        # (fused hand-overs: written by the gather itself, next to the reads; otherwise copied after the barrier above).
        view = (lambda t: t.view(fmax, W * b)) if ex.field_major else (lambda t: t.view(W * b, fmax).t())
        ids_g = view(ex.ids_local)
        orefs = []
        for k, f in enumerate(mine):
            tbl = front.layer.table_of(f)
            orefs.append(ops.FieldRef(tbl.weight, ids_g[k], tbl.padding_idx, k * dim, -1))
        if orefs:
            dest = (ctypes.c_void_p * W)(*[ex.rows_ptrs[s] + me * b * fmax * dim * 4 for s in range(W)])
            if fs is not None:
                ids_in = view(ex.ids)
                arr = ops._field_array([ops.FieldRef(r.weight, ids_in[k], None, k * dim, -1) for k, r in enumerate(orefs)])
                sy = _lib.RhSync()
                sy.wait_flags, sy.wait_mask, sy.step = fs["flags"].data_ptr(), (1 << W) - 1, fs["step"].data_ptr()  # every rank's ids
                sy.sig_flags, sy.sig_world, sy.sig_rank = ctypes.cast(fs["rows_ptrs"], ctypes.c_void_p), W, me  # -> "owner me's rows have landed"
                sy.ticket = fs["tickets"].data_ptr() + 4
                sy.id_snapshot_delta = ex.ids_local.data_ptr() - ex.ids.data_ptr()
                ops.check(L.rh_fields_fwd_sync(arr, len(orefs), dim, None, 0, W * b, None, fmax * dim, None, None, None, None, None, dest, W, b, ctypes.byref(sy), err, st), "rh_fields_fwd_sync (gather)")
            else:
                ops.check(L.rh_fields_fwd_p2p(ops._field_array(orefs), len(orefs), dim, W * b, dest, W, b, fmax * dim, err, st), "rh_fields_fwd_p2p")
        if fs is None:
            ex.barrier()
        # F5: sample side — the received rows are the "table"
        table = ex.rows.view(W * b * fmax, dim)
        fm_slot = {f.name: j for j, f in enumerate(fm_features)} if fm_features else {}
        srefs, col = [], 0
        for f, rid in zip(sparse, p["row_ids"]):
            srefs.append(ops.FieldRef(table, rid, None, col, fm_slot.get(f.name, -1), is_act=True))
            col += dim
        drefs = []
        for f in dense:
            v = x[f.name]
            if v.dtype not in ops._DENSE_CODES:
                v = v.float()
            width = 1 if v.dim() == 1 else v.shape[1]
            drefs.append(ops.DenseRef(v, width, col))
            col += width
        width_total = col
        ld = ops._pad4(width_total)
        tile = torch.empty((b, ld), dtype=torch.float32, device=dev)
        want_fm = bool(fm_features)
        y_fm = torch.empty(b, dtype=torch.float32, device=dev) if want_fm else None
        y_lr = torch.empty(b, dtype=torch.float32, device=dev) if (want_fm and lr_w is not None) else None
        fsum = torch.empty((b, dim), dtype=torch.float32, device=dev) if want_fm else None
        fwd_args = (ops._field_array(srefs), len(srefs), dim, ops._dense_array(drefs) if drefs else None, len(drefs), b, tile.data_ptr(), ld, ops.ptr(lr_w) if y_lr is not None else None,
                    ops.ptr(lr_b) if y_lr is not None else None, ops.ptr(y_fm), ops.ptr(y_lr), ops.ptr(fsum))
        if fs is not None:
            sy = _lib.RhSync()  # wait inside the launch for the rows of every owner that has fields
            sy.wait_flags, sy.step = fs["flags"].data_ptr() + 32, fs["step"].data_ptr()
            sy.wait_mask = sum(1 << r for r in range(W) if p["by_owner"][r])
            ops.check(L.rh_fields_fwd_sync(*fwd_args, None, 0, 0, ctypes.byref(sy), err, st), "rh_fields_fwd_sync (unpack)")
        else:
            ops.check(L.rh_fields_fwd(*fwd_args, err, st), "rh_fields_fwd")
        ctx.front, ctx.p, ctx.srefs, ctx.orefs, ctx.ld, ctx.want_lr = front, p, srefs, orefs, ld, y_lr is not None
        ctx.direct, ctx.my_ids = direct, my_ids
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(tile, fsum, lr_w)
        return (tile if ld == width_total else tile[:, :width_total]), y_fm, y_lr

    @staticmethod
    def backward(ctx, d_tile, d_yfm, d_ylr):
        from . import _lib, ops, table as _table
        L = _lib.lib()
        front, p = ctx.front, ctx.p
        ex, W, me = p["ex"], front.world, front.rank
        fmax, dim, b = p["fmax"], p["dim"], p["batch"]
        tile, fsum, lr_w = ctx.saved_tensors
        dev = front.device
        st = ops.stream_ptr()
        err = _lib.err_flag(dev).data_ptr()
        d_ld = 0
        if d_tile is not None:
            d_tile = ops._rowmajor(d_tile)
            d_ld = d_tile.stride(0) if b > 1 else d_tile.shape[1]
        d_yfm = d_yfm.contiguous() if d_yfm is not None else None
        d_ylr = d_ylr.contiguous() if d_ylr is not None else None
        d_lrw = d_lrb = None
        if d_ylr is not None and lr_w is not None:
            n = lr_w.numel()
            buf = torch.zeros(ops._pad4(n) + 1, dtype=torch.float32, device=dev)
            d_lrw, d_lrb = buf[:n].view_as(lr_w), buf[ops._pad4(n):ops._pad4(n) + 1]
        arr = ops._field_array(ctx.srefs)
        if ctx.direct is not None:
            # B1 (direct): row gradients go over NVLink straight into the OWNER's gradient buffer, at the row the id names.
            # The rows themselves come from the saved tile, so `table` is never dereferenced here.
            for n, f in enumerate(p["sparse"]):
                ids = ctx.my_ids[f.name]
                ptr_, vocab, pad = ctx.direct[f.name if f.shared_with is None else f.shared_with]
                arr[n].table = tile.data_ptr()
                arr[n].table_grad = ptr_
                arr[n].ids = ids.data_ptr()
                arr[n].id_stride = ids.stride(0) if ids.shape[0] > 1 else 1
                arr[n].ids_are_i32 = int(ids.dtype == torch.int32)
                arr[n].vocab = vocab
                arr[n].padding_idx = pad
        else:
            # B1 (staged): gradients of the received rows go to their owners' staging buffers (vector RED into drows[owner][me])
            owners = [front.owner_of(f) for f in p["sparse"]]
            for n, r in enumerate(owners):
                # the kernel indexes the gradient buffer with the FORWARD row id (r*b + i)*fmax + k; the owner's block for my
                # samples starts at row me*b*fmax: shift the base so that both agree
                arr[n].table_grad = ex.drows_ptrs[r] + (me - r) * b * fmax * dim * 4
        has_fm = (d_yfm is not None or d_ylr is not None) and any(s.fm_slot >= 0 for s in ctx.srefs)
        ops.check(
            L.rh_fields_bwd(arr, len(ctx.srefs), dim, b, tile.data_ptr(), ctx.ld, ops.ptr(d_tile), d_ld, ops.ptr(d_yfm) if has_fm else None, ops.ptr(d_ylr) if has_fm else None,
                            ops.ptr(lr_w) if has_fm else None, ops.ptr(fsum) if has_fm else None, ops.ptr(d_lrw) if has_fm else None, ops.ptr(d_lrb) if has_fm else None, err, st), "rh_fields_bwd")
        if ctx.direct is not None and front.defer_barrier:
            front.deferred = ex  # the step runs this barrier after launching the dense all-reduce, next to the row-wise update
        else:
            ex.barrier()
        if ctx.direct is not None:
            # once every rank's REDs have landed in my buffers only the bookkeeping of which rows are dirty is left
            for r in ctx.orefs:
                g, slot = _table.grad_target(r.weight)
                if g is not None:
                    _table.note_dirty(slot, _snapshot(front, r.ids))
            return (None, None, None, None, d_lrw, d_lrb, None) + (None,) * len(ctx.orefs)
        # B3: the owner scatter-adds what every rank sent into its tables
        if ctx.orefs:
            targets = [_table.grad_target(r.weight) for r in ctx.orefs]
            ops.check(
                L.rh_fields_bwd(ops._field_array(ctx.orefs, [t[0] for t in targets]), len(ctx.orefs), dim, W * b, None, 0, ex.drows.data_ptr(), fmax * dim, None, None, None, None, None, None, err, st),
                "rh_fields_bwd")
            for r, (g, slot) in zip(ctx.orefs, targets):
                if g is not None:
                    _table.note_dirty(slot, _snapshot(front, r.ids))
        ex.drows.zero_()  # ready for the next step's REDs: they can only start after the next forward's two barriers
        return (None, None, None, None, d_lrw, d_lrb, None) + (None,) * len(ctx.orefs)


class ShardedFront(object):
    """Field-sharded replacement for ``EmbeddingLayer.forward`` (installed by :func:`attach`).

    CUDA route (default, ``config.p2p_exchange``): :class:`_ShardedP2P` — ids, rows and row gradients move through NVLink peer
    memory inside the engine's own kernels (see its docstring).
    NCCL route (``RECHUB_B200_P2P=0``; the checker of the peer-memory route), all launches static-shaped and graph-capturable:
      1 kernel    pack my samples' ids per owner                       (index_select / stack)
      NCCL        ids all-to-all
      1 launch    owner-side ``rh_fields_fwd`` over the global batch   -> rows (W*b, fmax*dim)
      NCCL        rows all-to-all (autograd Function; backward = the reverse exchange)
      1 launch    sample-side ``rh_fields_fwd`` reading the RECEIVED rows as its table: unpacks them into the model's
                  column order, appends the dense columns and (DeepFM) reduces FM + LR in the same pass
    CPU route (gloo tests): the same data flow with stock torch ops.
    """

    def __init__(self, layer, group, device):
        self.layer = layer
        self.group = group
        self.device = device
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.owner = field_owners(list(layer.embed_dict.keys()), self.world)
        self._plans = {}
        from . import config
        self.use_p2p = device.type == "cuda" and config.p2p_exchange
        self.lazy_clean = True  # a dense optimiser cleans the gradient rows lazily, in the next step (set per step by DistEngine)
        self.defer_barrier = False  # DistEngine.train_step issues the post-backward barrier itself, after launching the all-reduce
        self.deferred = None
        self.direct = None  # table name -> (peer-mapped gradient buffer pointer, vocab, padding_idx); set by DistEngine

    def owner_of(self, fea):
        return self.owner[fea.name if fea.shared_with is None else fea.shared_with]

    def _anchor(self):
        """A requires-grad scalar handed to the exchange Functions so that their backward (barriers, REDs, collectives) runs on
        every rank, also on one that owns no trainable table and passes no LR weights (fewer tables than ranks, frozen tables)."""
        a = getattr(self, "_anchor_t", None)
        if a is None:
            a = self._anchor_t = torch.zeros((), dtype=torch.float32, device=self.device, requires_grad=True)
        return a

    def _plan(self, features, batch):
        key = (tuple(id(f) for f in features), batch)
        p = self._plans.get(key)
        if p is None:
            sparse = [f for f in features if isinstance(f, (SparseFeature, SequenceFeature))]
            dense = [f for f in features if isinstance(f, DenseFeature)]
            for f in sparse:
                if isinstance(f, SequenceFeature):
                    raise NotImplementedError("field sharding covers SparseFeature columns; sequence features stay replicated (DIN tables are tiny: SURVEY §8e)")
            dims = {f.embed_dim for f in sparse}
            if len(dims) > 1:
                raise NotImplementedError("field sharding needs one embed_dim across the sharded features")
            W = self.world
            by_owner = [[f for f in sparse if self.owner_of(f) == r] for r in range(W)]
            fmax = max(len(b) for b in by_owner) if sparse else 0
            dim = sparse[0].embed_dim if sparse else 0
            # received-row index of (sample i, feature f): rows are laid out (owner, sample, slot)
            row_ids, cols = [], []
            ar = torch.arange(batch, dtype=torch.int32, device=self.device)
            for f in sparse:
                r = self.owner_of(f)
                k = by_owner[r].index(f)
                row_ids.append((r * batch + ar) * fmax + k)
                base = (r * fmax + k) * dim
                cols.extend(range(base, base + dim))
            p = {"sparse": sparse, "dense": dense, "by_owner": by_owner, "fmax": fmax, "dim": dim, "row_ids": row_ids, "batch": batch, "ex": None,
                 "cols": torch.tensor(cols, dtype=torch.long, device=self.device), "slots": [(r, k, f) for r in range(W) for k, f in enumerate(by_owner[r])]}
            self._plans[key] = p
        return p

    def _pack_ids(self, x, p, b):
        """(W, fmax, b) int64: chunk r = my samples' ids of the fields rank r owns (unused slots hold id 0)."""
        W, fmax = self.world, p["fmax"]
        ids = getattr(x, "ids", None)
        if ids is not None and ids.dim() == 2 and ids.dtype == torch.int64:  # PackedColumns: one gather of the transposed block
            sel = p.get("sel")
            if sel is None:
                pos = {n: j for j, n in enumerate(x.id_names)}
                sel_list = [0] * (W * fmax)
                for r, k, f in p["slots"]:
                    sel_list[r * fmax + k] = pos[f.name]
                sel = p["sel"] = torch.tensor(sel_list, dtype=torch.long, device=ids.device)
            return ids.t().index_select(0, sel).view(W, fmax, b)
        cols = [None] * (W * fmax)
        for r, k, f in p["slots"]:
            cols[r * fmax + k] = x[f.name].long()
        filler = next(c for c in cols if c is not None)
        return torch.stack([c if c is not None else filler for c in cols], dim=0).view(W, fmax, b)

    def run(self, x, features, fm_features=None, lr=None):
        """-> (tile (b, width) with sparse block first and dense appended, y_fm, y_lr)."""
        first = next(f for f in features if not isinstance(f, DenseFeature)) if any(not isinstance(f, DenseFeature) for f in features) else None
        if first is None:
            return self.layer._forward_local(x, features, squeeze_dim=True), None, None
        b = x[first.name].shape[0]
        p = self._plan(features, b)
        sparse, dense, W, fmax, dim = p["sparse"], p["dense"], self.world, p["fmax"], p["dim"]
        cuda = self.device.type == "cuda"
        if cuda and self.use_p2p and dim % 4 == 0 and W <= 8:
            if p["ex"] is None:
                p["ex"] = P2PExchange(self.group, self.device, W, b, fmax, dim, with_drows=self.direct is None)
            mine = p["by_owner"][self.rank]
            owned = [self.layer.table_of(f).weight for f in mine]
            return _ShardedP2P.apply(self, p, x, fm_features, lr[0] if lr else None, lr[1] if lr else None, self._anchor(), *owned)
        send = self._pack_ids(x, p, b)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)  # recv[s, k, i] = id of sample i of rank s for my k-th field
        mine = p["by_owner"][self.rank]
        # owner side: fused gather of MY fields over the global batch (W*b samples)
        ids_g = recv.permute(0, 2, 1).contiguous().view(W * b, fmax)  # [global sample, my k-th field]: one strided column per field
        xg = {f.name: ids_g[:, k] for k, f in enumerate(mine)}
        if cuda:
            from . import ops
            rows = torch.empty((W * b, fmax * dim), dtype=torch.float32, device=self.device)
            if mine:
                oplan = self.layer.build_plan(xg, mine, with_dense=False)
                oplan.out_tile = rows
                oplan.tile_width = fmax * dim  # the whole buffer travels; slots >= len(mine) are never read on the other side
                rows = ops.fused_tile(oplan)[0]
        else:
            rows = self.layer._forward_local(xg, mine, squeeze_dim=True) if mine else torch.zeros((W * b, 0))
            if len(mine) < fmax:
                rows = torch.nn.functional.pad(rows, (0, (fmax - len(mine)) * dim))
        got = _AllToAllRows.apply(rows.view(W, b, fmax * dim), self.group, self._anchor())  # got[r] = rank r's fields for MY samples
        if cuda:
            from . import ops
            table = got.view(W * b * fmax, dim)
            plan = ops.TilePlan(b, self.device)
            fm_slot = {f.name: j for j, f in enumerate(fm_features)} if fm_features else {}
            col = 0
            for f, rid in zip(sparse, p["row_ids"]):
                plan.fields.append(ops.FieldRef(table, rid, None, col, fm_slot.get(f.name, -1), is_act=True))
                col += dim
            for f in dense:
                v = x[f.name]
                if v.dtype not in ops._DENSE_CODES:
                    v = v.float()
                width = 1 if v.dim() == 1 else v.shape[1]
                plan.dense.append(ops.DenseRef(v, width, col))
                col += width
            plan.tile_width = col
            plan.act_table = table
            if fm_features:
                plan.n_fm, plan.fm_dim, plan.want_fm, plan.want_lr = len(fm_features), dim, True, lr is not None
            return ops.fused_tile(plan, lr[0] if lr else None, lr[1] if lr else None)
        grouped = got.permute(1, 0, 2).reshape(b, W * fmax * dim)
        parts = [grouped.index_select(1, p["cols"])]
        for f in dense:
            v = x[f.name].float()
            parts.append(v if v.dim() > 1 else v.unsqueeze(1))
        return (torch.cat(parts, dim=1) if len(parts) > 1 else parts[0]), None, None

    def forward(self, x, features, squeeze_dim):
        n_sparse = sum(1 for f in features if not isinstance(f, DenseFeature))
        if squeeze_dim:
            if not features:
                raise ValueError("The input features can note be empty")
            return self.run(x, features)[0]
        if n_sparse == 0:
            raise ValueError("If keep the original shape:[batch_size, num_features, embed_dim], expected %s in feature list, got %s" % ("SparseFeatures", features))
        only_sparse = [f for f in features if not isinstance(f, DenseFeature)]
        tile = self.run(x, only_sparse)[0]
        return tile.unflatten(1, (n_sparse, only_sparse[0].embed_dim))


class DistEngine(object):
    """Per-process state of a sharded run: the front end(s), the dense-gradient bucket, the step."""

    def __init__(self, model, device, group=None):
        self.model = model
        self.device = torch.device(device)
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        from ..basic.layers import EmbeddingLayer
        from .table import FieldTable
        self.fronts = []
        self.owned, self.foreign = [], []
        self.replicated_ids = set()  # table weights that stay whole on every rank (dense all-reduce of their gradients)
        for mod in model.modules():
            if isinstance(mod, EmbeddingLayer):
                if any(isinstance(f, SequenceFeature) for f in mod.features):
                    # a layer with sequence features (DIN: SURVEY §8e — its tables are a few MB) is not sharded: plain data
                    # parallelism, the tables' dense gradients ride the all-reduce bucket with the tower's
                    self.replicated_ids.update(id(t.weight) for t in mod.embed_dict.values())
                    continue
                front = ShardedFront(mod, self.group, self.device)
                mod._dist = front
                self.fronts.append(front)
                front._tables = [(name, front.owner[name], tuple(tbl.weight.shape), bool(tbl.weight.requires_grad), tbl.padding_idx) for name, tbl in mod.embed_dict.items()]
                for name, tbl in mod.embed_dict.items():
                    if front.owner[name] == self.rank:
                        self.owned.append(tbl.weight)
                    else:
                        self.foreign.append(tbl.weight)
                        tbl._dist_shape = tuple(tbl.weight.shape)
                        tbl.weight.data = torch.empty((0, tbl.weight.shape[1]), dtype=tbl.weight.dtype, device=tbl.weight.device)  # free it
                        tbl.weight.requires_grad_(False)
        self.grad_pool = None
        self._inv_world = torch.full((), 1.0 / self.world, dtype=torch.float32, device=self.device)
        self._map_gradient_buffers()
        skip = {id(p) for p in self.owned} | {id(p) for p in self.foreign}
        self.dense_params = [p for p in model.parameters() if id(p) not in skip and p.requires_grad]
        # replicated parameters start identical on every rank
        for p in self.dense_params:
            dist.broadcast(p.data, src=0, group=self.group)
        for b in model.buffers():
            if b.dtype.is_floating_point:
                dist.broadcast(b.data, src=0, group=self.group)
        self.peer_reduce = None
        self._setup_peer_reduce()
        n = sum(p.numel() for p in self.dense_params)
        self.n_dense = n
        self.flat = torch.zeros(n + 1, dtype=torch.float32, device=self.device)  # [+1]: the loss rides along
        self.views, off = [], 0
        for p in self.dense_params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def _setup_peer_reduce(self):
        """Peer-mapped staging buffers + flags of the engine's own all-reduce (rh_dense_pack_signal / rh_dense_reduce_update)."""
        from . import _lib, config
        if self.device.type != "cuda" or not config.p2p_allreduce or not config.p2p_exchange or self.world > 8 or not self.dense_params or len(self.dense_params) > 96:
            return
        if any(p.dtype != torch.float32 or not p.is_contiguous() for p in self.dense_params):
            return
        import ctypes
        import torch.distributed._symmetric_memory as symm
        n = len(self.dense_params)
        numel = (ctypes.c_int64 * n)(*[p.numel() for p in self.dense_params])
        total = int(_lib.lib().rh_dense_stage_floats(n, numel))
        stage = symm.empty(2 * self.world * total, dtype=torch.float32, device=self.device)  # 2 step parities x world slots: rank s pushes into slot s of everybody
        flags = symm.empty(8, dtype=torch.int32, device=self.device)
        stage.zero_()
        flags.zero_()
        h_stage, h_flags = symm.rendezvous(stage, self.group), symm.rendezvous(flags, self.group)
        self.peer_reduce = {
            "stage": stage, "flags": flags, "handles": (h_stage, h_flags), "numel": numel, "n": n,
            "stage_ptrs": (ctypes.c_void_p * self.world)(*[int(p) for p in h_stage.buffer_ptrs]),
            "flag_ptrs": (ctypes.c_void_p * self.world)(*[int(p) for p in h_flags.buffer_ptrs]),
            "epoch": torch.zeros(1, dtype=torch.int32, device=self.device), "ticket": torch.zeros(1, dtype=torch.int32, device=self.device),
            "extra_out": torch.zeros(4, dtype=torch.float32, device=self.device),
        }
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)  # nobody publishes into flags that are still being zeroed

    def _peer_reduce_step(self, opt, loss):
        """Pack + publish my gradients, run the owned tables' row-wise update while the peers publish theirs, then the fused
        wait + sum-in-rank-order + dense optimiser update.  Returns the global mean loss (device scalar)."""
        import ctypes
        from . import _lib
        L, pr, st = _lib.lib(), self.peer_reduce, _lib.stream_ptr()
        n = pr["n"]
        grads = (ctypes.c_void_p * n)(*[None if p.grad is None else (p.grad if p.grad.is_contiguous() else p.grad.contiguous()).data_ptr() for p in self.dense_params])
        keep = [p.grad for p in self.dense_params]  # alive until the launch is queued
        lw = (loss.detach() * self._inv_world).reshape(1)
        extra = (ctypes.c_void_p * 1)(lw.data_ptr())
        _lib.check(L.rh_dense_pack_signal(n, grads, pr["numel"], extra, 1, pr["stage_ptrs"], pr["flag_ptrs"], self.rank, self.world, pr["epoch"].data_ptr(), pr["ticket"].data_ptr(), st), "rh_dense_pack_signal")
        del keep
        from . import config
        if any(f.deferred is not None for f in self.fronts):
            # Every rank's row-gradient REDs must have landed before the owners consume their buffers.  A rank publishes its
            # gradients (above) behind its backward kernels and a system fence, so waiting for everybody's publication flag IS
            # that barrier — without a signal round of its own.
            if config.p2p_fold_barrier:
                _lib.check(L.rh_peer_wait(pr["flags"].data_ptr(), self.world, pr["epoch"].data_ptr(), st), "rh_peer_wait")
            else:
                next(f for f in self.fronts if f.deferred is not None).deferred.barrier()
        for f in self.fronts:
            f.deferred = None
        rw, eng = opt.rowwise, opt.dense_engine
        rw.set_lr(float(opt.dense.param_groups[0]["lr"]))
        rw.step()  # the owned tables' update does not need the peers' gradients: it overlaps their publication
        for p in self.dense_params:
            if id(p) not in eng.state:
                eng.state[id(p)] = (torch.zeros_like(p, memory_format=torch.contiguous_format) if rw.kind != 0 else None, torch.zeros_like(p, memory_format=torch.contiguous_format) if rw.kind == 1 else None)
        ptrs = lambda ts: (ctypes.c_void_p * n)(*[None if t is None else t.data_ptr() for t in ts])
        s1 = ptrs([eng.state[id(p)][0] for p in self.dense_params]) if rw.kind != 0 else None
        s2 = ptrs([eng.state[id(p)][1] for p in self.dense_params]) if rw.kind == 1 else None
        _lib.check(
            L.rh_dense_reduce_update(n, ptrs(self.dense_params), s1, s2, pr["numel"], 1, pr["extra_out"].data_ptr(), pr["stage"].data_ptr(), pr["flags"].data_ptr(), self.rank, self.world, pr["epoch"].data_ptr(),
                                     pr["ticket"].data_ptr(), rw.kind, rw._lr_dev.data_ptr(), rw._bc_dev.data_ptr(), rw.betas[0], rw.betas[1], rw.eps, rw.weight_decay, st), "rh_dense_reduce_update")
        return pr["extra_out"][0]

    def _map_gradient_buffers(self):
        """Direct gradients: carve every owned table's persistent gradient buffer out of ONE symmetric-memory pool per rank, so
        that the peers' backward kernels can RED row gradients straight into it.  Every rank derives the same layout from the
        model description (no communication besides the rendezvous)."""
        from . import config, table as _table
        fronts = [f for f in self.fronts if f.use_p2p]
        if self.device.type != "cuda" or not config.p2p_direct_grads or not fronts or self.world > 8:
            return
        layout, totals = [], [0] * self.world  # (front, name, owner, offset in floats, shape, trainable, padding_idx)
        for front in fronts:
            for name, owner, shape, trainable, pad in front._tables:
                if len(shape) != 2 or shape[1] % 4 != 0:
                    return  # the vector RED needs 16-byte rows; keep the staged route
                n = (shape[0] * shape[1] + 63) // 64 * 64
                layout.append((front, name, owner, totals[owner], shape, trainable, pad))
                totals[owner] += n
        import torch.distributed._symmetric_memory as symm
        self.grad_pool = symm.empty(max(max(totals), 64), dtype=torch.float32, device=self.device)
        self.grad_pool.zero_()
        handle = symm.rendezvous(self.grad_pool, self.group)
        ptrs = [int(p) for p in handle.buffer_ptrs]
        self._grad_pool_handle = handle
        for front in fronts:
            front.direct = {}
        for front, name, owner, off, shape, trainable, pad in layout:
            front.direct[name] = (ptrs[owner] + 4 * off if trainable else None, shape[0], -1 if pad is None else int(pad))
            if owner == self.rank and trainable:
                w = front.layer.embed_dict[name].weight
                slot = _table.slot_of(w)
                slot.buffer = self.grad_pool[off:off + shape[0] * shape[1]].view(shape)
                slot.pending, slot.all_dirty = [], False
                w.grad = None
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)  # nobody REDs into a pool that is still being zeroed

    # -- one training step -------------------------------------------------------------------------------------
    def train_step(self, trainer, x_dict, y):
        # zero_grad BEFORE the forward: with direct gradients the owned buffers are cleaned and re-attached ahead of the forward
        # barriers, i.e. before any peer's backward can RED into them
        for p in self.owned:
            p.grad = None
        for p in self.dense_params:
            p.grad = None  # autograd then hands over fresh gradient tensors (no accumulate kernels)
        opt = trainer.optimizer
        split = hasattr(opt, "rowwise") and hasattr(opt, "dense_engine")
        from . import config
        if split and self.device.type == "cuda":
            opt.rowwise.advance_early()  # joined by rowwise.step() below
        peer = split and self.peer_reduce is not None
        for f in self.fronts:
            # peer-memory reduction: the post-backward barrier is replaced by a wait on the gradient-publication flags (below)
            f.lazy_clean, f.defer_barrier, f.deferred = not split, bool(config.p2p_defer_barrier) or (peer and config.p2p_fold_barrier), None
        loss = trainer._loss(x_dict, y)
        loss.backward(self._inv_world)  # d(loss / world): the all-reduce SUM then yields the gradient of the global-batch mean
        for f in self.fronts:
            f.defer_barrier = False
        if split and self.peer_reduce is not None:
            return self._peer_reduce_step(opt, loss)
        # ONE bucket: [dense gradients ..., loss]; a parameter that got no gradient contributes zeros
        pieces = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self.dense_params]
        pieces.append((loss.detach() / self.world).reshape(1))
        torch.cat(pieces, out=self.flat)
        work = dist.all_reduce(self.flat, group=self.group, async_op=True)  # SUM of (local grad / world) = grad of the global-batch mean
        for f in self.fronts:  # every rank's row-gradient REDs must have landed before the owners consume their buffers
            if f.deferred is not None:
                f.deferred.barrier()
                f.deferred = None
                break  # one barrier orders all exchanges: every backward kernel of every rank precedes it
        for f in self.fronts:
            f.deferred = None
        if split:
            opt.rowwise.set_lr(float(opt.dense.param_groups[0]["lr"]))
            opt.rowwise.step()  # the owned tables' update does not need the all-reduce: it overlaps with it
        work.wait()
        for p, v in zip(self.dense_params, self.views):
            p.grad = v
        if split:
            opt.dense_engine.step()
        else:
            opt.step()
        return self.flat[-1]  # global mean loss

    def full_state_dict(self):
        """Reference-layout state_dict with every table gathered to all ranks' host memory (checkpoint compatibility)."""
        sd = {}
        for k, v in self.model.state_dict().items():
            sd[k] = v
        for front in self.fronts:
            for name, tbl in front.layer.embed_dict.items():
                owner = front.owner[name]
                shape = tuple(tbl.weight.shape) if owner == self.rank else tbl._dist_shape
                buf = tbl.weight.detach().clone() if owner == self.rank else torch.empty(shape, dtype=torch.float32, device=self.device)
                dist.broadcast(buf, src=owner, group=self.group)
                for k in list(sd.keys()):
                    if sd[k] is tbl.weight or (k.endswith("embed_dict.%s.weight" % name) and sd[k].shape[0] == 0):
                        sd[k] = buf.cpu()
        return sd


def attach(model, device, group=None):
    """Shard ``model``'s embedding tables over the ranks of the default process group; returns the engine."""
    if not dist.is_initialized():
        backend = "nccl" if torch.device(device).type == "cuda" else "gloo"
        dist.init_process_group(backend)
    return DistEngine(model, device, group)
