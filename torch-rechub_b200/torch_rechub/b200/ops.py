"""Autograd wrappers around the C-ABI kernels (CUDA tensors only).

Every function here launches hand-written sm_100a kernels from ``librechub_b200.so`` on the current
stream.  Table gradients never travel through autograd as tensors: the backward kernels scatter-add
into the table's persistent dense gradient buffer (``table.grad_target``), which is what
``weight.grad`` points at afterwards.
"""
import ctypes

import torch

from . import _lib, table as _table
from ._lib import RhDense, RhField, check, ptr, stream_ptr

_ID_DTYPES = (torch.int64, torch.int32)
_DENSE_CODES = {torch.float32: 0, torch.float64: 1, torch.int64: 2, torch.int32: 3}


def _as_ids(t):
    """ids as the kernels read them: int64/int32, no copy when already so (reference: ``.long()``, layers.py:83)."""
    return t if t.dtype in _ID_DTYPES else t.long()


def _rowmajor(t):
    """A 2-D view with unit inner stride (copy only when the layout forces it)."""
    if t.dim() != 2:
        t = t.reshape(t.shape[0], -1)
    if t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t


def _pad4(n):
    return (n + 3) // 4 * 4


# =====================================================================================================
# single-table lookup of any id shape  (FieldTable.forward)
# =====================================================================================================
class _RowsGather(torch.autograd.Function):

    @staticmethod
    def forward(ctx, weight, ids, padding_idx):
        L = _lib.lib()
        ids_c = _as_ids(ids).contiguous()
        n = ids_c.numel()
        vocab, dim = weight.shape
        out = torch.empty(tuple(ids.shape) + (dim,), dtype=torch.float32, device=weight.device)
        check(L.rh_rows_gather(weight.data_ptr(), vocab, dim, ids_c.data_ptr(), int(ids_c.dtype == torch.int32), n, out.data_ptr(), _lib.err_flag(weight.device).data_ptr(), stream_ptr()), "rh_rows_gather")
        ctx.weight = weight
        ctx.ids = ids_c
        ctx.padding_idx = -1 if padding_idx is None else int(padding_idx)
        return out

    @staticmethod
    def backward(ctx, d_out):
        weight, ids = ctx.weight, ctx.ids
        target, slot = _table.grad_target(weight)
        if target is not None:
            L = _lib.lib()
            vocab, dim = weight.shape
            g = d_out.contiguous()
            check(L.rh_rows_scatter_add(target.data_ptr(), vocab, dim, ctx.padding_idx, ids.data_ptr(), int(ids.dtype == torch.int32), ids.numel(), g.data_ptr(), _lib.err_flag(weight.device).data_ptr(), stream_ptr()),
                  "rh_rows_scatter_add")
            _table.note_dirty(slot, ids)
        return None, None, None


def table_lookup(tbl, ids):
    """``tbl(ids)`` for a CUDA :class:`FieldTable` (reference: nn.Embedding call at basic/layers.py:83-99)."""
    w = tbl.weight
    if w.dtype != torch.float32:
        raise NotImplementedError("the sm_100a engine keeps tables in fp32 (got %s)" % w.dtype)
    if not ids.is_cuda:
        raise RuntimeError("Expected all tensors to be on the same device: table on %s, ids on cpu" % w.device)
    return _RowsGather.apply(w, ids, tbl.padding_idx)


# =====================================================================================================
# the fused front end: multi-field gather (+ pooled sequences + dense columns) -> tile, (+ FM, + LR)
# =====================================================================================================
class FieldRef(object):
    """One SparseFeature column of a plan."""
    __slots__ = ("weight", "ids", "vocab", "dim", "padding_idx", "tile_col", "fm_slot", "is_act")

    def __init__(self, weight, ids, padding_idx, tile_col, fm_slot, is_act=False):
        self.is_act = is_act  # the "table" is an activation (rows received from peer GPUs): its gradient is returned, not stored
        self.weight = weight
        self.ids = ids
        self.vocab, self.dim = weight.shape
        self.padding_idx = -1 if padding_idx is None else int(padding_idx)
        self.tile_col = tile_col
        self.fm_slot = fm_slot


class SeqRef(object):
    """One pooled SequenceFeature column block of a plan (sum / mean pooling)."""
    __slots__ = ("weight", "ids", "vocab", "dim", "padding_idx", "mask_id", "mode", "tile_col")

    def __init__(self, weight, ids, padding_idx, mode, tile_col, mask_id="same"):
        self.weight = weight
        self.ids = ids
        self.vocab, self.dim = weight.shape
        # gradient skip = the OWNING table's nn.Embedding.padding_idx (aten::embedding_dense_backward); pooling mask = the
        # FEATURE's padding_idx (InputMask rule, layers.py:154-157).  They differ when a SequenceFeature shares a table
        # (shared_with) whose padding_idx is not its own.
        self.padding_idx = -1 if padding_idx is None else int(padding_idx)
        if isinstance(mask_id, str):
            mask_id = padding_idx
        self.mask_id = -1 if mask_id is None else int(mask_id)
        self.mode = mode
        self.tile_col = tile_col


class DenseRef(object):
    __slots__ = ("values", "width", "tile_col")

    def __init__(self, values, width, tile_col):
        self.values = values
        self.width = width
        self.tile_col = tile_col


class TilePlan(object):
    """What one fused front-end call has to produce."""

    def __init__(self, batch, device):
        self.batch = batch
        self.device = device
        self.fields = []  # FieldRef
        self.seqs = []  # SeqRef
        self.dense = []  # DenseRef
        self.tile_width = 0  # logical columns of the tile (0: no tile)
        self.n_fm = 0
        self.fm_dim = 0
        self.want_fm = False
        self.want_lr = False
        self.by_name = {}
        self.act_table = None  # (rows, dim) activation used as the table of every field flagged is_act
        self.out_tile = None  # optional preallocated (batch, ld) output buffer

    def weights(self):
        seen, out = set(), []
        for r in list(self.fields) + list(self.seqs):
            if getattr(r, "is_act", False):
                continue
            if id(r.weight) not in seen:
                seen.add(id(r.weight))
                out.append(r.weight)
        return out

    def field_groups(self):
        """Fields grouped by row width, in chunks of <= RH_MAX_FIELDS (one launch each).
        FM fields must all sit in ONE group (checked by the caller)."""
        by_dim = {}
        for f in sorted(self.fields, key=lambda r: r.fm_slot < 0):  # FM fields first (stable): they share one launch
            by_dim.setdefault(f.dim, []).append(f)
        groups = []
        for dim, fs in by_dim.items():
            for i in range(0, len(fs), _lib.RH_MAX_FIELDS):
                groups.append((dim, fs[i:i + _lib.RH_MAX_FIELDS]))
        return groups


def _field_array(refs, grads=None):
    arr = (RhField * len(refs))()
    for i, r in enumerate(refs):
        a = arr[i]
        ids = r.ids
        a.table = r.weight.data_ptr()
        a.table_grad = None if grads is None or grads[i] is None else grads[i].data_ptr()
        a.ids = ids.data_ptr()
        a.id_stride = ids.stride(0) if ids.dim() > 0 and ids.shape[0] > 1 else 1
        a.ids_are_i32 = int(ids.dtype == torch.int32)
        a.vocab = r.vocab
        a.padding_idx = r.padding_idx
        a.tile_col = r.tile_col
        a.fm_slot = r.fm_slot
    return arr


def _dense_array(refs):
    arr = (RhDense * max(len(refs), 1))()
    for i, r in enumerate(refs):
        v = r.values
        a = arr[i]
        a.values = v.data_ptr()
        a.stride = v.stride(0) if v.shape[0] > 1 else r.width
        a.dtype = _DENSE_CODES[v.dtype]
        a.width = r.width
        a.tile_col = r.tile_col
    return arr


class _FusedTile(torch.autograd.Function):
    """tile, y_fm, y_lr = front(plan).  Inputs carried for autograd connectivity only: lr weight/bias, table weights."""

    @staticmethod
    def forward(ctx, plan, lr_weight, lr_bias, act_table, *weights):
        L = _lib.lib()
        dev = plan.device
        B = plan.batch
        st = stream_ptr()
        err = _lib.err_flag(dev).data_ptr()
        tile = None
        ld = 0
        if plan.tile_width > 0:
            if plan.out_tile is not None:
                tile = plan.out_tile
                ld = tile.stride(0)
            else:
                ld = _pad4(plan.tile_width)
                tile = torch.empty((B, ld), dtype=torch.float32, device=dev)  # columns >= tile_width are padding, never read
        y_fm = torch.empty(B, dtype=torch.float32, device=dev) if plan.want_fm else None
        y_lr = torch.empty(B, dtype=torch.float32, device=dev) if plan.want_lr else None
        fsum = torch.empty((B, plan.fm_dim), dtype=torch.float32, device=dev) if plan.want_fm else None

        groups = plan.field_groups()
        dense_left = list(plan.dense)
        if not groups and dense_left:
            groups = [(4, [])]
        for dim, refs in groups:
            has_fm = any(r.fm_slot >= 0 for r in refs)
            dn = dense_left[:_lib.RH_MAX_DENSE]
            dense_left = dense_left[len(dn):]
            check(
                L.rh_fields_fwd(_field_array(refs) if refs else None, len(refs), dim,
                                _dense_array(dn) if dn else None, len(dn), B, ptr(tile), ld, ptr(lr_weight) if (has_fm and plan.want_lr) else None,
                                ptr(lr_bias) if (has_fm and plan.want_lr) else None,
                                ptr(y_fm) if has_fm else None, ptr(y_lr) if (has_fm and plan.want_lr) else None, ptr(fsum) if has_fm else None, err, st), "rh_fields_fwd")
        while dense_left:  # more than RH_MAX_DENSE numeric columns
            dn = dense_left[:_lib.RH_MAX_DENSE]
            dense_left = dense_left[len(dn):]
            check(L.rh_fields_fwd(None, 0, 4, _dense_array(dn), len(dn), B, ptr(tile), ld, None, None, None, None, None, err, st), "rh_fields_fwd")
        for s in plan.seqs:
            ids = s.ids
            check(
                L.rh_seq_pool_fwd(s.weight.data_ptr(), s.vocab, s.dim, ids.data_ptr(), int(ids.dtype == torch.int32), B, ids.shape[1], s.mode, s.mask_id, tile.data_ptr() + 4 * s.tile_col, ld, err, st),
                "rh_seq_pool_fwd")

        ctx.plan = plan
        ctx.ld = ld
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(tile, fsum, lr_weight)
        tile_view = None
        if tile is not None:
            tile_view = tile if tile.shape[1] == plan.tile_width else tile[:, :plan.tile_width]
        return tile_view, y_fm, y_lr

    @staticmethod
    def backward(ctx, d_tile, d_yfm, d_ylr):
        L = _lib.lib()
        plan = ctx.plan
        tile, fsum, lr_weight = ctx.saved_tensors
        B = plan.batch
        st = stream_ptr()
        err = _lib.err_flag(plan.device).data_ptr()
        d_ld = 0
        if d_tile is not None:
            d_tile = _rowmajor(d_tile)
            d_ld = d_tile.stride(0) if B > 1 else d_tile.shape[1]
        if d_yfm is not None:
            d_yfm = d_yfm.contiguous()
        if d_ylr is not None:
            d_ylr = d_ylr.contiguous()
        d_lrw = d_lrb = None
        if d_ylr is not None and lr_weight is not None:
            n = lr_weight.numel()
            buf = torch.zeros(_pad4(n) + 1, dtype=torch.float32, device=plan.device)
            d_lrw = buf[:n].view_as(lr_weight)
            d_lrb = buf[_pad4(n):_pad4(n) + 1]

        d_act = None
        if plan.act_table is not None:
            d_act = torch.zeros_like(plan.act_table)  # every received row is written once by the scatter below
        for dim, refs in plan.field_groups():
            targets = [(d_act, None) if r.is_act else _table.grad_target(r.weight) for r in refs]
            grads = [t[0] for t in targets]
            has_fm = any(r.fm_slot >= 0 for r in refs) and (d_yfm is not None or d_ylr is not None)
            if d_tile is None and not has_fm:
                continue
            check(
                L.rh_fields_bwd(_field_array(refs, grads), len(refs), dim, B, ptr(tile), ctx.ld, ptr(d_tile), d_ld, ptr(d_yfm) if has_fm else None,
                                ptr(d_ylr) if has_fm else None, ptr(lr_weight) if has_fm else None, ptr(fsum) if has_fm else None, ptr(d_lrw) if has_fm else None,
                                ptr(d_lrb) if has_fm else None, err, st), "rh_fields_bwd")
            for r, (g, slot) in zip(refs, targets):
                if g is not None and not r.is_act:
                    _table.note_dirty(slot, r.ids)
        if d_tile is not None:
            for s in plan.seqs:
                g, slot = _table.grad_target(s.weight)
                if g is None:
                    continue
                ids = s.ids
                check(
                    L.rh_seq_pool_bwd(g.data_ptr(), s.vocab, s.dim, s.padding_idx, ids.data_ptr(), int(ids.dtype == torch.int32), B, ids.shape[1], s.mode, s.mask_id, d_tile.data_ptr() + 4 * s.tile_col, d_ld, err, st),
                    "rh_seq_pool_bwd")
                _table.note_dirty(slot, ids)
        return (None, d_lrw, d_lrb, d_act) + (None,) * len(plan.weights())


def fused_tile(plan, lr_weight=None, lr_bias=None):
    """Run a :class:`TilePlan`; returns ``(tile (B, width) or None, y_fm (B,) or None, y_lr (B,) or None)``."""
    ws = plan.weights()
    for w in ws:
        if w.dtype != torch.float32:
            raise NotImplementedError("the sm_100a engine keeps tables in fp32 (got %s)" % w.dtype)
    return _FusedTile.apply(plan, lr_weight, lr_bias, plan.act_table, *ws)


# =====================================================================================================
# FM on a materialised (B, F, D) tensor
# =====================================================================================================
class _FM(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, reduce_sum):
        L = _lib.lib()
        xc = x.contiguous()
        B, F, D = xc.shape
        y = torch.empty((B, 1) if reduce_sum else (B, D), dtype=torch.float32, device=x.device)
        check(L.rh_fm_fwd(xc.data_ptr(), B, F, D, int(reduce_sum), y.data_ptr(), stream_ptr()), "rh_fm_fwd")
        ctx.save_for_backward(xc)
        ctx.reduce_sum = reduce_sum
        return y

    @staticmethod
    def backward(ctx, d_y):
        L = _lib.lib()
        (xc,) = ctx.saved_tensors
        B, F, D = xc.shape
        d_x = torch.empty_like(xc)
        g = d_y.contiguous()
        check(L.rh_fm_bwd(xc.data_ptr(), g.data_ptr(), B, F, D, int(ctx.reduce_sum), d_x.data_ptr(), stream_ptr()), "rh_fm_bwd")
        return d_x, None


def fm(x, reduce_sum=True):
    if x.dtype != torch.float32 or x.dim() != 3:
        raise NotImplementedError("FM kernel takes a (B, F, D) fp32 tensor")
    return _FM.apply(x, bool(reduce_sum))


# =====================================================================================================
# CrossNetwork
# =====================================================================================================
def _ptr_array(tensors):
    arr = (ctypes.c_void_p * max(len(tensors), 1))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


class _Cross(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, n_layers, *params):
        L = _lib.lib()
        ws, bs = params[:n_layers], params[n_layers:]
        x2 = _rowmajor(x)
        B, W = x2.shape
        x_ld = x2.stride(0) if B > 1 else W
        out_ld = _pad4(W)
        out = torch.empty((B, out_ld), dtype=torch.float32, device=x.device)
        xw = torch.empty((max(n_layers, 1), B), dtype=torch.float32, device=x.device)
        wc = [w.contiguous() for w in ws]
        bc = [b.contiguous() for b in bs]
        check(L.rh_cross_fwd(x2.data_ptr(), x_ld, B, W, n_layers, _ptr_array(wc), _ptr_array(bc), out.data_ptr(), out_ld, xw.data_ptr(), stream_ptr()), "rh_cross_fwd")
        ctx.n_layers = n_layers
        ctx.save_for_backward(x2, xw, *wc, *bc)
        return out if out_ld == W else out[:, :W]

    @staticmethod
    def backward(ctx, d_out):
        L = _lib.lib()
        n = ctx.n_layers
        saved = ctx.saved_tensors
        x2, xw = saved[0], saved[1]
        wc, bc = saved[2:2 + n], saved[2 + n:2 + 2 * n]
        B, W = x2.shape
        x_ld = x2.stride(0) if B > 1 else W
        g = _rowmajor(d_out)
        g_ld = g.stride(0) if B > 1 else W
        dx_ld = _pad4(W)
        d_x = torch.empty((B, dx_ld), dtype=torch.float32, device=x2.device)
        Wp = _pad4(W)
        gbuf = torch.zeros((2 * max(n, 1), Wp), dtype=torch.float32, device=x2.device)
        d_ws = [gbuf[l, :W] for l in range(n)]
        d_bs = [gbuf[n + l, :W] for l in range(n)]
        check(
            L.rh_cross_bwd(x2.data_ptr(), x_ld, B, W, n, _ptr_array(wc), _ptr_array(bc), xw.data_ptr(), g.data_ptr(), g_ld, d_x.data_ptr(), dx_ld, _ptr_array(d_ws), _ptr_array(d_bs), stream_ptr()), "rh_cross_bwd")
        d_x = d_x if dx_ld == W else d_x[:, :W]
        return (d_x, None) + tuple(dw.view_as(w) for dw, w in zip(d_ws, wc)) + tuple(d_bs)


def cross_network(x, weights, biases):
    """CrossNetwork.forward on CUDA; ``weights[l]``: (1, W) Linear weight, ``biases[l]``: (W,) (layers.py:409-420)."""
    n = len(weights)
    if x.dtype != torch.float32:
        x = x.float()
    return _Cross.apply(x, n, *weights, *biases)


# =====================================================================================================
# DCN-v2 cross networks: CrossNetMix (low-rank mixture of experts) and CrossNetV2 (full rank)
# =====================================================================================================
def _buf(rows, cols, dev, zero=False):
    """(rows, cols) fp32 view of a buffer whose row stride is a multiple of 4 floats (TMA / 16-byte rows)."""
    ld = _pad4(cols)
    t = (torch.zeros if zero else torch.empty)((rows, ld), dtype=torch.float32, device=dev)
    return t if ld == cols else t[:, :cols]


def _padded_rows(x):
    """x as a (rows, cols) fp32 view with unit inner stride, 16-byte aligned rows (copy only when the layout forces it)."""
    x2 = _rowmajor(x if x.dtype == torch.float32 else x.float())
    if x2.stride(0) % 4 != 0 or x2.data_ptr() % 16 != 0:
        b = _buf(x2.shape[0], x2.shape[1], x2.device)
        b.copy_(x2)
        x2 = b
    return x2


def _mm(A, a_mn, Bm, b_mn, M, N, K, tc, split_k=1):
    """C[M, N] = op(A) op(B)^T: rh_gemm_tf32x3 on the tensor cores when ``tc``, else the library GEMM on the same views (thin
    shapes: fewer than 128 rows or an operand narrower than one 32-float TMA box)."""
    out = _buf(M, N, A.device, zero=bool(tc and split_k > 1))
    if tc:
        gemm3x(A, a_mn, Bm, b_mn, M, N, K, split_k=split_k, out=out)
    else:
        Al = A.t() if a_mn else A
        Bl = Bm.t() if b_mn else Bm
        out.copy_(Al[:M, :K] @ Bl[:N, :K].t())
    return out


def _ptrs_of(tensors):
    arr = (ctypes.c_void_p * max(len(tensors), 1))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


class _CrossMix(torch.autograd.Function):
    """All layers of CrossNetMix (basic/layers.py:470-506) on the packed operands of rh_crossmix_pack: per layer three GEMMs
    (rh_gemm_tf32x3) + rh_crossmix_mid1_fwd / _mid2_fwd / _out_fwd; backward: six GEMMs + three maps per layer, then
    rh_crossmix_unpack_grads.  ``params``: u_list (L), v_list (L), c_list (L), gating weights (E), bias (L)."""

    @staticmethod
    def forward(ctx, x, L_, E, *params):
        lib = _lib.lib()
        st = stream_ptr()
        us, vs, cs = params[:L_], params[L_:2 * L_], params[2 * L_:3 * L_]
        gates, biases = params[3 * L_:3 * L_ + E], params[3 * L_ + E:]
        x0 = _padded_rows(x)
        B, W = x0.shape
        dev = x0.device
        r = vs[0].shape[2]
        Er, N1 = E * r, E * r + E
        ldw = _pad4(W)
        wcat = [torch.empty((N1, ldw), dtype=torch.float32, device=dev) for _ in range(L_)]
        cbd = [torch.empty((Er, Er), dtype=torch.float32, device=dev) for _ in range(L_)]
        ucat = [torch.empty((W, Er), dtype=torch.float32, device=dev) for _ in range(L_)]
        uc, vc, cc = [t.contiguous() for t in us], [t.contiguous() for t in vs], [t.contiguous() for t in cs]
        gc = [g.contiguous() for g in gates]
        check(lib.rh_crossmix_pack(L_, E, W, r, _ptrs_of(uc), _ptrs_of(vc), _ptrs_of(cc), _ptrs_of(gc), _ptrs_of(wcat), ldw, _ptrs_of(cbd), _ptrs_of(ucat), st), "rh_crossmix_pack")
        tc = B >= 128 and min(W, Er) >= 32 and Er % 4 == 0 and N1 % 4 == 0 and _tc_ok(B, x0)
        saved = []
        xl = x0
        for l in range(L_):
            ag = _mm(xl, False, wcat[l][:, :W], False, B, N1, W, tc)
            t1 = torch.empty((B, Er), dtype=torch.float32, device=dev)
            s_ = torch.empty((B, E), dtype=torch.float32, device=dev)
            check(lib.rh_crossmix_mid1_fwd(ag.data_ptr(), ag.stride(0), B, E, r, t1.data_ptr(), s_.data_ptr(), st), "rh_crossmix_mid1_fwd")
            P = _mm(t1, False, cbd[l], False, B, Er, Er, tc)
            if P.stride(0) != Er:
                P = P.contiguous()
            t2 = torch.empty((B, Er), dtype=torch.float32, device=dev)
            z = torch.empty((B, Er), dtype=torch.float32, device=dev)
            check(lib.rh_crossmix_mid2_fwd(P.data_ptr(), s_.data_ptr(), B, E, r, t2.data_ptr(), z.data_ptr(), st), "rh_crossmix_mid2_fwd")
            u = _mm(z, False, ucat[l], False, B, W, Er, tc)
            bl = biases[l].contiguous()
            nxt = _buf(B, W, dev)
            check(lib.rh_crossmix_out_fwd(x0.data_ptr(), x0.stride(0), xl.data_ptr(), xl.stride(0), u.data_ptr(), u.stride(0), bl.data_ptr(), 0, B, W, nxt.data_ptr(), nxt.stride(0), st), "rh_crossmix_out_fwd")
            saved.append((xl, t1, s_, t2, z, u, bl))
            xl = nxt
        ctx.saved_layers = saved
        ctx.packed = (wcat, cbd, ucat)
        ctx.meta = (L_, E, r, W, B, tc)
        ctx.shapes = ([t.shape for t in us], [t.shape for t in cs], [g.shape for g in gates], [b.shape for b in biases])
        ctx.x0 = x0
        return xl

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.lib()
        st = stream_ptr()
        L_, E, r, W, B, tc = ctx.meta
        wcat, cbd, ucat = ctx.packed
        x0 = ctx.x0
        dev = x0.device
        Er, N1 = E * r, E * r + E
        g1 = _padded_rows(d_out)
        g2 = None
        d_x0 = _buf(B, W, dev, zero=True)
        d_bias = torch.zeros((L_, _pad4(W)), dtype=torch.float32, device=dev)
        d_wcat, d_cbd, d_ucat = [None] * L_, [None] * L_, [None] * L_
        split = _split_k_for(2, 4, (B + 31) // 32, budget=128) if tc else 1
        for l in reversed(range(L_)):
            xl, t1, s_, t2, z, u, bl = ctx.saved_layers[l]
            g_sum = _buf(B, W, dev)
            d_u = _buf(B, W, dev)
            check(
                lib.rh_crossmix_out_bwd(g1.data_ptr(), g1.stride(0), ptr(g2), g2.stride(0) if g2 is not None else 0, x0.data_ptr(), x0.stride(0), u.data_ptr(), u.stride(0), bl.data_ptr(), 0, B, W, g_sum.data_ptr(),
                                        g_sum.stride(0), d_u.data_ptr(), d_u.stride(0), d_x0.data_ptr(), d_x0.stride(0), d_bias[l].data_ptr(), st), "rh_crossmix_out_bwd")
            d_z = _mm(d_u, False, ucat[l], True, B, Er, W, tc)  # d_z = d_u Ucat
            d_ucat[l] = _mm(d_u, True, z, True, W, Er, B, tc, split_k=split)  # d_Ucat = d_u^T z
            d_P = torch.empty((B, Er), dtype=torch.float32, device=dev)
            d_ag = _buf(B, N1, dev)
            check(lib.rh_crossmix_mid2_bwd(d_z.data_ptr(), d_z.stride(0), s_.data_ptr(), t2.data_ptr(), B, E, r, d_P.data_ptr(), d_ag.data_ptr(), d_ag.stride(0), st), "rh_crossmix_mid2_bwd")
            d_t1 = _mm(d_P, False, cbd[l], True, B, Er, Er, tc)  # d_t1 = d_P Cbd
            d_cbd[l] = _mm(d_P, True, t1, True, Er, Er, B, tc, split_k=split)  # d_Cbd = d_P^T t1
            check(lib.rh_crossmix_mid1_bwd(d_t1.data_ptr(), d_t1.stride(0), t1.data_ptr(), B, E, r, d_ag.data_ptr(), d_ag.stride(0), st), "rh_crossmix_mid1_bwd")
            g2 = _mm(d_ag, False, wcat[l][:, :W], True, B, W, N1, tc)  # (d x_l through the projections) = d_ag Wcat
            d_wcat[l] = _mm(d_ag, True, xl, True, N1, W, B, tc, split_k=split)  # d_Wcat = d_ag^T x_l
            g1 = g_sum
        d_x = _buf(B, W, dev)
        check(lib.rh_sum3(g1.data_ptr(), g1.stride(0), g2.data_ptr(), g2.stride(0), d_x0.data_ptr(), d_x0.stride(0), B, W, d_x.data_ptr(), d_x.stride(0), st), "rh_sum3")
        u_shapes, c_shapes, g_shapes, b_shapes = ctx.shapes
        d_us = [torch.empty(tuple(sh), dtype=torch.float32, device=dev) for sh in u_shapes]
        d_vs = [torch.empty(tuple(sh), dtype=torch.float32, device=dev) for sh in u_shapes]
        d_cs = [torch.empty(tuple(sh), dtype=torch.float32, device=dev) for sh in c_shapes]
        d_gs = [torch.empty(W, dtype=torch.float32, device=dev) for _ in g_shapes]
        check(
            lib.rh_crossmix_unpack_grads(L_, E, W, r, _ptrs_of(d_wcat), d_wcat[0].stride(0), _ptrs_of(d_cbd), d_cbd[0].stride(0), _ptrs_of(d_ucat), d_ucat[0].stride(0), _ptrs_of(d_us), _ptrs_of(d_vs), _ptrs_of(d_cs),
                                         _ptrs_of(d_gs), st), "rh_crossmix_unpack_grads")
        d_bs = [d_bias[l, :W].reshape(tuple(sh)) for l, sh in enumerate(b_shapes)]
        return (d_x, None, None) + tuple(d_us) + tuple(d_vs) + tuple(d_cs) + tuple(g.view(tuple(sh)) for g, sh in zip(d_gs, g_shapes)) + tuple(d_bs)


def cross_net_mix(x, u_list, v_list, c_list, gate_weights, biases):
    """CrossNetMix.forward on CUDA (returns (B, width); the caller applies the reference's ``squeeze()``)."""
    L_, E = len(u_list), len(gate_weights)
    if L_ > 8 or E > 8:
        raise NotImplementedError("CrossNetMix kernels cover <= 8 layers and <= 8 experts (got %d, %d)" % (L_, E))
    return _CrossMix.apply(x, L_, E, *u_list, *v_list, *c_list, *gate_weights, *biases)


class _CrossV2(torch.autograd.Function):
    """CrossNetV2 (basic/layers.py:440-444): per layer one width x width GEMM (rh_gemm_tf32x3) + rh_crossmix_out_fwd."""

    @staticmethod
    def forward(ctx, x, L_, *params):
        lib = _lib.lib()
        st = stream_ptr()
        ws, bs = params[:L_], params[L_:]
        x0 = _padded_rows(x)
        B, W = x0.shape
        dev = x0.device
        tc = B >= 128 and W >= 32 and _tc_ok(B, x0)
        wp = [_padded_weight(w) if tc else w for w in ws]
        saved = []
        xl = x0
        for l in range(L_):
            u = _mm(xl, False, wp[l], False, B, W, W, tc)
            bl = bs[l].contiguous()
            nxt = _buf(B, W, dev)
            check(lib.rh_crossmix_out_fwd(x0.data_ptr(), x0.stride(0), xl.data_ptr(), xl.stride(0), u.data_ptr(), u.stride(0), bl.data_ptr(), 1, B, W, nxt.data_ptr(), nxt.stride(0), st), "rh_crossmix_out_fwd")
            saved.append((xl, u, bl))
            xl = nxt
        ctx.saved_layers, ctx.wp, ctx.meta, ctx.x0 = saved, wp, (L_, W, B, tc), x0
        return xl

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.lib()
        st = stream_ptr()
        L_, W, B, tc = ctx.meta
        x0 = ctx.x0
        dev = x0.device
        g1, g2 = _padded_rows(d_out), None
        d_x0 = _buf(B, W, dev, zero=True)
        d_bias = torch.zeros((L_, _pad4(W)), dtype=torch.float32, device=dev)
        d_ws = [None] * L_
        split = _split_k_for((W + 127) // 128, (W + 127) // 128, (B + 31) // 32, budget=128) if tc else 1
        for l in reversed(range(L_)):
            xl, u, bl = ctx.saved_layers[l]
            g_sum, d_u = _buf(B, W, dev), _buf(B, W, dev)
            check(
                lib.rh_crossmix_out_bwd(g1.data_ptr(), g1.stride(0), ptr(g2), g2.stride(0) if g2 is not None else 0, x0.data_ptr(), x0.stride(0), u.data_ptr(), u.stride(0), bl.data_ptr(), 1, B, W, g_sum.data_ptr(),
                                        g_sum.stride(0), d_u.data_ptr(), d_u.stride(0), d_x0.data_ptr(), d_x0.stride(0), d_bias[l].data_ptr(), st), "rh_crossmix_out_bwd")
            g2 = _mm(d_u, False, ctx.wp[l], True, B, W, W, tc)  # d_u W
            d_ws[l] = _mm(d_u, True, xl, True, W, W, B, tc, split_k=split)  # d_u^T x_l
            g1 = g_sum
        d_x = _buf(B, W, dev)
        check(lib.rh_sum3(g1.data_ptr(), g1.stride(0), g2.data_ptr(), g2.stride(0), d_x0.data_ptr(), d_x0.stride(0), B, W, d_x.data_ptr(), d_x.stride(0), st), "rh_sum3")
        return (d_x, None) + tuple(d_ws) + tuple(d_bias[l, :W] for l in range(L_))


def cross_net_v2(x, weights, biases):
    return _CrossV2.apply(x, len(weights), *weights, *biases)


# =====================================================================================================
# BatchNorm1d + activation + dropout
# =====================================================================================================
ACT_CODES = {"none": 0, "relu": 1, "dice": 2, "prelu": 3, "sigmoid": 4, "leakyrelu": 5}
_scratch = {}


def _colstats_scratch(device, cols):
    key = (device.index, cols)
    t = _scratch.get(key)
    if t is None:
        t = torch.zeros(2 * cols + 1, dtype=torch.float32, device=device)
        _scratch[key] = t
    return t


def gemm3x(A, a_mn, B, b_mn, M, N, K, bias=None, split_k=1, out=None, out_cols=None):
    """C[M, N] = op(A) @ op(B)^T (+ bias) on the tensor cores, fp32-accurate (rh_gemm_tf32x3).
    ``A``: (M, K) row-major when ``a_mn`` is False, else the stored (K, M) matrix whose transpose is the operand; same for B.
    ``out``: optional (M, >= N) buffer with unit inner stride; with ``split_k > 1`` it must be zero."""
    L = _lib.lib()
    if out is None:
        ld = _pad4(N) if out_cols is None else out_cols
        out = torch.zeros((M, ld), dtype=torch.float32, device=A.device) if split_k > 1 else torch.empty((M, ld), dtype=torch.float32, device=A.device)
    check(
        L.rh_gemm_tf32x3(A.data_ptr(), A.stride(0), int(a_mn), B.data_ptr(), B.stride(0), int(b_mn), out.data_ptr(), out.stride(0), M, N, K, ptr(bias), int(split_k), stream_ptr()),
        "rh_gemm_tf32x3")
    return out if out.shape[1] == N else out[:, :N]


def _tc_ok(rows, *mats):
    """Shapes/strides the TMA-fed kernel accepts; tiny problems stay on the library GEMM."""
    from . import config
    if not config.tensor_core_gemm or rows < 128:
        return False
    if any(m.dim() == 2 and min(m.shape) < 32 for m in mats):  # thinner than one 32-float TMA box: not worth a 128 x 128 tile
        return False
    return all(m.dtype == torch.float32 and m.dim() == 2 and m.stride(1) == 1 and m.stride(0) % 4 == 0 and m.data_ptr() % 16 == 0 for m in mats)


_padded_weights = {}


def _padded_weight(W):
    """W (N, K) with a row stride that is a multiple of 4 floats (TMA): W itself, or a refreshed padded copy."""
    if W.stride(1) == 1 and W.stride(0) % 4 == 0 and W.data_ptr() % 16 == 0:
        return W
    key = id(W)
    ent = _padded_weights.get(key)
    N, K = W.shape
    if ent is None or ent[0]() is not W or ent[1].device != W.device:
        import weakref
        buf = torch.zeros((N, _pad4(K)), dtype=torch.float32, device=W.device)
        _padded_weights[key] = (weakref.ref(W, lambda _r, _k=key: _padded_weights.pop(_k, None)), buf)
        ent = _padded_weights[key]
    ent[1][:, :K].copy_(W.detach())
    return ent[1][:, :K]


def _split_k_for(m_tiles, n_tiles, k_blocks, budget=128):
    """Enough K slices to put ~``budget`` CTAs on the 148 SMs when the output has few tiles (weight gradients)."""
    tiles = max(1, m_tiles * n_tiles)
    return max(1, min(k_blocks, budget // tiles))


_aux_streams = {}


def _aux_stream(dev):
    """Second stream of a device for the independent half of a backward step (forked and joined inside one autograd node)."""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    st = _aux_streams.get(key)
    if st is None:
        st = _aux_streams[key] = torch.cuda.Stream(device=dev)
    return st


_salt_counter = [0]


def _dropout_seed(module):
    """Per-layer dropout stream seed: deterministic under torch.manual_seed and the order layers are first used in."""
    salt = getattr(module, "_rh_salt", None)
    if salt is None:
        _salt_counter[0] += 1
        salt = module._rh_salt = _salt_counter[0]
    return (torch.initial_seed() * 0x9E3779B1 + salt * 0x85EBCA6B) & 0xFFFFFFFF


_fused_bn_scratch = {}


def _bn_fused_scratch(dev, cols):
    """Scratch of the fused BatchNorm kernels per (device, width): column sums + barrier counters, zeroed once, left zeroed by
    every launch.  Launches that share it are ordered on one stream (a tower's layers run one after the other)."""
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), cols)
    buf = _fused_bn_scratch.get(key)
    if buf is None:
        buf = _fused_bn_scratch[key] = torch.zeros(int(_lib.lib().rh_bn_fused_scratch_floats(cols)), dtype=torch.float32, device=dev)
    return buf


def _linear_fwd(x2, W, b):
    """h = x2 @ W^T + b -> (h, padded weight view or None, ran on the tensor cores?)."""
    rows, K = x2.shape
    cols = W.shape[0]
    h = torch.empty((rows, cols), dtype=torch.float32, device=x2.device)
    use_tc = cols % 4 == 0 and cols >= 32 and _tc_ok(rows, x2)
    Wp = None
    if use_tc:
        Wp = _padded_weight(W)  # (cols, K) view with a 16-byte row stride
        gemm3x(x2, False, Wp, False, rows, cols, K, bias=b, out=h)
    elif b is not None:
        torch.addmm(b, x2, W.t(), out=h)
    else:
        torch.mm(x2, W.t(), out=h)
    return h, Wp, use_tc


def _dw_split(rows, cols, K, need_dx):
    """(split_k of the weight-gradient GEMM, runs next to the input-gradient GEMM?) — shared by _prezero_dw and _linear_bwd."""
    from . import config
    concurrent = bool(config.concurrent_tower_bwd and need_dx)
    return _split_k_for((cols + 127) // 128, (K + 127) // 128, (rows + 31) // 32, budget=64 if concurrent else 128), concurrent


def _prezero_dw(ctx, x2, W, use_tc, need_dx):
    """Forward-time half of the weight gradient: its split-K target has to be zero, and that fill was a node between the BatchNorm
    backward and the gradient GEMMs (on the main stream) or in front of the dW GEMM on the side stream (dW ~ dX in duration, so it
    lengthened the join: measured +4 us per step).  The FORWARD has the side stream idle: allocate and zero the buffer there, next
    to this layer's forward GEMM; backward finds it ready (``ctx.dw_pre``, consumed once)."""
    from . import config
    ctx.dw_pre = None
    if not (use_tc and config.prezero_dw and x2.is_cuda and len(ctx.needs_input_grad) > 1 and ctx.needs_input_grad[1]):
        return
    rows, K = x2.shape
    cols = W.shape[0]
    split, _ = _dw_split(rows, cols, K, need_dx)
    if split <= 1:
        return
    dev = x2.device
    d_W = torch.empty((cols, K), dtype=torch.float32, device=dev)
    cur, aux = torch.cuda.current_stream(), _aux_stream(dev)
    aux.wait_stream(cur)  # the block may have been freed by work queued on the main stream
    with torch.cuda.stream(aux):
        d_W.zero_()
    ctx.dw_pre = d_W  # first touched again on the side stream (the dW GEMM) or after a join with it


def _linear_bwd(d_h, x2, W, Wp, use_tc, need_dx, pre=None):
    """(d_x or None, d_W) of h = x2 @ W^T from d_h; on the tensor cores the weight-gradient GEMM runs on a second stream next to
    the input-gradient GEMM (config.concurrent_tower_bwd).  ``pre``: the zeroed split-K target from _prezero_dw."""
    rows, K = x2.shape
    cols = W.shape[0]
    dev = d_h.device
    fork = None
    if use_tc:  # dW[cols, K] = d_h^T x: both operands read as stored (MN-major), K = rows split over CTAs
        split, concurrent = _dw_split(rows, cols, K, need_dx)
        fresh = pre is None or tuple(pre.shape) != (cols, K)
        d_W = torch.empty((cols, K), dtype=torch.float32, device=dev) if fresh else pre
        if fresh and split > 1:
            d_W.zero_()
        if concurrent or not fresh:
            # dW and dX only share their input d_h: dW runs on a second stream (half the SMs each), joined before returning
            # (a pre-zeroed target was filled on that stream: the GEMM follows it there in stream order)
            cur, fork = torch.cuda.current_stream(), _aux_stream(dev)
            fork.wait_stream(cur)
            with torch.cuda.stream(fork):
                gemm3x(d_h, True, x2, True, cols, K, rows, split_k=split, out=d_W)
        else:
            gemm3x(d_h, True, x2, True, cols, K, rows, split_k=split, out=d_W)
    else:
        d_W = torch.mm(d_h.t(), x2)
    d_x = None
    if need_dx:
        ld = _pad4(K)
        buf = torch.empty((rows, ld), dtype=torch.float32, device=dev)
        d_x = buf if ld == K else buf[:, :K]
        if use_tc:  # dX[rows, K] = d_h W: W (cols, K) is the MN-major B operand as stored
            gemm3x(d_h, False, Wp, True, rows, K, cols, out=buf)
        else:
            torch.mm(d_h, W, out=d_x)
    if fork is not None:
        torch.cuda.current_stream().wait_stream(fork)
    return d_x, d_W


def _fused_bn_ok(rows, cols, training, head):
    from . import config
    return bool(training and config.fused_bn and cols % 4 == 0 and _lib.lib().rh_bn_fused_supported(rows, cols, int(head)))


class _TowerLayer(torch.autograd.Function):
    """y = dropout(act(bn(x @ W^T + b))): one tower layer (basic/layers.py:282-285).

    The GEMMs run on the tensor cores through rh_gemm_tf32x3 (fp32-accurate 3xTF32; batches under 128 rows use torch.mm).
    Training-mode BatchNorm + activation + dropout is ONE launch each way when the layer fits the fused kernels
    (rh_bn_act_fused_fwd / _bwd: rows in registers across a grid barrier); otherwise rh_colstats + rh_bn_act_fwd and the
    two passes of rh_bn_act_bwd.  No activation, mask or normalised copy is stored: backward recomputes from h.
    """

    @staticmethod
    def forward(ctx, x, W, b, gamma, beta, act_param, cfg):
        L = _lib.lib()
        x2 = _rowmajor(x if x.dtype == torch.float32 else x.float())
        rows, K = x2.shape
        cols = W.shape[0]
        dev = x2.device
        st = stream_ptr()
        training = cfg["training"]
        h, Wp, use_tc = _linear_fwd(x2, W, b)
        _prezero_dw(ctx, x2, W, use_tc, ctx.needs_input_grad[0])
        p = float(cfg["p_drop"])
        y = torch.empty((rows, cols), dtype=torch.float32, device=dev)
        fused = _fused_bn_ok(rows, cols, training, False)
        stats = None
        if fused:
            stats = torch.empty(2 * cols + 1, dtype=torch.float32, device=dev)
            check(
                L.rh_bn_act_fused_fwd(h.data_ptr(), cols, rows, cols, float(cfg["eps"]), ptr(gamma), ptr(beta), cfg["act"], ptr(act_param), float(cfg["dice_eps"]), p, cfg["seed"], ptr(cfg["running_mean"]),
                                      ptr(cfg["running_var"]), ptr(cfg["num_batches_tracked"]), float(cfg["momentum"]), stats.data_ptr(), _bn_fused_scratch(dev, cols).data_ptr(), y.data_ptr(), cols, None, None, None,
                                      None, 0, None, st), "rh_bn_act_fused_fwd")
            mean = var = None
        else:
            counter = None
            if training:
                stats = torch.empty(2 * cols + 1, dtype=torch.float32, device=dev)
                check(L.rh_colstats(h.data_ptr(), cols, rows, cols, stats.data_ptr(), _colstats_scratch(dev, cols).data_ptr(), ptr(cfg["running_mean"]), ptr(cfg["running_var"]), ptr(cfg["num_batches_tracked"]),
                                    float(cfg["momentum"]), st), "rh_colstats")
                mean, var, counter = stats[:cols], stats[cols:2 * cols], stats[2 * cols:]
            else:
                mean, var = cfg["running_mean"], cfg["running_var"]
            check(
                L.rh_bn_act_fwd(h.data_ptr(), cols, rows, cols, ptr(mean), ptr(var), float(cfg["eps"]), ptr(gamma), ptr(beta), cfg["act"], ptr(act_param), float(cfg["dice_eps"]), p, cfg["seed"], ptr(counter), y.data_ptr(),
                                cols, st), "rh_bn_act_fwd")
        ctx.cfg = cfg
        ctx.has_param = act_param is not None
        ctx.has_bias = b is not None
        ctx.use_tc = use_tc
        ctx.Wp = Wp
        ctx.fused = fused
        ctx.save_for_backward(x2, W, h, stats, mean if stats is None else None, var if stats is None else None, gamma, beta, act_param)
        return y

    @staticmethod
    def backward(ctx, d_y):
        L = _lib.lib()
        cfg = ctx.cfg
        x2, W, h, stats, rmean, rvar, gamma, beta, act_param = ctx.saved_tensors
        rows, K = x2.shape
        cols = W.shape[0]
        dev = h.device
        training = stats is not None
        g = _rowmajor(d_y)
        g_ld = g.stride(0) if rows > 1 else cols
        d_h = torch.empty((rows, cols), dtype=torch.float32, device=dev)
        if ctx.fused and g_ld % 4 == 0 and g.data_ptr() % 16 == 0:
            gbuf = torch.empty(3 * cols + 4, dtype=torch.float32, device=dev)  # every slice is written by the kernel
            d_gamma, d_beta, d_b, d_alpha = gbuf[:cols], gbuf[cols:2 * cols], gbuf[2 * cols:3 * cols], gbuf[3 * cols:3 * cols + 1]
            check(
                L.rh_bn_act_fused_bwd(h.data_ptr(), cols, rows, cols, stats.data_ptr(), float(cfg["eps"]), ptr(gamma), ptr(beta), cfg["act"], ptr(act_param), float(cfg["dice_eps"]), float(cfg["p_drop"]), cfg["seed"],
                                      g.data_ptr(), g_ld, None, None, None, 0, _bn_fused_scratch(dev, cols).data_ptr(), d_h.data_ptr(), cols, d_gamma.data_ptr(), d_beta.data_ptr(),
                                      d_alpha.data_ptr(), None, None, None, d_b.data_ptr(), stream_ptr()), "rh_bn_act_fused_bwd")
        else:
            if training:
                mean, var, counter = stats[:cols], stats[cols:2 * cols], stats[2 * cols:]
            else:
                mean, var, counter = rmean, rvar, None
            gbuf = torch.zeros(3 * cols + 1, dtype=torch.float32, device=dev)
            d_gamma, d_beta, d_b, d_alpha = gbuf[:cols], gbuf[cols:2 * cols], gbuf[2 * cols:3 * cols], gbuf[3 * cols:]
            check(
                L.rh_bn_act_bwd(h.data_ptr(), cols, rows, cols, ptr(mean), ptr(var), float(cfg["eps"]), ptr(gamma), ptr(beta), cfg["act"], ptr(act_param), float(cfg["dice_eps"]), float(cfg["p_drop"]), cfg["seed"],
                                ptr(counter), g.data_ptr(), g_ld, int(training), d_h.data_ptr(), cols, d_gamma.data_ptr(), d_beta.data_ptr(), d_alpha.data_ptr() if ctx.has_param else None, stream_ptr()), "rh_bn_act_bwd")
            if not training and ctx.has_bias:  # eval: BN is affine, the Linear bias sees sum_rows d_h = d_beta * gamma * rstd
                d_b = d_beta * (gamma if gamma is not None else 1.0) / torch.sqrt(var + cfg["eps"])
        # training: the bias in front of a batch-statistics BN has a gradient of exactly 0 (BN removes any per-column
        # shift); the reference's autograd produces rounding noise ~1e-8 there.  d_b is the zero slice of gbuf.
        pre, ctx.dw_pre = getattr(ctx, "dw_pre", None), None
        d_x, d_W = _linear_bwd(d_h, x2, W, ctx.Wp, ctx.use_tc, ctx.needs_input_grad[0], pre)
        return (d_x, d_W, d_b if ctx.has_bias else None, d_gamma if gamma is not None else None, d_beta if beta is not None else None, d_alpha.view_as(act_param) if ctx.has_param else None, None)


def _layer_cfg(bn, act_code, dice_eps, p_drop, training):
    return {
        "running_mean": bn.running_mean,
        "running_var": bn.running_var,
        "num_batches_tracked": bn.num_batches_tracked,
        "momentum": bn.momentum,
        "eps": bn.eps,
        "training": bool(training),
        "act": act_code,
        "dice_eps": dice_eps,
        "p_drop": float(p_drop),
        "seed": _dropout_seed(bn) if p_drop > 0 else 0,
    }


def tower_layer(x, linear, bn, act_code, act_param, dice_eps, p_drop, training):
    """One ``Linear -> BatchNorm1d -> activation -> Dropout`` group of the reference's MLP on CUDA."""
    return _TowerLayer.apply(x, linear.weight, linear.bias, bn.weight, bn.bias, act_param, _layer_cfg(bn, act_code, dice_eps, p_drop, training))


class _TowerLayerHead(torch.autograd.Function):
    """p = f(Linear(K2, 1)(dropout(act(bn(x @ W^T + b)))) + extras...): the LAST hidden layer of a tower together with the output
    layer and the model's tail (basic/layers.py:279-285 + e.g. models/ranking/deepfm.py:41-43), training mode.

    GEMM + ONE fused launch forward (column statistics, BatchNorm, activation, dropout, the head's dot product, side terms,
    sigmoid — the activation never reaches HBM), ONE fused launch + the two gradient GEMMs backward."""

    @staticmethod
    def forward(ctx, x, W, b, gamma, beta, act_param, head_W, head_b, cfg, apply_sigmoid, *extras):
        L = _lib.lib()
        x2 = _rowmajor(x if x.dtype == torch.float32 else x.float())
        rows, K = x2.shape
        cols = W.shape[0]
        dev = x2.device
        h, Wp, use_tc = _linear_fwd(x2, W, b)
        _prezero_dw(ctx, x2, W, use_tc, ctx.needs_input_grad[0])
        ex = [e.contiguous() for e in extras]
        stats = torch.empty(2 * cols + 1, dtype=torch.float32, device=dev)
        out = torch.empty(rows, dtype=torch.float32, device=dev)
        check(
            L.rh_bn_act_fused_fwd(h.data_ptr(), cols, rows, cols, float(cfg["eps"]), ptr(gamma), ptr(beta), cfg["act"], ptr(act_param), float(cfg["dice_eps"]), float(cfg["p_drop"]), cfg["seed"], ptr(cfg["running_mean"]),
                                  ptr(cfg["running_var"]), ptr(cfg["num_batches_tracked"]), float(cfg["momentum"]), stats.data_ptr(), _bn_fused_scratch(dev, cols).data_ptr(), None, 0, head_W.data_ptr(), ptr(head_b),
                                  ptr(ex[0]) if len(ex) > 0 else None, ptr(ex[1]) if len(ex) > 1 else None, int(apply_sigmoid), out.data_ptr(), stream_ptr()), "rh_bn_act_fused_fwd")
        ctx.cfg, ctx.sig, ctx.n_extra = cfg, bool(apply_sigmoid), len(ex)
        ctx.has_param, ctx.has_bias, ctx.has_head_bias = act_param is not None, b is not None, head_b is not None
        ctx.use_tc, ctx.Wp = use_tc, Wp
        ctx.save_for_backward(x2, W, h, stats, gamma, beta, act_param, head_W, out)
        return out

    @staticmethod
    def backward(ctx, d_out):
        L = _lib.lib()
        cfg = ctx.cfg
        x2, W, h, stats, gamma, beta, act_param, head_W, out = ctx.saved_tensors
        rows, K = x2.shape
        cols = W.shape[0]
        dev = h.device
        d_out = d_out.contiguous()
        d_h = torch.empty((rows, cols), dtype=torch.float32, device=dev)
        gbuf = torch.empty(4 * cols + 4, dtype=torch.float32, device=dev)  # every slice is written by the kernel
        d_gamma, d_beta, d_b, d_hw = gbuf[:cols], gbuf[cols:2 * cols], gbuf[2 * cols:3 * cols], gbuf[3 * cols:4 * cols]
        d_alpha, d_hb = gbuf[4 * cols:4 * cols + 1], gbuf[4 * cols + 1:4 * cols + 2]
        d_e = torch.empty(rows, dtype=torch.float32, device=dev) if ctx.n_extra else None
        check(
            L.rh_bn_act_fused_bwd(h.data_ptr(), cols, rows, cols, stats.data_ptr(), float(cfg["eps"]), ptr(gamma), ptr(beta), cfg["act"], ptr(act_param), float(cfg["dice_eps"]), float(cfg["p_drop"]), cfg["seed"], None, 0,
                                  head_W.data_ptr(), out.data_ptr(), d_out.data_ptr(), int(ctx.sig), _bn_fused_scratch(dev, cols).data_ptr(), d_h.data_ptr(), cols, d_gamma.data_ptr(), d_beta.data_ptr(),
                                  d_alpha.data_ptr(), d_hw.data_ptr(), d_hb.data_ptr(), ptr(d_e), d_b.data_ptr(), stream_ptr()), "rh_bn_act_fused_bwd")
        pre, ctx.dw_pre = getattr(ctx, "dw_pre", None), None
        d_x, d_W = _linear_bwd(d_h, x2, W, ctx.Wp, ctx.use_tc, ctx.needs_input_grad[0], pre)
        return (d_x, d_W, d_b if ctx.has_bias else None, d_gamma if gamma is not None else None, d_beta if beta is not None else None, d_alpha.view_as(act_param) if ctx.has_param else None,
                d_hw.view_as(head_W), d_hb if ctx.has_head_bias else None, None, None) + (d_e,) * ctx.n_extra


def tower_layer_head(x, linear, bn, act_code, act_param, dice_eps, p_drop, head_linear, extras=(), sigmoid=True):
    """Training-mode last hidden layer + output layer + tail as :class:`_TowerLayerHead`, or None when the shape is outside the
    fused kernel (callers then run ``tower_layer`` + ``output_head``)."""
    from . import config
    Wh = head_linear.weight
    x2 = x
    if (not config.fused_bn_head or not x.is_cuda or x.dim() != 2 or Wh.shape[0] != 1 or Wh.dtype != torch.float32 or not Wh.is_contiguous() or Wh.data_ptr() % 16 != 0 or len(extras) > 2
            or any(e.dim() != 1 or e.shape[0] != x.shape[0] or e.dtype != torch.float32 for e in extras)):
        return None
    cols = linear.weight.shape[0]
    if Wh.shape[1] != cols or not _fused_bn_ok(x2.shape[0], cols, True, True):
        return None
    return _TowerLayerHead.apply(x, linear.weight, linear.bias, bn.weight, bn.bias, act_param, Wh, head_linear.bias, _layer_cfg(bn, act_code, dice_eps, p_drop, True), bool(sigmoid), *extras)


# =====================================================================================================
# DIN target attention pieces
# =====================================================================================================
class _Head(torch.autograd.Function):
    """p = f(x @ w^T + b + extras...), f = sigmoid or identity: the tower's ``Linear(K, 1)`` output layer fused with the model's
    tail (reference basic/layers.py:279-280 + e.g. models/ranking/deepfm.py:41-43).  One launch forward, one backward."""

    @staticmethod
    def forward(ctx, x, W, b, apply_sigmoid, *extras):
        L = _lib.lib()
        x2 = _rowmajor(x)
        rows, K = x2.shape
        ex = [e.contiguous() for e in extras]
        out = torch.empty(rows, dtype=torch.float32, device=x2.device)
        check(L.rh_head_fwd(x2.data_ptr(), x2.stride(0) if rows > 1 else K, rows, K, W.data_ptr(), ptr(b), ptr(ex[0]) if len(ex) > 0 else None, ptr(ex[1]) if len(ex) > 1 else None, int(apply_sigmoid),
                            out.data_ptr(), stream_ptr()), "rh_head_fwd")
        ctx.sig, ctx.n_extra, ctx.has_bias = bool(apply_sigmoid), len(ex), b is not None
        ctx.save_for_backward(x2, W, out)
        return out

    @staticmethod
    def backward(ctx, d_out):
        L = _lib.lib()
        x2, W, out = ctx.saved_tensors
        rows, K = x2.shape
        dev = x2.device
        d_out = d_out.contiguous()
        gbuf = torch.zeros(K + 1, dtype=torch.float32, device=dev)  # [d_w | d_b]
        need_x = ctx.needs_input_grad[0]
        d_x = torch.empty((rows, K), dtype=torch.float32, device=dev) if need_x else None
        d_e = torch.empty(rows, dtype=torch.float32, device=dev) if ctx.n_extra else None
        check(L.rh_head_bwd(x2.data_ptr(), x2.stride(0) if rows > 1 else K, rows, K, W.data_ptr(), out.data_ptr(), d_out.data_ptr(), int(ctx.sig), ptr(d_x), K, gbuf.data_ptr(),
                            gbuf.data_ptr() + 4 * K, ptr(d_e), stream_ptr()), "rh_head_bwd")
        return (d_x, gbuf[:K].view_as(W), gbuf[K:K + 1] if ctx.has_bias else None, None) + (d_e,) * ctx.n_extra


def output_head(x, linear, extras=(), sigmoid=True):
    """``f(linear(x).squeeze(1) + sum(extras))`` for a ``Linear(K, 1)`` on CUDA, or None when the shape is outside the kernel."""
    from . import config
    W = linear.weight
    if (not config.fused_head or not x.is_cuda or x.dim() != 2 or x.dtype != torch.float32 or W.shape[0] != 1 or W.shape[1] > 1024 or W.dtype != torch.float32 or len(extras) > 2
            or not W.is_contiguous() or any(e.dim() != 1 or e.shape[0] != x.shape[0] or e.dtype != torch.float32 for e in extras)):
        return None
    return _Head.apply(x, W, linear.bias, sigmoid, *extras)


class _DinAttnInput(torch.autograd.Function):
    """(att_in (B*L, 4D), hist (B, L, D), target (B, D)) from the two tables; backward ends in the scatter-add."""

    @staticmethod
    def forward(ctx, hist_w, tgt_w, hist_ids, tgt_ids, hist_pad, tgt_pad):
        L = _lib.lib()
        dev = hist_w.device
        hi = _as_ids(hist_ids).contiguous()
        ti = _as_ids(tgt_ids)
        if ti.dtype != hi.dtype:
            ti = ti.to(hi.dtype)
        B, S = hi.shape
        D = hist_w.shape[1]
        att_in = torch.empty((B * S, 4 * D), dtype=torch.float32, device=dev)
        hist = torch.empty((B, S, D), dtype=torch.float32, device=dev)
        tgt = torch.empty((B, D), dtype=torch.float32, device=dev)
        t_stride = ti.stride(0) if B > 1 else 1
        check(
            L.rh_din_attn_input_fwd(hist_w.data_ptr(), hist_w.shape[0], tgt_w.data_ptr(), tgt_w.shape[0], D, hi.data_ptr(), ti.data_ptr(), int(hi.dtype == torch.int32), t_stride, B, S, att_in.data_ptr(),
                                    hist.data_ptr(), tgt.data_ptr(), _lib.err_flag(dev).data_ptr(), stream_ptr()), "rh_din_attn_input_fwd")
        ctx.hist_w, ctx.tgt_w, ctx.hi, ctx.ti = hist_w, tgt_w, hi, ti
        ctx.pads = (-1 if hist_pad is None else int(hist_pad), -1 if tgt_pad is None else int(tgt_pad))
        ctx.t_stride = t_stride
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(hist, tgt)
        return att_in, hist, tgt

    @staticmethod
    def backward(ctx, d_att_in, d_hist, d_tgt):
        L = _lib.lib()
        hist, tgt = ctx.saved_tensors
        B, S, D = hist.shape
        dev = hist.device
        if d_att_in is None:
            d_att_in = torch.zeros((B * S, 4 * D), dtype=torch.float32, device=dev)
        gh, hslot = _table.grad_target(ctx.hist_w)
        gt, tslot = _table.grad_target(ctx.tgt_w)
        check(
            L.rh_din_attn_input_bwd(ptr(gh), ctx.hist_w.shape[0], ctx.pads[0], ptr(gt), ctx.tgt_w.shape[0], ctx.pads[1], D, ctx.hi.data_ptr(), ctx.ti.data_ptr(), int(ctx.hi.dtype == torch.int32), ctx.t_stride, B, S,
                                    hist.data_ptr(), tgt.data_ptr(), d_att_in.contiguous().data_ptr(), ptr(d_hist.contiguous()) if d_hist is not None else None,
                                    ptr(d_tgt.contiguous()) if d_tgt is not None else None, _lib.err_flag(dev).data_ptr(), stream_ptr()), "rh_din_attn_input_bwd")
        if gh is not None:
            _table.note_dirty(hslot, ctx.hi)
        if gt is not None:
            _table.note_dirty(tslot, ctx.ti)
        return None, None, None, None, None, None


def din_attention_input(hist_table, tgt_table, hist_ids, tgt_ids):
    return _DinAttnInput.apply(hist_table.weight, tgt_table.weight, hist_ids, tgt_ids, hist_table.padding_idx, tgt_table.padding_idx)


class _DinWeightedSum(torch.autograd.Function):

    @staticmethod
    def forward(ctx, att_w, hist, use_softmax):
        L = _lib.lib()
        B, S, D = hist.shape
        w = att_w.contiguous()
        h = hist.contiguous()
        out = torch.empty((B, D), dtype=torch.float32, device=hist.device)
        w_used = torch.empty((B, S), dtype=torch.float32, device=hist.device) if use_softmax else None
        check(L.rh_din_weighted_sum_fwd(w.data_ptr(), h.data_ptr(), B, S, D, int(use_softmax), ptr(w_used), out.data_ptr(), stream_ptr()), "rh_din_weighted_sum_fwd")
        ctx.use_softmax = use_softmax
        ctx.save_for_backward(w_used if use_softmax else w, h)
        return out

    @staticmethod
    def backward(ctx, d_out):
        L = _lib.lib()
        w_used, h = ctx.saved_tensors
        B, S, D = h.shape
        d_w = torch.empty((B, S), dtype=torch.float32, device=h.device)
        d_h = torch.empty_like(h)
        check(L.rh_din_weighted_sum_bwd(w_used.data_ptr(), h.data_ptr(), d_out.contiguous().data_ptr(), B, S, D, int(ctx.use_softmax), d_w.data_ptr(), d_h.data_ptr(), stream_ptr()), "rh_din_weighted_sum_bwd")
        return d_w, d_h, None


def din_weighted_sum(att_w, hist, use_softmax):
    """``sum_l w[b,l] * hist[b,l,:]`` with optional softmax over l (din.py:86-92)."""
    return _DinWeightedSum.apply(att_w, hist, bool(use_softmax))


# =====================================================================================================
# two-tower in-batch negatives (MatchTrainer's in-batch branch)
# =====================================================================================================
class _ScoresNT(torch.autograd.Function):
    """scores = U V^T (B x B) on the tensor cores (rh_gemm_tf32x3); backward: d_U = dS V, d_V = dS^T U with dS read as stored."""

    @staticmethod
    def forward(ctx, u, v):
        u2, v2 = _padded_rows(u), _padded_rows(v)
        B, D = u2.shape
        out = _buf(B, v2.shape[0], u2.device)
        gemm3x(u2, False, v2, False, B, v2.shape[0], D, out=out)
        ctx.save_for_backward(u2, v2)
        return out

    @staticmethod
    def backward(ctx, d_s):
        u2, v2 = ctx.saved_tensors
        B, D = u2.shape
        N = v2.shape[0]
        g = _padded_rows(d_s)
        d_u = _mm(g, False, v2, True, B, D, N, True)  # d_U[B, D] = dS[B, N] V[N, D]: V is the MN-major operand as stored
        d_v = _mm(g, True, u2, True, N, D, B, True, split_k=_split_k_for((N + 127) // 128, 1, (B + 31) // 32, budget=128))  # d_V = dS^T U
        return d_u, d_v


def scores_nt(u, v):
    """``u @ v.T`` for the in-batch score matrix, or None when the shape is outside the tensor-core kernel."""
    if not (u.is_cuda and u.dim() == 2 and v.dim() == 2 and u.dtype == torch.float32 and u.shape[0] >= 128 and v.shape[0] >= 128 and u.shape[1] >= 32 and _tc_ok(u.shape[0], _padded_rows(u.detach()))):
        return None
    return _ScoresNT.apply(u, v)


def inbatch_sample(scores, k, hard, seed_dev):
    """(B, k) int64 picks by the engine's samplers (``scores`` is only read for hard negatives), or None outside their range."""
    B = scores.shape[0]
    if not scores.is_cuda or (not hard and k != B - 1 and k > 128):
        return None
    L = _lib.lib()
    picks = torch.empty((B, k), dtype=torch.int64, device=scores.device)
    if hard:
        s2 = _rowmajor(scores.detach())
        check(L.rh_inbatch_sample_hard(s2.data_ptr(), s2.stride(0), B, k, picks.data_ptr(), stream_ptr()), "rh_inbatch_sample_hard")
    else:
        check(L.rh_inbatch_sample_random(B, k, seed_dev.data_ptr(), picks.data_ptr(), stream_ptr()), "rh_inbatch_sample_random")
    return picks


class _InbatchCE(torch.autograd.Function):
    """mean cross entropy over [positive | picks] with logits = <u_i, v_c> (rh_inbatch_ce_fwd / _bwd): no gathered logits tensor, no
    dense (B, B) score gradient."""

    @staticmethod
    def forward(ctx, u, v, picks):
        L = _lib.lib()
        u2, v2 = _rowmajor(u), _rowmajor(v)
        B, D = u2.shape
        K = picks.shape[1]
        prob = torch.empty((B, K + 1), dtype=torch.float32, device=u2.device)
        rows = torch.empty(B, dtype=torch.float32, device=u2.device)
        check(L.rh_inbatch_ce_fwd(u2.data_ptr(), u2.stride(0), v2.data_ptr(), v2.stride(0), D, picks.data_ptr(), B, K, prob.data_ptr(), rows.data_ptr(), stream_ptr()), "rh_inbatch_ce_fwd")
        ctx.save_for_backward(u2, v2, picks, prob)
        return rows.mean()

    @staticmethod
    def backward(ctx, d_loss):
        L = _lib.lib()
        u2, v2, picks, prob = ctx.saved_tensors
        B, D = u2.shape
        K = picks.shape[1]
        d_u = torch.empty((B, D), dtype=torch.float32, device=u2.device)
        d_v = torch.zeros((v2.shape[0], D), dtype=torch.float32, device=u2.device)
        g = d_loss.detach().reshape(1).float().contiguous()
        check(L.rh_inbatch_ce_bwd(u2.data_ptr(), u2.stride(0), v2.data_ptr(), v2.stride(0), D, picks.data_ptr(), prob.data_ptr(), g.data_ptr(), B, K, d_u.data_ptr(), D, d_v.data_ptr(), D, stream_ptr()),
              "rh_inbatch_ce_bwd")
        return d_u, d_v, None


def inbatch_cross_entropy(u, v, picks):
    """CrossEntropyLoss(mean) over ``[<u_i, v_i> | <u_i, v_picks[i, :]>]`` (positive in column 0), fused; None outside the kernel's range."""
    if not (u.is_cuda and u.dim() == 2 and v.dim() == 2 and u.dtype == torch.float32 and v.dtype == torch.float32 and u.shape[1] == v.shape[1] and u.shape[1] <= 256 and u.shape[0] == v.shape[0]):
        return None
    return _InbatchCE.apply(u, v, picks.contiguous())


# =====================================================================================================
# BCELoss(mean) on probabilities: the CTR trainer's criterion
# =====================================================================================================
_bce_scratch = {}


class _BCEMean(torch.autograd.Function):

    @staticmethod
    def forward(ctx, prob, target):
        L = _lib.lib()
        p, y = prob.contiguous(), target.contiguous()
        dev = p.device
        key = dev.index if dev.index is not None else torch.cuda.current_device()
        sc = _bce_scratch.get(key)
        if sc is None:
            sc = _bce_scratch[key] = torch.zeros(65, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        check(L.rh_bce_fwd(p.data_ptr(), y.data_ptr(), p.numel(), sc.data_ptr(), loss.data_ptr(), stream_ptr()), "rh_bce_fwd")
        ctx.save_for_backward(p, y)
        return loss

    @staticmethod
    def backward(ctx, d_loss):
        L = _lib.lib()
        p, y = ctx.saved_tensors
        d_p = torch.empty_like(p)
        g = d_loss.reshape(1).float().contiguous()
        check(L.rh_bce_bwd(p.data_ptr(), y.data_ptr(), g.data_ptr(), p.numel(), d_p.data_ptr(), stream_ptr()), "rh_bce_bwd")
        return d_p, None


class EngineBCELoss(torch.nn.BCELoss):
    """``torch.nn.BCELoss`` whose mean reduction over fp32 CUDA probabilities is one launch each way (``rh_bce_fwd`` / ``rh_bce_bwd``);
    any other input (CPU, weights, other reductions, dtypes) takes the stock implementation."""

    def forward(self, input, target):
        if (input.is_cuda and self.reduction == "mean" and self.weight is None and input.dtype == torch.float32 and target.dtype == torch.float32 and input.shape == target.shape and input.numel() > 0
                and not target.requires_grad):
            return _BCEMean.apply(input, target)
        return super().forward(input, target)
