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

    def __init__(self, weight, ids, padding_idx, mode, tile_col):
        self.weight = weight
        self.ids = ids
        self.vocab, self.dim = weight.shape
        self.padding_idx = -1 if padding_idx is None else int(padding_idx)
        self.mask_id = -1 if padding_idx is None else int(padding_idx)  # InputMask rule, layers.py:154-157
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


_gemm_scratch = {}


def _gemm_stats_scratch(dev, rows, cols):
    """Per (device, shape) scratch of rh_gemm_tf32x3_stats: per-tile partial statistics + tickets; zeroed once, the kernel keeps
    its tickets zero.  Launches that share it are ordered on one stream (the tower's forward)."""
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), rows, cols)
    buf = _gemm_scratch.get(key)
    if buf is None:
        n = int(_lib.lib().rh_gemm_stats_scratch_floats(rows, cols))
        buf = _gemm_scratch[key] = torch.zeros(n, dtype=torch.float32, device=dev)
    return buf


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


class _TowerLayer(torch.autograd.Function):
    """y = dropout(act(bn(x @ W^T + b))): one tower layer (basic/layers.py:282-285) as GEMM + statistics + ONE fused pass.

    The GEMMs run on the tensor cores through rh_gemm_tf32x3 (fp32-accurate 3xTF32; batches under 128 rows use torch.mm);
    everything between them is rh_colstats / rh_bn_act_fwd / rh_bn_act_bwd.  No activation, mask or normalised copy is
    stored: backward recomputes from h.  In backward the weight-gradient GEMM runs on a second stream next to the
    input-gradient GEMM (config.concurrent_tower_bwd).
    """

    @staticmethod
    def forward(ctx, x, W, b, gamma, beta, act_param, cfg):
        L = _lib.lib()
        x2 = _rowmajor(x if x.dtype == torch.float32 else x.float())
        rows, K = x2.shape
        cols = W.shape[0]
        dev = x2.device
        st = stream_ptr()
        h = torch.empty((rows, cols), dtype=torch.float32, device=dev)
        Wp = None
        use_tc = cols % 4 == 0 and cols >= 32 and _tc_ok(rows, x2)
        training = cfg["training"]
        stats = None
        if use_tc:
            Wp = _padded_weight(W)  # (cols, K) view with a 16-byte row stride
            from . import config
            if training and config.gemm_colstats:  # GEMM + BatchNorm column statistics in ONE launch
                stats = torch.empty(2 * cols + 1, dtype=torch.float32, device=dev)
                check(L.rh_gemm_tf32x3_stats(x2.data_ptr(), x2.stride(0), 0, Wp.data_ptr(), Wp.stride(0), 0, h.data_ptr(), h.stride(0), rows, cols, K, ptr(b), stats.data_ptr(),
                                             _gemm_stats_scratch(dev, rows, cols).data_ptr(), ptr(cfg["running_mean"]), ptr(cfg["running_var"]), ptr(cfg["num_batches_tracked"]), float(cfg["momentum"]), st),
                      "rh_gemm_tf32x3_stats")
            else:
                gemm3x(x2, False, Wp, False, rows, cols, K, bias=b, out=h)
        elif b is not None:
            torch.addmm(b, x2, W.t(), out=h)
        else:
            torch.mm(x2, W.t(), out=h)
        counter = None
        if training and stats is not None:
            mean, var, counter = stats[:cols], stats[cols:2 * cols], stats[2 * cols:]
        elif training:
            stats = torch.empty(2 * cols + 1, dtype=torch.float32, device=dev)
            check(L.rh_colstats(h.data_ptr(), cols, rows, cols, stats.data_ptr(), _colstats_scratch(dev, cols).data_ptr(), ptr(cfg["running_mean"]), ptr(cfg["running_var"]), ptr(cfg["num_batches_tracked"]),
                                float(cfg["momentum"]), st), "rh_colstats")
            mean, var, counter = stats[:cols], stats[cols:2 * cols], stats[2 * cols:]
        else:
            stats = None
            mean, var = cfg["running_mean"], cfg["running_var"]
        p = float(cfg["p_drop"])
        y = torch.empty((rows, cols), dtype=torch.float32, device=dev)
        check(
            L.rh_bn_act_fwd(h.data_ptr(), cols, rows, cols, ptr(mean), ptr(var), float(cfg["eps"]), ptr(gamma), ptr(beta), cfg["act"], ptr(act_param), float(cfg["dice_eps"]), p, cfg["seed"], ptr(counter), y.data_ptr(), cols,
                            st), "rh_bn_act_fwd")
        ctx.cfg = cfg
        ctx.has_param = act_param is not None
        ctx.has_bias = b is not None
        ctx.use_tc = use_tc
        ctx.Wp = Wp
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
        if training:
            mean, var, counter = stats[:cols], stats[cols:2 * cols], stats[2 * cols:]
        else:
            mean, var, counter = rmean, rvar, None
        g = _rowmajor(d_y)
        g_ld = g.stride(0) if rows > 1 else cols
        d_h = torch.empty((rows, cols), dtype=torch.float32, device=dev)
        gbuf = torch.zeros(3 * cols + 1, dtype=torch.float32, device=dev)
        d_gamma, d_beta, d_b, d_alpha = gbuf[:cols], gbuf[cols:2 * cols], gbuf[2 * cols:3 * cols], gbuf[3 * cols:]
        check(
            L.rh_bn_act_bwd(h.data_ptr(), cols, rows, cols, ptr(mean), ptr(var), float(cfg["eps"]), ptr(gamma), ptr(beta), cfg["act"], ptr(act_param), float(cfg["dice_eps"]), float(cfg["p_drop"]), cfg["seed"],
                            ptr(counter), g.data_ptr(), g_ld, int(training), d_h.data_ptr(), cols, d_gamma.data_ptr(), d_beta.data_ptr(), d_alpha.data_ptr() if ctx.has_param else None, stream_ptr()), "rh_bn_act_bwd")
        from . import config
        fork = None
        if ctx.use_tc:  # dW[cols, K] = d_h^T x: both operands read as stored (MN-major), K = rows split over CTAs
            concurrent = config.concurrent_tower_bwd and ctx.needs_input_grad[0]
            split = _split_k_for((cols + 127) // 128, (K + 127) // 128, (rows + 31) // 32, budget=64 if concurrent else 128)
            d_W = torch.zeros((cols, K), dtype=torch.float32, device=dev) if split > 1 else torch.empty((cols, K), dtype=torch.float32, device=dev)
            if concurrent:
                # dW and dX only share their input d_h: dW runs on a second stream (half the SMs each), joined before returning
                cur, fork = torch.cuda.current_stream(), _aux_stream(dev)
                fork.wait_stream(cur)
                with torch.cuda.stream(fork):
                    gemm3x(d_h, True, x2, True, cols, K, rows, split_k=split, out=d_W)
            else:
                gemm3x(d_h, True, x2, True, cols, K, rows, split_k=split, out=d_W)
        else:
            d_W = torch.mm(d_h.t(), x2)
        if not training and ctx.has_bias:  # eval: BN is affine, the Linear bias sees sum_rows d_h = d_beta * gamma * rstd
            d_b = d_beta * (gamma if gamma is not None else 1.0) / torch.sqrt(var + cfg["eps"])
        # training: the bias in front of a batch-statistics BN has a gradient of exactly 0 (BN removes any per-column
        # shift); the reference's autograd produces rounding noise ~1e-8 there.  d_b stays the zero slice of gbuf.
        d_x = None
        if ctx.needs_input_grad[0]:
            ld = _pad4(K)
            buf = torch.empty((rows, ld), dtype=torch.float32, device=dev)
            d_x = buf if ld == K else buf[:, :K]
            if ctx.use_tc:  # dX[rows, K] = d_h W: W (cols, K) is the MN-major B operand as stored
                gemm3x(d_h, False, ctx.Wp, True, rows, K, cols, out=buf)
            else:
                torch.mm(d_h, W, out=d_x)
        if fork is not None:
            torch.cuda.current_stream().wait_stream(fork)
        return (d_x, d_W, d_b if ctx.has_bias else None, d_gamma if gamma is not None else None, d_beta if beta is not None else None, d_alpha.view_as(act_param) if ctx.has_param else None, None)


def tower_layer(x, linear, bn, act_code, act_param, dice_eps, p_drop, training):
    """One ``Linear -> BatchNorm1d -> activation -> Dropout`` group of the reference's MLP on CUDA."""
    cfg = {
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
    return _TowerLayer.apply(x, linear.weight, linear.bias, bn.weight, bn.bias, act_param, cfg)


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
