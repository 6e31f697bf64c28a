"""TEST INFRASTRUCTURE ONLY — numpy restatement (forward AND hand-derived backward) of the reference's CTR hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may import this file; the product
(``torch-rechub_b200/``) never does.  Every function cites the reference lines it restates (paths relative to the
reference root, torch-rechub v0.8.0 @ 5a73b2e).  No torch: plain numpy, float64 by default, so it is an
independent check of both the CUDA kernels and the package's torch composite path.

Parity status: the reference's own tests hold NO numeric golden vectors for this path (SURVEY.md §8c,
``tests/test_e2e_ranking.py:106-107`` asserts only 0 <= AUC <= 1).  The oracle is therefore pinned against outputs of
the live reference itself: ``tests/golden/*.npz`` (generated in the build container by ``tests/golden/make_golden.py``
importing ``/root/reference``) hold state_dicts, inputs, logits and dense table gradients of the reference's
DeepFM / DCN / DCNv2 / DIN, and ``tests/test_oracle.py`` checks this file against them to 1e-6.

Conventions: ``sd`` is a model ``state_dict`` as numpy arrays with the reference's key layout;
``x`` maps feature name -> numpy array; models return a dict with ``logit`` (pre-sigmoid), ``prob`` and — for the
``*_forward_backward`` functions — ``grads`` keyed like ``sd`` for BCELoss(mean) against ``y``.
Dropout is not modelled (p = 0 / eval only): the reference's RNG stream cannot be reproduced.
"""
import numpy as np

BN_EPS = 1e-5  # torch.nn.BatchNorm1d default
DICE_EPS = 1e-3  # basic/activation.py:10


def _f(a, dtype):
    return np.asarray(a, dtype=dtype)


# ---------------------------------------------------------------------------------------------------------
# primitives
# ---------------------------------------------------------------------------------------------------------
def embedding_lookup(weight, ids):
    """nn.Embedding forward: rows of ``weight`` (basic/layers.py:83,85).  Float ids truncate like ``.long()``."""
    idx = np.asarray(ids).astype(np.int64)
    if idx.size and (idx.min() < 0 or idx.max() >= weight.shape[0]):
        raise IndexError("index out of range in self")
    return weight[idx]


class CompactGrad(object):
    """A dense (V, D) table gradient stored as its non-zero rows: ``ids`` (sorted, unique) and ``rows``.  Same numbers as the
    dense form (``dense()``); used for full-size tables (26 x 1M x 16) where 52 dense float64 gradients would not fit the
    "oracle finishes in seconds" budget.  Supports ``+`` with 0 and with another CompactGrad (two lookups of one table)."""

    def __init__(self, shape, ids, rows):
        self.shape, self.ids, self.rows = tuple(shape), ids, rows

    def __add__(self, other):
        if isinstance(other, (int, float)) and other == 0:
            return self
        ids = np.concatenate([self.ids, other.ids])
        rows = np.concatenate([self.rows, other.rows])
        u, inv = np.unique(ids, return_inverse=True)
        out = np.zeros((len(u), self.shape[1]), dtype=rows.dtype)
        np.add.at(out, inv, rows)
        return CompactGrad(self.shape, u, out)

    __radd__ = __add__

    def dense(self):
        g = np.zeros(self.shape, dtype=self.rows.dtype)
        g[self.ids] = self.rows
        return g


COMPACT_TABLE_GRADS = False  # tests at full table size switch this on (see ``compact_table_grads``)


class compact_table_grads(object):
    """``with compact_table_grads():`` -> table gradients come back as :class:`CompactGrad` instead of dense arrays."""

    def __enter__(self):
        global COMPACT_TABLE_GRADS
        self.prev, COMPACT_TABLE_GRADS = COMPACT_TABLE_GRADS, True

    def __exit__(self, *a):
        global COMPACT_TABLE_GRADS
        COMPACT_TABLE_GRADS = self.prev


def embedding_grad(weight_shape, ids, d_out, padding_idx=None, dtype=np.float64):
    """aten::embedding_dense_backward: dense (V, D) gradient, duplicates accumulate, padding row stays 0."""
    idx = np.asarray(ids).astype(np.int64).reshape(-1)
    d = np.asarray(d_out, dtype=dtype).reshape(-1, weight_shape[1])
    if padding_idx is not None:
        keep = idx != padding_idx
        idx, d = idx[keep], d[keep]
    if COMPACT_TABLE_GRADS:
        u, inv = np.unique(idx, return_inverse=True)
        rows = np.zeros((len(u), weight_shape[1]), dtype=dtype)
        np.add.at(rows, inv, d)
        return CompactGrad(weight_shape, u, rows)
    g = np.zeros(weight_shape, dtype=dtype)
    np.add.at(g, idx, d)
    return g


def table_rows(sd, key, ids, dtype=np.float64):
    """Rows ``ids`` of the table ``sd[key]`` in ``dtype``: the lookup first, the widening second (a 1M-row table is never
    converted whole)."""
    return _f(embedding_lookup(np.asarray(sd[key]), ids), dtype)


# ---------------------------------------------------------------------------------------------------------
# optimiser step (trainers/ctr_trainer.py:60-61,99: ``optimizer_fn(params, lr=1e-3, weight_decay=1e-5)`` -> torch.optim.Adam)
# ---------------------------------------------------------------------------------------------------------
def adam_update(w, g, m, v, step, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """One torch.optim.Adam step (L2 ``weight_decay`` added to the gradient, bias corrections by ``step`` >= 1) on arrays of any
    shape; returns (w, m, v).  float64 throughout.  Applied to the rows a batch touched it restates the row-wise (lazy) mode:
    untouched rows keep w, m, v while ``step`` is the global step count."""
    g = g + weight_decay * w
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1**step
    bc2 = 1 - beta2**step
    w = w - (lr / bc1) * m / (np.sqrt(v) / np.sqrt(bc2) + eps)
    return w, m, v


def fm_forward(e, reduce_sum=True):
    """FM.forward (basic/layers.py:313-319): 0.5 * sum_d[(sum_f e)^2 - sum_f e^2]; e: (B, F, D)."""
    s = e.sum(axis=1)
    ix = s * s - (e * e).sum(axis=1)
    if reduce_sum:
        ix = ix.sum(axis=1, keepdims=True)
    return 0.5 * ix


def fm_backward(e, d_y, reduce_sum=True):
    """d/de of fm_forward: d_y * (sum_f e - e)."""
    s = e.sum(axis=1, keepdims=True)
    g = d_y[:, None, :] if not reduce_sum else d_y.reshape(-1, 1, 1)
    return g * (s - e)


def input_mask(ids, padding_idx):
    """InputMask.forward (basic/layers.py:148-161): ids != padding_idx, or != -1 when the feature has none."""
    return (np.asarray(ids).astype(np.int64) != (padding_idx if padding_idx is not None else -1))


def seq_pool(weight, ids, pooling, padding_idx, dtype=np.float64):
    """SumPooling / AveragePooling with the mask (basic/layers.py:209-251): mean = bmm(mask, x) / (sum(mask) + 1e-16)."""
    e = embedding_lookup(weight, ids).astype(dtype)  # (B, L, D)
    m = input_mask(ids, padding_idx).astype(dtype)  # (B, L)
    pooled = (e * m[:, :, None]).sum(axis=1)
    if pooling == "sum":
        return pooled
    if pooling == "mean":
        return pooled / (m.sum(axis=1, keepdims=True) + 1e-16)
    raise ValueError(pooling)


def dice_forward(x, alpha, eps=DICE_EPS):
    """Dice.forward (basic/activation.py:15-25): per-ROW mean / summed variance over the neuron axis."""
    avg = x.mean(axis=1, keepdims=True)
    var = ((x - avg)**2 + eps).sum(axis=1, keepdims=True)
    ps = 1.0 / (1.0 + np.exp(-(x - avg) / np.sqrt(var)))
    return ps * x + (1 - ps) * alpha * x, (avg, var, ps)


def dice_backward(x, alpha, cache, d_out):
    avg, var, ps = cache
    n = x.shape[1]
    s = np.sqrt(var)
    c = 1.0 - alpha
    a = d_out * x * c * ps * (1 - ps)
    a1 = a.sum(axis=1, keepdims=True)
    a2 = (a * (x - avg)).sum(axis=1, keepdims=True)
    d_x = d_out * (alpha + c * ps) + a / s - a1 / (n * s) - (x - avg) * a2 / s**3
    d_alpha = (d_out * x * (1 - ps)).sum()
    return d_x, d_alpha


def _act_forward(kind, z, param):
    if kind == "relu":
        return np.maximum(z, 0), None
    if kind == "dice":
        return dice_forward(z, param)
    if kind == "prelu":
        return np.where(z > 0, z, param * z), None
    if kind == "sigmoid":
        return 1 / (1 + np.exp(-z)), None
    if kind == "leakyrelu":
        return np.where(z > 0, z, 0.01 * z), None
    raise ValueError(kind)


def _act_backward(kind, z, param, cache, d_a):
    if kind == "relu":
        return d_a * (z > 0), None
    if kind == "dice":
        return dice_backward(z, param, cache, d_a)
    if kind == "prelu":
        return np.where(z > 0, d_a, param * d_a), (d_a * np.minimum(z, 0)).sum()
    if kind == "sigmoid":
        sg = 1 / (1 + np.exp(-z))
        return d_a * sg * (1 - sg), None
    if kind == "leakyrelu":
        return np.where(z > 0, d_a, 0.01 * d_a), None
    raise ValueError(kind)


def mlp_forward(sd, prefix, x, n_hidden, activation="relu", train=True, output_layer=True, dtype=np.float64):
    """MLP.forward (basic/layers.py:276-292): [Linear -> BatchNorm1d -> act -> Dropout(p=0)] * n_hidden (+ Linear(.,1)).
    ``prefix`` like ``"mlp.mlp."``; module indices 4i (Linear), 4i+1 (BN), 4i+2 (activation params)."""
    caches = []
    h = x
    for i in range(n_hidden):
        W, b = _f(sd[prefix + "%d.weight" % (4 * i)], dtype), _f(sd[prefix + "%d.bias" % (4 * i)], dtype)
        gamma, beta = _f(sd[prefix + "%d.weight" % (4 * i + 1)], dtype), _f(sd[prefix + "%d.bias" % (4 * i + 1)], dtype)
        lin = h @ W.T + b
        if train:
            mu, var = lin.mean(axis=0), lin.var(axis=0)  # biased variance, as BatchNorm1d normalises with
        else:
            mu, var = _f(sd[prefix + "%d.running_mean" % (4 * i + 1)], dtype), _f(sd[prefix + "%d.running_var" % (4 * i + 1)], dtype)
        rstd = 1.0 / np.sqrt(var + BN_EPS)
        xhat = (lin - mu) * rstd
        z = xhat * gamma + beta
        pkey = prefix + ("%d.alpha" % (4 * i + 2) if activation == "dice" else "%d.weight" % (4 * i + 2))
        param = _f(sd[pkey], dtype)[0] if activation in ("dice", "prelu") else None
        a, acache = _act_forward(activation, z, param)
        caches.append((h, W, lin, rstd, xhat, gamma, z, param, acache))
        h = a
    if output_layer:
        W, b = _f(sd[prefix + "%d.weight" % (4 * n_hidden)], dtype), _f(sd[prefix + "%d.bias" % (4 * n_hidden)], dtype)
        caches.append((h, W))
        h = h @ W.T + b
    return h, caches


def mlp_backward(prefix, caches, d_out, n_hidden, activation="relu", train=True, output_layer=True):
    grads = {}
    g = d_out
    if output_layer:
        h, W = caches[-1]
        grads[prefix + "%d.weight" % (4 * n_hidden)] = g.T @ h
        grads[prefix + "%d.bias" % (4 * n_hidden)] = g.sum(axis=0)
        g = g @ W
    for i in reversed(range(n_hidden)):
        h, W, lin, rstd, xhat, gamma, z, param, acache = caches[i]
        d_z, d_param = _act_backward(activation, z, param, acache, g)
        if d_param is not None:
            grads[prefix + ("%d.alpha" % (4 * i + 2) if activation == "dice" else "%d.weight" % (4 * i + 2))] = np.array([d_param])
        grads[prefix + "%d.weight" % (4 * i + 1)] = (d_z * xhat).sum(axis=0)
        grads[prefix + "%d.bias" % (4 * i + 1)] = d_z.sum(axis=0)
        if train:  # batch statistics take part in the gradient
            n = lin.shape[0]
            d_lin = gamma * rstd * (d_z - d_z.mean(axis=0) - xhat * (d_z * xhat).mean(axis=0))
        else:
            d_lin = d_z * gamma * rstd
        grads[prefix + "%d.weight" % (4 * i)] = d_lin.T @ h
        grads[prefix + "%d.bias" % (4 * i)] = d_lin.sum(axis=0)
        g = d_lin @ W
    return g, grads


def bce_and_dlogit(logit, y):
    """BCELoss(mean) on sigmoid(logit) (trainers/ctr_trainer.py:68,88) and dLoss/dlogit = (p - y) / B."""
    p = 1 / (1 + np.exp(-logit))
    eps = 1e-300
    loss = -np.mean(y * np.log(np.maximum(p, eps)) + (1 - y) * np.log(np.maximum(1 - p, eps)))
    return loss, (p - y) / logit.shape[0], p


# ---------------------------------------------------------------------------------------------------------
# EmbeddingLayer
# ---------------------------------------------------------------------------------------------------------
def embedding_tile(sd, x, sparse_names, dense_names, table_of=None, dtype=np.float64):
    """EmbeddingLayer.forward(squeeze_dim=True) (basic/layers.py:77-127): sparse block first (list order), dense appended."""
    table_of = table_of or {}
    embs = [table_rows(sd, "embedding.embed_dict.%s.weight" % table_of.get(n, n), x[n], dtype) for n in sparse_names]
    parts = [np.concatenate(embs, axis=1)] if embs else []
    dense = [_f(x[n], np.float32).astype(dtype).reshape(len(x[n]), -1) for n in dense_names]
    if dense:
        parts.append(np.concatenate(dense, axis=1))
    return np.concatenate(parts, axis=1), embs


# ---------------------------------------------------------------------------------------------------------
# models
# ---------------------------------------------------------------------------------------------------------
def deepfm_forward_backward(sd, x, y, dense_names, deep_sparse_names, fm_names, n_hidden, activation="relu", train=True, backward=True, dtype=np.float64):
    """DeepFM.forward (models/ranking/deepfm.py:34-43) + BCELoss backward.
    deep tile = [deep sparse embeddings..., dense values]; FM/LR over ``fm_names`` embeddings."""
    y = _f(y, dtype)
    tile, deep_embs = embedding_tile(sd, x, deep_sparse_names, dense_names, dtype=dtype)
    e_fm = np.stack([table_rows(sd, "embedding.embed_dict.%s.weight" % n, x[n], dtype) for n in fm_names], axis=1)  # (B, F, D)
    B, F, D = e_fm.shape
    lw, lb = _f(sd["linear.fc.weight"], dtype), _f(sd["linear.fc.bias"], dtype)
    y_lin = e_fm.reshape(B, F * D) @ lw.T + lb
    y_fm = fm_forward(e_fm)
    y_deep, caches = mlp_forward(sd, "mlp.mlp.", tile, n_hidden, activation, train, True, dtype)
    logit = (y_lin + y_fm + y_deep)[:, 0]
    loss, d_logit, prob = bce_and_dlogit(logit, y)
    out = {"logit": logit, "prob": prob, "loss": loss, "y_fm": y_fm[:, 0], "y_linear": y_lin[:, 0], "tile": tile}
    if not backward:
        return out
    g = d_logit[:, None]
    d_tile, grads = mlp_backward("mlp.mlp.", caches, g, n_hidden, activation, train, True)
    grads["linear.fc.weight"] = g.T @ e_fm.reshape(B, F * D)
    grads["linear.fc.bias"] = g.sum(axis=0)
    d_e = fm_backward(e_fm, g) + (g @ lw).reshape(B, F, D)
    for k, n in enumerate(fm_names):
        key = "embedding.embed_dict.%s.weight" % n
        grads[key] = grads.get(key, 0) + embedding_grad(sd[key].shape, x[n], d_e[:, k, :], dtype=dtype)
    col = 0
    for n, e in zip(deep_sparse_names, deep_embs):
        key = "embedding.embed_dict.%s.weight" % n
        w = e.shape[1]
        grads[key] = grads.get(key, 0) + embedding_grad(sd[key].shape, x[n], d_tile[:, col:col + w], dtype=dtype)
        col += w
    out["grads"] = grads
    return out


def cross_forward(x0, ws, bs):
    """CrossNetwork.forward (basic/layers.py:412-420): x <- x0 * (w_i . x) + b_i + x."""
    xs, x = [x0], x0
    for w, b in zip(ws, bs):
        x = x0 * (x @ w.reshape(-1, 1)) + b + x
        xs.append(x)
    return x, xs


def cross_backward(x0, ws, bs, xs, d_out):
    g, g_x0 = d_out, np.zeros_like(x0)
    d_ws, d_bs = [None] * len(ws), [None] * len(ws)
    for l in reversed(range(len(ws))):
        s = xs[l] @ ws[l].reshape(-1, 1)  # (B,1)
        t = (g * x0).sum(axis=1, keepdims=True)  # dL/ds
        d_bs[l] = g.sum(axis=0)
        d_ws[l] = (t * xs[l]).sum(axis=0)
        g_x0 = g_x0 + g * s
        g = g + t * ws[l].reshape(1, -1)
    return g + g_x0, d_ws, d_bs


def dcn_forward_backward(sd, x, y, dense_names, sparse_names, n_cross, n_hidden, activation="relu", train=True, backward=True, dtype=np.float64):
    """DCN.forward (models/ranking/dcn.py:32-38): cross(e) || MLP(e) -> cat -> LR -> sigmoid."""
    y = _f(y, dtype)
    tile, embs = embedding_tile(sd, x, sparse_names, dense_names, dtype=dtype)
    ws = [_f(sd["cn.w.%d.weight" % i], dtype).reshape(-1) for i in range(n_cross)]
    bs = [_f(sd["cn.b.%d" % i], dtype) for i in range(n_cross)]
    cn_out, xs = cross_forward(tile, ws, bs)
    mlp_out, caches = mlp_forward(sd, "mlp.mlp.", tile, n_hidden, activation, train, False, dtype)
    stack = np.concatenate([cn_out, mlp_out], axis=1)
    lw, lb = _f(sd["linear.fc.weight"], dtype), _f(sd["linear.fc.bias"], dtype)
    logit = (stack @ lw.T + lb)[:, 0]
    loss, d_logit, prob = bce_and_dlogit(logit, y)
    out = {"logit": logit, "prob": prob, "loss": loss, "cross_out": cn_out, "tile": tile}
    if not backward:
        return out
    g = d_logit[:, None]
    grads = {"linear.fc.weight": g.T @ stack, "linear.fc.bias": g.sum(axis=0)}
    d_stack = g @ lw
    W = tile.shape[1]
    d_tile_c, d_ws, d_bs = cross_backward(tile, ws, bs, xs, d_stack[:, :W])
    d_tile_m, mg = mlp_backward("mlp.mlp.", caches, d_stack[:, W:], n_hidden, activation, train, False)
    grads.update(mg)
    for i in range(n_cross):
        grads["cn.w.%d.weight" % i] = d_ws[i].reshape(1, -1)
        grads["cn.b.%d" % i] = d_bs[i]
    d_tile = d_tile_c + d_tile_m
    col = 0
    for n, e in zip(sparse_names, embs):
        key = "embedding.embed_dict.%s.weight" % n
        w = e.shape[1]
        grads[key] = grads.get(key, 0) + embedding_grad(sd[key].shape, x[n], d_tile[:, col:col + w], dtype=dtype)
        col += w
    out["grads"] = grads
    return out


def crossnetmix_forward(sd, prefix, x, n_layers, n_experts, dtype=np.float64):
    """CrossNetMix.forward (basic/layers.py:470-506)."""
    x0 = x[:, :, None]
    xl = x0
    gates = [_f(sd[prefix + "gating.%d.weight" % e], dtype) for e in range(n_experts)]
    for i in range(n_layers):
        U, V, C = (_f(sd[prefix + "%s.%d" % (k, i)], dtype) for k in ("u_list", "v_list", "c_list"))
        bias = _f(sd[prefix + "bias.%d" % i], dtype)
        outs, scores = [], []
        for e in range(n_experts):
            scores.append(xl[:, :, 0] @ gates[e].T)  # (B,1)
            v = np.tanh(V[e].T @ xl)  # (B, r, 1)
            v = np.tanh(C[e] @ v)
            uv = U[e] @ v  # (B, W, 1)
            outs.append((x0 * (uv + bias))[:, :, 0])
        outs = np.stack(outs, axis=2)  # (B, W, E)
        sc = np.stack(scores, axis=1)  # (B, E, 1)
        sc = np.exp(sc - sc.max(axis=1, keepdims=True))
        sc = sc / sc.sum(axis=1, keepdims=True)
        xl = outs @ sc + xl
    return xl[:, :, 0]


def crossnetmix_forward_backward(sd, prefix, x, n_layers, n_experts, d_out=None, dtype=np.float64):
    """CrossNetMix.forward (basic/layers.py:470-506) in 2-D form with its hand-derived backward.

    Per layer i and expert e:  g_e = x_l Wg_e^T;  t1 = tanh(x_l V_e);  t2 = tanh(t1 C_e^T);  o_e = x_0 * (t2 U_e^T + b_i);
    s = softmax_e(g);  x_{l+1} = sum_e s_e o_e + x_l.   The gating Linears are shared by all layers (``:466``), so their
    gradients accumulate across layers.  Returns ``(x_L, d_x, grads)`` (the last two None without ``d_out``)."""
    x0 = _f(x, dtype)
    gates = [_f(sd[prefix + "gating.%d.weight" % e], dtype) for e in range(n_experts)]  # (1, W) each
    layers, xl = [], x0
    for i in range(n_layers):
        U, V, C = (_f(sd[prefix + "%s.%d" % (k, i)], dtype) for k in ("u_list", "v_list", "c_list"))
        b = _f(sd[prefix + "bias.%d" % i], dtype)[:, 0]
        g = np.stack([xl @ gates[e][0] for e in range(n_experts)], axis=1)  # (B, E)
        t1 = [np.tanh(xl @ V[e]) for e in range(n_experts)]
        t2 = [np.tanh(t1[e] @ C[e].T) for e in range(n_experts)]
        uvb = [t2[e] @ U[e].T + b for e in range(n_experts)]
        o = [x0 * uvb[e] for e in range(n_experts)]
        sm = np.exp(g - g.max(axis=1, keepdims=True))
        sm /= sm.sum(axis=1, keepdims=True)
        nxt = sum(o[e] * sm[:, e:e + 1] for e in range(n_experts)) + xl
        layers.append((xl, U, V, C, t1, t2, uvb, o, sm))
        xl = nxt
    if d_out is None:
        return xl, None, None
    grads = {prefix + "gating.%d.weight" % e: np.zeros_like(gates[e]) for e in range(n_experts)}
    d_x0 = np.zeros_like(x0)
    d_next = _f(d_out, dtype)
    for i in reversed(range(n_layers)):
        xin, U, V, C, t1, t2, uvb, o, sm = layers[i]
        d_xl = d_next.copy()  # the residual
        d_s = np.stack([(d_next * o[e]).sum(axis=1) for e in range(n_experts)], axis=1)
        d_g = sm * (d_s - (sm * d_s).sum(axis=1, keepdims=True))
        dU, dV, dC = np.zeros_like(U), np.zeros_like(V), np.zeros_like(C)
        d_b = np.zeros(x0.shape[1], dtype=dtype)
        for e in range(n_experts):
            grads[prefix + "gating.%d.weight" % e] += (d_g[:, e] @ xin)[None, :]
            d_xl += d_g[:, e:e + 1] * gates[e]
            d_o = d_next * sm[:, e:e + 1]
            d_x0 += d_o * uvb[e]
            d_uv = d_o * x0
            d_b += d_uv.sum(axis=0)
            dU[e] = d_uv.T @ t2[e]
            d_c = (d_uv @ U[e]) * (1 - t2[e]**2)
            dC[e] = d_c.T @ t1[e]
            d_a = (d_c @ C[e]) * (1 - t1[e]**2)
            dV[e] = xin.T @ d_a
            d_xl += d_a @ V[e].T
        grads[prefix + "u_list.%d" % i], grads[prefix + "v_list.%d" % i], grads[prefix + "c_list.%d" % i] = dU, dV, dC
        grads[prefix + "bias.%d" % i] = d_b[:, None]
        d_next = d_xl
    return xl, d_next + d_x0, grads


def dcnv2_forward_backward(sd, x, y, dense_names, sparse_names, n_cross, n_hidden, n_experts=4, activation="relu", train=True, backward=True, dtype=np.float64):
    """DCNv2.forward, default ``parallel`` structure with CrossNetMix (models/ranking/dcn_v2.py:47-59) + BCELoss backward."""
    y = _f(y, dtype)
    tile, embs = embedding_tile(sd, x, sparse_names, dense_names, dtype=dtype)
    W = tile.shape[1]
    cross_out, _, _ = crossnetmix_forward_backward(sd, "crossnet.", tile, n_cross, n_experts, None, dtype)
    dnn_out, caches = mlp_forward(sd, "parallel_dnn.mlp.", tile, n_hidden, activation, train, False, dtype)
    final = np.concatenate([cross_out, dnn_out], axis=1)
    lw, lb = _f(sd["linear.fc.weight"], dtype), _f(sd["linear.fc.bias"], dtype)
    logit = (final @ lw.T + lb)[:, 0]
    loss, d_logit, prob = bce_and_dlogit(logit, y)
    out = {"logit": logit, "prob": prob, "loss": loss, "tile": tile}
    if not backward:
        return out
    g = d_logit[:, None]
    grads = {"linear.fc.weight": g.T @ final, "linear.fc.bias": g.sum(axis=0)}
    d_final = g @ lw
    _, d_tile_c, cg = crossnetmix_forward_backward(sd, "crossnet.", tile, n_cross, n_experts, d_final[:, :W], dtype)
    d_tile_m, mg = mlp_backward("parallel_dnn.mlp.", caches, d_final[:, W:], n_hidden, activation, train, False)
    grads.update(cg)
    grads.update(mg)
    d_tile = d_tile_c + d_tile_m
    col = 0
    for n, e in zip(sparse_names, embs):
        key = "embedding.embed_dict.%s.weight" % n
        w = e.shape[1]
        grads[key] = grads.get(key, 0) + embedding_grad(sd[key].shape, x[n], d_tile[:, col:col + w], dtype=dtype)
        col += w
    out["grads"] = grads
    return out


def dcnv2_forward(sd, x, dense_names, sparse_names, n_cross, n_hidden, n_experts=4, activation="relu", train=True, dtype=np.float64):
    """DCNv2.forward, default ``parallel`` structure with CrossNetMix (models/ranking/dcn_v2.py:47-59)."""
    tile, _ = embedding_tile(sd, x, sparse_names, dense_names, dtype=dtype)
    cross_out = crossnetmix_forward(sd, "crossnet.", tile, n_cross, n_experts, dtype)
    dnn_out, _ = mlp_forward(sd, "parallel_dnn.mlp.", tile, n_hidden, activation, train, False, dtype)
    final = np.concatenate([cross_out, dnn_out], axis=1)
    logit = (final @ _f(sd["linear.fc.weight"], dtype).T + _f(sd["linear.fc.bias"], dtype))[:, 0]
    return {"logit": logit, "prob": 1 / (1 + np.exp(-logit)), "tile": tile}


def din_forward_backward(sd, x, y, feature_names, history_names, target_names, shared_with, n_att_hidden, n_hidden, use_softmax=False, train=True, backward=True, dtype=np.float64):
    """DIN.forward + ActivationUnit.forward (models/ranking/din.py:38-55, 77-93), Dice everywhere, no padding mask.
    ``shared_with[h]`` names the table a history feature looks up."""
    y = _f(y, dtype)
    rows = lambda n: table_rows(sd, "embedding.embed_dict.%s.weight" % shared_with.get(n, n), x[n], dtype)
    e_feat = [rows(n) for n in feature_names]
    e_hist = [rows(n) for n in history_names]  # (B, L, D) each
    e_tgt = [rows(n) for n in target_names]
    pooled, att_caches = [], []
    for i, h in enumerate(e_hist):
        B, L, D = h.shape
        t = np.broadcast_to(e_tgt[i][:, None, :], (B, L, D))
        att_in = np.concatenate([t, h, t - h, t * h], axis=-1).reshape(B * L, 4 * D)
        w, caches = mlp_forward(sd, "attention_layers.%d.attention.mlp." % i, att_in, n_att_hidden, "dice", train, True, dtype)
        w = w.reshape(B, L)
        w_used = w
        if use_softmax:
            ex = np.exp(w - w.max(axis=1, keepdims=True))
            w_used = ex / ex.sum(axis=1, keepdims=True)
        pooled.append((w_used[:, :, None] * h).sum(axis=1))
        att_caches.append((caches, w_used, t, h))
    mlp_in = np.concatenate(pooled + e_tgt + e_feat, axis=1)
    out_, caches = mlp_forward(sd, "mlp.mlp.", mlp_in, n_hidden, "dice", train, True, dtype)
    logit = out_[:, 0]
    loss, d_logit, prob = bce_and_dlogit(logit, y)
    out = {"logit": logit, "prob": prob, "loss": loss, "mlp_in": mlp_in}
    if not backward:
        return out
    d_in, grads = mlp_backward("mlp.mlp.", caches, d_logit[:, None], n_hidden, "dice", train, True)
    D = e_tgt[0].shape[1]
    col = 0
    d_pooled = []
    for _ in pooled:
        d_pooled.append(d_in[:, col:col + D])
        col += D
    d_tgt = []
    for _ in e_tgt:
        d_tgt.append(d_in[:, col:col + D].copy())
        col += D
    d_feat = []
    for e in e_feat:
        d_feat.append(d_in[:, col:col + e.shape[1]])
        col += e.shape[1]

    def add(name, ids, d):
        key = "embedding.embed_dict.%s.weight" % shared_with.get(name, name)
        grads[key] = grads.get(key, 0) + embedding_grad(sd[key].shape, ids, d, dtype=dtype)

    for i, (caches_i, w_used, t, h) in enumerate(att_caches):
        B, L, Dh = h.shape
        d_w = (d_pooled[i][:, None, :] * h).sum(axis=2)  # (B, L)
        d_h = w_used[:, :, None] * d_pooled[i][:, None, :]
        if use_softmax:
            d_w = w_used * (d_w - (w_used * d_w).sum(axis=1, keepdims=True))
        d_att_in, g_i = mlp_backward("attention_layers.%d.attention.mlp." % i, caches_i, d_w.reshape(B * L, 1), n_att_hidden, "dice", train, True)
        grads.update(g_i)
        d_att_in = d_att_in.reshape(B, L, 4, Dh)
        d0, d1, d2, d3 = (d_att_in[:, :, k, :] for k in range(4))
        d_h = d_h + d1 - d2 + d3 * t
        d_tgt[i] += (d0 + d2 + d3 * h).sum(axis=1)
        add(history_names[i], x[history_names[i]], d_h)
    for n, d in zip(target_names, d_tgt):
        add(n, x[n], d)
    for n, d in zip(feature_names, d_feat):
        add(n, x[n], d)
    out["grads"] = grads
    return out


# ---------------------------------------------------------------------------------------------------------
# two-tower retrieval
# ---------------------------------------------------------------------------------------------------------
def l2_normalize(h, eps=1e-12):
    """F.normalize(h, p=2, dim=1) (models/matching/dssm.py:61,71): h / max(||h||_2, eps)."""
    n = np.maximum(np.sqrt((h * h).sum(axis=1, keepdims=True)), eps)
    return h / n, n


def l2_normalize_backward(u, n, d_u):
    """d/dh of h / ||h||: (d_u - u <u, d_u>) / ||h|| (rows whose norm was clamped are never produced by the towers)."""
    return (d_u - u * (u * d_u).sum(axis=1, keepdims=True)) / n


def tower_tile(sd, x, features, dtype=np.float64):
    """EmbeddingLayer.forward(squeeze_dim=True) over sparse AND pooled sequence features in list order (basic/layers.py:80-117).
    ``features``: ``("sparse", name, table)`` or ``("seq", name, table, pooling)`` with ``table`` the owning table's name."""
    parts = []
    for f in features:
        w = _f(sd["embedding.embed_dict.%s.weight" % f[2]], dtype)
        parts.append(embedding_lookup(w, x[f[1]]).astype(dtype) if f[0] == "sparse" else seq_pool(w, x[f[1]], f[3], None, dtype))
    return np.concatenate(parts, axis=1)


def tower_tile_backward(sd, x, features, d_tile, grads, dtype=np.float64):
    """Dense table gradients of tower_tile, accumulated into ``grads`` (several features may share one table)."""
    col = 0
    for f in features:
        key = "embedding.embed_dict.%s.weight" % f[2]
        shape = sd[key].shape
        d = d_tile[:, col:col + shape[1]]
        col += shape[1]
        ids = np.asarray(x[f[1]]).astype(np.int64)
        if f[0] == "sparse":
            g = embedding_grad(shape, ids, d, None, dtype)
        else:
            m = input_mask(ids, None).astype(dtype)  # (B, L)
            scale = m / (m.sum(axis=1, keepdims=True) + 1e-16) if f[3] == "mean" else m
            g = embedding_grad(shape, ids.reshape(-1), (scale[:, :, None] * d[:, None, :]).reshape(-1, shape[1]), None, dtype)
        grads[key] = grads.get(key, 0) + g
    return grads


def hard_negative_indices(scores, k):
    """utils/match.py:131-135: per row the k best-scoring columns other than the diagonal, best first."""
    masked = scores.copy()
    np.fill_diagonal(masked, -np.inf)
    return np.argsort(-masked, axis=1, kind="stable")[:, :k]


def dssm_forward_backward(sd, x, user_features, item_features, n_user_hidden, n_item_hidden, neg_ratio, activation="relu", train=True, backward=True, dtype=np.float64, neg_idx=None):
    """DSSM towers (models/matching/dssm.py:40-72) + MatchTrainer's in-batch branch with HARD negatives and cross entropy
    (trainers/match_trainer.py:118-140, utils/match.py:104-161).  Returns the tower embeddings, the point-wise probability
    ``sigmoid(<u, v>)``, the (B, B) scores, the sampled columns, the ``[positive | negatives]`` logits, the loss and — with
    ``backward`` — every parameter gradient of that loss.  ``neg_idx`` overrides the sampler (any (B, K) column choice: random
    negatives, or another implementation's hard negatives when two scores tie within rounding)."""
    tu = tower_tile(sd, x, user_features, dtype)
    ti = tower_tile(sd, x, item_features, dtype)
    hu, cu = mlp_forward(sd, "user_mlp.mlp.", tu, n_user_hidden, activation, train, output_layer=False, dtype=dtype)
    hi, ci = mlp_forward(sd, "item_mlp.mlp.", ti, n_item_hidden, activation, train, output_layer=False, dtype=dtype)
    u, nu = l2_normalize(hu)
    v, nv = l2_normalize(hi)
    scores = u @ v.T
    B = scores.shape[0]
    k = neg_ratio if (neg_ratio is not None and 0 < neg_ratio <= B - 1) else B - 1
    neg = hard_negative_indices(scores, k) if neg_idx is None else np.asarray(neg_idx).astype(np.int64)
    k = neg.shape[1]
    rows = np.arange(B)
    logits = np.concatenate([scores[rows, rows][:, None], scores[rows[:, None], neg]], axis=1)
    mx = logits.max(axis=1, keepdims=True)
    lse = mx[:, 0] + np.log(np.exp(logits - mx).sum(axis=1))
    loss = float(np.mean(lse - logits[:, 0]))  # CrossEntropyLoss(mean) with the positive in column 0
    out = {"user_emb": u, "item_emb": v, "prob": 1 / (1 + np.exp(-(u * v).sum(axis=1))), "scores": scores, "neg_idx": neg, "logits": logits, "loss": loss}
    if not backward:
        return out
    d_logits = np.exp(logits - lse[:, None])
    d_logits[:, 0] -= 1.0
    d_logits /= B
    d_scores = np.zeros_like(scores)
    d_scores[rows, rows] += d_logits[:, 0]
    np.add.at(d_scores, (rows[:, None].repeat(k, axis=1), neg), d_logits[:, 1:])
    d_hu = l2_normalize_backward(u, nu, d_scores @ v)
    d_hi = l2_normalize_backward(v, nv, d_scores.T @ u)
    d_tu, gu = mlp_backward("user_mlp.mlp.", cu, d_hu, n_user_hidden, activation, train, output_layer=False)
    d_ti, gi = mlp_backward("item_mlp.mlp.", ci, d_hi, n_item_hidden, activation, train, output_layer=False)
    grads = {}
    grads.update(gu)
    grads.update(gi)
    tower_tile_backward(sd, x, user_features, d_tu, grads, dtype)
    tower_tile_backward(sd, x, item_features, d_ti, grads, dtype)
    out["grads"] = grads
    return out
