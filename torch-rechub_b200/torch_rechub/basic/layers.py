"""Layers of the ranking hot path (mirror of reference ``torch_rechub/basic/layers.py:12-506``).

Same classes, constructor signatures, parameter names (= state_dict keys) and output-shape rules as the
reference.  Each ``forward`` has two arithmetic routes:

* CUDA tensors -> the hand-written sm_100a kernels of ``torch_rechub.b200`` (fused multi-field gather,
  FM, cross network, BN+activation+dropout ...).  No eager fallback: a missing library raises.
* CPU tensors -> the same sequence of stock torch ops as the reference (quick-start, ONNX export).

Out of scope here: CIN, SENET, BiLinear, MultiInterestSA, Capsule, FFM, CEN, HSTU*,
InteractingLayer — they are plain PyTorch compositions on top of ``EmbeddingLayer`` in the reference.
"""
import torch
import torch.nn as nn

from .activation import Dice, activation_layer
from .features import DenseFeature, SequenceFeature, SparseFeature

_POOL_MODES = {"sum": 1, "mean": 2}


class PredictionLayer(nn.Module):
    """Sigmoid for ``"classification"``, identity for ``"regression"`` (reference ``layers.py:12-30``)."""

    def __init__(self, task_type='classification'):
        super(PredictionLayer, self).__init__()
        if task_type not in ["classification", "regression"]:
            raise ValueError("task_type must be classification or regression")
        self.task_type = task_type

    def forward(self, x):
        return torch.sigmoid(x) if self.task_type == "classification" else x


class ConcatPooling(nn.Module):
    """Identity on ``(B, L, D)`` (reference ``layers.py:192-206``; the mask is ignored there too)."""

    def forward(self, x, mask=None):
        return x


class AveragePooling(nn.Module):
    """Masked mean over the sequence axis: ``bmm(mask, x) / (sum(mask) + 1e-16)`` (reference ``layers.py:209-229``)."""

    def forward(self, x, mask=None):
        if mask is None:
            return torch.mean(x, dim=1)
        pooled = torch.bmm(mask, x).squeeze(1)
        valid = mask.sum(dim=-1)
        return pooled / (valid.float() + 1e-16)


class SumPooling(nn.Module):
    """Masked sum over the sequence axis (reference ``layers.py:232-251``)."""

    def forward(self, x, mask=None):
        if mask is None:
            return torch.sum(x, dim=1)
        return torch.bmm(mask, x).squeeze(1)


_POOLERS = {"sum": SumPooling, "mean": AveragePooling, "concat": ConcatPooling}


class InputMask(nn.Module):
    """1.0 where an id is a real token: ``id != padding_idx`` (``!= -1`` when the feature has no padding_idx).

    Reference ``layers.py:130-161``.  Output ``(B, n_features)`` for sparse, ``(B, n_seq, L)`` for sequences.
    """

    def forward(self, x, features):
        if not isinstance(features, list):
            features = [features]
        masks = []
        for fea in features:
            if not isinstance(fea, (SparseFeature, SequenceFeature)):
                raise ValueError("Only SparseFeature or SequenceFeature support to get mask.")
            sentinel = fea.padding_idx if fea.padding_idx is not None else -1
            masks.append((x[fea.name].long() != sentinel).unsqueeze(1).float())
        return torch.cat(masks, dim=1)


def _all_cuda(tensors):
    return all(t.is_cuda for t in tensors)


class EmbeddingLayer(nn.Module):
    """Per-feature tables + the lookup/concat front end of every model (reference ``layers.py:33-127``).

    ``forward(x, features, squeeze_dim=False)``:
      * ``squeeze_dim=True``  -> ``(B, sum(embed dims) [+ dense widths])``, sparse block first, dense appended;
      * ``squeeze_dim=False`` -> ``(B, n_sparse, D)`` (or ``(B, n_seq, L, D)`` for ``pooling="concat"``); dense ignored.

    On CUDA the whole call is ONE fused launch per row width (``rh_fields_fwd``) plus one per pooled
    sequence feature, writing straight into the output tile; the backward is the scatter-add into the
    tables' dense gradient buffers.
    """

    def __init__(self, features):
        super().__init__()
        self.features = features
        self.embed_dict = nn.ModuleDict()
        self.n_dense = 0
        self.input_mask = InputMask()
        self._dist = None  # b200.dist.ShardedFront when the tables are sharded by field over several GPUs
        for fea in features:
            if fea.name in self.embed_dict:
                continue
            if isinstance(fea, (SparseFeature, SequenceFeature)):
                if fea.shared_with is None:
                    self.embed_dict[fea.name] = fea.get_embedding_layer()
            elif isinstance(fea, DenseFeature):
                self.n_dense += 1

    # -- helpers -----------------------------------------------------------------------------------
    def table_of(self, fea):
        return self.embed_dict[fea.name if fea.shared_with is None else fea.shared_with]

    def _on_cuda(self, x, features):
        for fea in features:
            return x[fea.name].is_cuda
        return False

    # -- reference-order composite (CPU tensors, and CUDA layouts the fused plan does not cover) ----
    def _forward_composite(self, x, features, squeeze_dim):
        sparse_emb, dense_values = [], []
        for fea in features:
            if isinstance(fea, SparseFeature):
                sparse_emb.append(self.table_of(fea)(x[fea.name].long()).unsqueeze(1))
            elif isinstance(fea, SequenceFeature):
                if fea.pooling not in _POOLERS:
                    raise ValueError("Sequence pooling method supports only pooling in %s, got %s." % (["sum", "mean"], fea.pooling))
                pooled = _POOLERS[fea.pooling]()(self.table_of(fea)(x[fea.name].long()), self.input_mask(x, fea))
                sparse_emb.append(pooled.unsqueeze(1))
            else:
                v = x[fea.name].float()
                dense_values.append(v if v.dim() > 1 else v.unsqueeze(1))
        dense = torch.cat(dense_values, dim=1) if dense_values else None
        sparse = torch.cat(sparse_emb, dim=1) if sparse_emb else None
        return self._shape_output(sparse, dense, squeeze_dim, features)

    @staticmethod
    def _shape_output(sparse, dense, squeeze_dim, features):
        if squeeze_dim:
            if dense is not None and sparse is None:
                return dense
            if dense is None and sparse is not None:
                return sparse.flatten(start_dim=1)
            if dense is not None and sparse is not None:
                return torch.cat((sparse.flatten(start_dim=1), dense), dim=1)
            raise ValueError("The input features can note be empty")
        if sparse is not None:
            return sparse
        raise ValueError("If keep the original shape:[batch_size, num_features, embed_dim], expected %s in feature list, got %s" % ("SparseFeatures", features))

    # -- fused CUDA plan ---------------------------------------------------------------------------
    def build_plan(self, x, features, with_dense=True):
        """TilePlan for ``features`` or None when a feature needs the composite route
        (``pooling="concat"``, ids that are not one column per sample, exotic dense dtypes)."""
        from ..b200 import ops
        batch, device = None, None
        plan_items = []
        col = 0
        dense_feas = []
        for fea in features:
            if isinstance(fea, SparseFeature):
                ids = x[fea.name]
                if ids.dim() != 1:
                    return None
                plan_items.append(("f", fea, ops._as_ids(ids), col))
                col += fea.embed_dim
            elif isinstance(fea, SequenceFeature):
                if fea.pooling not in _POOLERS:
                    raise ValueError("Sequence pooling method supports only pooling in %s, got %s." % (["sum", "mean"], fea.pooling))
                ids = x[fea.name]
                if fea.pooling == "concat" or ids.dim() != 2:
                    return None
                plan_items.append(("s", fea, ops._as_ids(ids).contiguous(), col))
                col += fea.embed_dim
            else:
                dense_feas.append(fea)
            if batch is None:
                batch, device = x[fea.name].shape[0], x[fea.name].device
        if batch is None:
            return None
        plan = ops.TilePlan(batch, device)
        plan.by_name = {}
        for kind, fea, ids, c in plan_items:
            tbl = self.table_of(fea)
            if tbl.weight.shape[1] != fea.embed_dim:
                return None
            if kind == "f":
                plan.fields.append(ops.FieldRef(tbl.weight, ids, tbl.padding_idx, c, -1))
                plan.by_name.setdefault(fea.name, plan.fields[-1])
            else:
                plan.seqs.append(ops.SeqRef(tbl.weight, ids, tbl.padding_idx, _POOL_MODES[fea.pooling], c, mask_id=fea.padding_idx))
        plan.n_sparse_cols = col
        if with_dense:
            for fea in dense_feas:
                v = x[fea.name]
                if v.dtype not in ops._DENSE_CODES:
                    v = v.float()
                if v.dim() == 1:
                    width = 1
                elif v.dim() == 2 and (v.stride(1) == 1 or v.shape[1] == 1):
                    width = v.shape[1]
                else:
                    v = v.float().reshape(batch, -1).contiguous()
                    width = v.shape[1]
                plan.dense.append(ops.DenseRef(v, width, col))
                col += width
        plan.tile_width = col
        plan.sparse_dims = [it[1].embed_dim for it in plan_items]
        return plan

    def forward(self, x, features, squeeze_dim=False):
        if self._dist is not None:
            return self._dist.forward(x, features, squeeze_dim)
        return self._forward_local(x, features, squeeze_dim)

    def _forward_local(self, x, features, squeeze_dim=False):
        if not self._on_cuda(x, features):
            return self._forward_composite(x, features, squeeze_dim)
        from ..b200 import ops
        plan = self.build_plan(x, features, with_dense=squeeze_dim)
        if plan is None:  # CUDA, but a layout the fused plan does not cover: per-table kernels, reference order
            return self._forward_composite(x, features, squeeze_dim)
        n_sparse = len(plan.fields) + len(plan.seqs)
        if squeeze_dim:
            if plan.tile_width == 0:
                raise ValueError("The input features can note be empty")
            tile, _, _ = ops.fused_tile(plan)
            return tile
        if n_sparse == 0:
            raise ValueError("If keep the original shape:[batch_size, num_features, embed_dim], expected %s in feature list, got %s" % ("SparseFeatures", features))
        dims = set(plan.sparse_dims)
        if len(dims) != 1:
            raise RuntimeError("Sizes of tensors must match except in dimension 1 (embed_dim differs across features: %s); use squeeze_dim=True" % sorted(dims))
        tile, _, _ = ops.fused_tile(plan)
        return tile.unflatten(1, (n_sparse, plan.sparse_dims[0]))


class LR(nn.Module):
    """``Linear(input_dim, 1)`` (+ optional sigmoid) — the first-order term (reference ``layers.py:164-189``)."""

    def __init__(self, input_dim, sigmoid=False):
        super().__init__()
        self.sigmoid = sigmoid
        self.fc = nn.Linear(input_dim, 1, bias=True)

    def forward(self, x):
        y = self.fc(x)
        return torch.sigmoid(y) if self.sigmoid else y

    def probability(self, x, extras=()):
        """``sigmoid(fc(x).squeeze(1) + sum(extras))`` — the tail of DCN / DCNv2 / WideDeep — as one launch on CUDA when
        ``config.fused_head_all`` is set (``rh_head_fwd``); the library ops otherwise."""
        if x.is_cuda:
            from ..b200 import config, ops
            if config.fused_head_all:
                p = ops.output_head(x, self.fc, extras, sigmoid=True)
                if p is not None:
                    return p
        y = self.fc(x).squeeze(1)
        for e in extras:
            y = y + e
        return torch.sigmoid(y)


_FUSED_ACTS = {nn.ReLU: "relu", Dice: "dice", nn.PReLU: "prelu", nn.Sigmoid: "sigmoid", nn.LeakyReLU: "leakyrelu"}


class MLP(nn.Module):
    """``[Linear -> BatchNorm1d -> activation -> Dropout] x len(dims)`` (+ ``Linear(.,1)``), reference ``layers.py:254-292``.

    The module list (``self.mlp``, an ``nn.Sequential``) is built exactly like the reference so the
    state_dict keys match (``mlp.<4i>.weight``, ``mlp.<4i+1>.running_mean`` ...).  On CUDA each hidden layer
    is a GEMM followed by ONE fused BatchNorm(+batch statistics)+activation+dropout pass
    (``rh_colstats`` + ``rh_bn_act_fwd``) instead of ~4 (ReLU) to ~14 (Dice) elementwise launches.
    """

    def __init__(self, input_dim, output_layer=True, dims=None, dropout=0, activation="relu"):
        super().__init__()
        if dims is None:
            dims = []
        layers = list()
        for i_dim in dims:
            layers.append(nn.Linear(input_dim, i_dim))
            layers.append(nn.BatchNorm1d(i_dim))
            layers.append(activation_layer(activation))
            layers.append(nn.Dropout(p=dropout))
            input_dim = i_dim
        if output_layer:
            layers.append(nn.Linear(input_dim, 1))
        self.mlp = nn.Sequential(*layers)

    @staticmethod
    def _fusable(bn, act, x):
        if x.dim() != 2 or type(act) not in _FUSED_ACTS:
            return False
        if not (bn.affine and bn.track_running_stats and bn.momentum is not None) or bn.num_features > 1024:
            return False
        if isinstance(act, nn.PReLU) and act.weight.numel() != 1:
            return False
        if isinstance(act, nn.LeakyReLU) and act.negative_slope != 0.01:
            return False
        return True

    def forward(self, x):
        if not x.is_cuda:
            return self.mlp(x)
        return self._forward_cuda(x, len(self.mlp))

    def forward_head(self, x, extras=(), sigmoid=True):
        """``f(self(x).squeeze(1) + sum(extras))`` with the ``Linear(., 1)`` output layer, the per-sample ``extras`` (``(B,)``
        tensors) and the sigmoid fused: in training mode the LAST hidden layer's BatchNorm + activation + dropout and the head
        are ONE launch (``rh_bn_act_fused_fwd`` in head mode), otherwise the head alone is (``rh_head_fwd``).  None when this
        tower has no such head or ``x`` is not on CUDA."""
        mods = list(self.mlp)
        if not x.is_cuda or not mods or not isinstance(mods[-1], nn.Linear) or mods[-1].out_features != 1:
            return None
        from ..b200 import ops
        n = len(mods) - 1
        if (n >= 4 and isinstance(mods[n - 4], nn.Linear) and isinstance(mods[n - 3], nn.BatchNorm1d) and isinstance(mods[n - 1], nn.Dropout) and mods[n - 3].training):
            h = self._forward_cuda(x, n - 4)
            lin, bn, act, drop = mods[n - 4], mods[n - 3], mods[n - 2], mods[n - 1]
            if self._fusable(bn, act, h):
                name = _FUSED_ACTS[type(act)]
                param = act.alpha if name == "dice" else (act.weight if name == "prelu" else None)
                y = ops.tower_layer_head(h, lin, bn, ops.ACT_CODES[name], param, getattr(act, "epsilon", 0.0), drop.p if drop.training else 0.0, mods[-1], extras, sigmoid)
                if y is not None:
                    return y
            h = self._forward_cuda(h, 4, start=n - 4)
        else:
            h = self._forward_cuda(x, n)
        y = ops.output_head(h, mods[-1], extras, sigmoid)
        if y is None:  # outside the kernel's shapes: finish with the library route
            y = mods[-1](h).squeeze(1)
            for e in extras:
                y = y + e
            y = torch.sigmoid(y) if sigmoid else y
        return y

    def _forward_cuda(self, x, n_mods, start=0):
        from ..b200 import ops
        mods = list(self.mlp)[start:start + n_mods]
        i = 0
        while i < len(mods):
            m = mods[i]
            if (isinstance(m, nn.Linear) and i + 3 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm1d) and isinstance(mods[i + 3], nn.Dropout) and self._fusable(mods[i + 1], mods[i + 2], x)):
                bn, act, drop = mods[i + 1], mods[i + 2], mods[i + 3]
                name = _FUSED_ACTS[type(act)]
                param = act.alpha if name == "dice" else (act.weight if name == "prelu" else None)
                x = ops.tower_layer(x, m, bn, ops.ACT_CODES[name], param, getattr(act, "epsilon", 0.0), drop.p if drop.training else 0.0, bn.training)
                i += 4
            else:
                x = m(x)
                i += 1
        return x


class FM(nn.Module):
    """Second-order FM term ``0.5 * sum_d[(sum_f x)^2 - sum_f x^2]`` (reference ``layers.py:295-319``)."""

    def __init__(self, reduce_sum=True):
        super().__init__()
        self.reduce_sum = reduce_sum

    def forward(self, x):
        if x.is_cuda and x.dim() == 3 and x.dtype == torch.float32:
            from ..b200 import ops
            return ops.fm(x, self.reduce_sum)
        square_of_sum = torch.sum(x, dim=1)**2
        sum_of_square = torch.sum(x**2, dim=1)
        ix = square_of_sum - sum_of_square
        if self.reduce_sum:
            ix = torch.sum(ix, dim=1, keepdim=True)
        return 0.5 * ix


class CrossNetwork(nn.Module):
    """DCN cross network: ``x_{l+1} = x0 * <w_l, x_l> + b_l + x_l`` (reference ``layers.py:390-420``).

    On CUDA all layers run in one launch with the row held in registers (``rh_cross_fwd``).
    """

    def __init__(self, input_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        self.w = torch.nn.ModuleList([torch.nn.Linear(input_dim, 1, bias=False) for _ in range(num_layers)])
        self.b = torch.nn.ParameterList([torch.nn.Parameter(torch.zeros((input_dim,))) for _ in range(num_layers)])

    def forward(self, x):
        if x.is_cuda and x.dim() == 2 and self.num_layers <= 16 and x.shape[1] <= 2048:
            from ..b200 import ops
            return ops.cross_network(x, [lin.weight for lin in self.w], list(self.b))
        x0 = x
        for i in range(self.num_layers):
            xw = self.w[i](x)
            x = x0 * xw + self.b[i] + x
        return x


class CrossNetV2(nn.Module):
    """DCN-v2 full-rank cross layers ``x0 * (W x) + b + x`` (reference ``layers.py:423-444``).
    On CUDA: one tensor-core GEMM (``rh_gemm_tf32x3``) + one fused cross step (``rh_crossmix_out_fwd``) per layer."""

    def __init__(self, input_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        self.w = torch.nn.ModuleList([torch.nn.Linear(input_dim, input_dim, bias=False) for _ in range(num_layers)])
        self.b = torch.nn.ParameterList([torch.nn.Parameter(torch.zeros((input_dim,))) for _ in range(num_layers)])

    def forward(self, x):
        if x.is_cuda and x.dim() == 2 and 0 < self.num_layers <= 8:
            from ..b200 import ops
            return ops.cross_net_v2(x, [lin.weight for lin in self.w], list(self.b))
        x0 = x
        for i in range(self.num_layers):
            x = x0 * self.w[i](x) + self.b[i] + x
        return x


class CrossNetMix(nn.Module):
    """DCN-v2 mixture of low-rank experts (reference ``layers.py:447-506``).

    Per layer and expert: ``x0 * (U tanh(C tanh(V^T x_l)) + bias)``, experts mixed by a softmax over
    ``Linear(width,1)`` gates (the gate modules are shared by all layers, reference ``:466``).
    On CUDA a layer is three tensor-core GEMMs over packed operands + three fused maps (``b200.ops._CrossMix``,
    ``csrc/rh_crossmix.cu``) instead of the reference's ~60 launches.
    """

    def __init__(self, input_dim, num_layers=2, low_rank=32, num_experts=4):
        super(CrossNetMix, self).__init__()
        self.num_layers = num_layers
        self.num_experts = num_experts
        mk = lambda *shape: nn.Parameter(nn.init.xavier_normal_(torch.empty(*shape)))
        # creation order U, V, C per kind matches the reference's RNG consumption (layers.py:455-461)
        self.u_list = torch.nn.ParameterList([mk(num_experts, input_dim, low_rank) for _ in range(self.num_layers)])
        self.v_list = torch.nn.ParameterList([mk(num_experts, input_dim, low_rank) for _ in range(self.num_layers)])
        self.c_list = torch.nn.ParameterList([mk(num_experts, low_rank, low_rank) for _ in range(self.num_layers)])
        self.gating = nn.ModuleList([nn.Linear(input_dim, 1, bias=False) for _ in range(self.num_experts)])
        self.bias = torch.nn.ParameterList([nn.Parameter(nn.init.zeros_(torch.empty(input_dim, 1))) for _ in range(self.num_layers)])

    def _forward_batched(self, x):
        """The packed formulation the CUDA kernels implement, written with stock torch ops (the CPU-side statement of the algebra,
        checked against the reference-order loop in tests/test_cpu_api.py):

            [a | g] = x_l [V_1 .. V_E | Wg^T]            one (B, W) x (W, E r + E) product: every expert's projection AND the gates
            t2_e    = tanh(C_e tanh(a_e)),  s = softmax(g)
            x_{l+1} = x_0 * ([s_1 t2_1 .. s_E t2_E] [U_1 .. U_E]^T + b) + x_l

        — the gate-weighted sum over experts moves inside the last product's K dimension (sum_e s_e = 1 keeps the bias whole)."""
        n, width = x.shape
        E = self.num_experts
        gates = torch.cat([g.weight for g in self.gating], dim=0)  # (E, W)
        x_0, x_l = x, x
        for i in range(self.num_layers):
            U, V, C = self.u_list[i], self.v_list[i], self.c_list[i]  # (E, W, r), (E, W, r), (E, r, r)
            r = V.shape[2]
            first = torch.cat([V.permute(1, 0, 2).reshape(width, E * r), gates.t()], dim=1)  # (W, E r + E)
            ag = x_l @ first
            t1 = torch.tanh(ag[:, :E * r]).view(n, E, r)
            t2 = torch.tanh(torch.einsum("ber,esr->bes", t1, C))  # C_e @ t1_e per expert
            z = (t2 * torch.softmax(ag[:, E * r:], dim=1).unsqueeze(2)).reshape(n, E * r)
            x_l = x_0 * (z @ U.permute(0, 2, 1).reshape(E * r, width) + self.bias[i].view(1, width)) + x_l
        return x_l.squeeze()  # the reference's squeeze() quirk at B == 1 (layers.py:505), kept on this route too

    def forward(self, x):
        if x.is_cuda and x.dim() == 2 and 0 < self.num_layers <= 8 and self.num_experts <= 8:
            from ..b200 import ops
            out = ops.cross_net_mix(x, list(self.u_list), list(self.v_list), list(self.c_list), [g.weight for g in self.gating], list(self.bias))
            return out.squeeze()  # NOTE reference quirk kept: squeeze() drops the batch axis when B == 1 (layers.py:505)
        x_0 = x.unsqueeze(2)  # (B, width, 1)
        x_l = x_0
        for i in range(self.num_layers):
            expert_out, gate_scores = [], []
            flat = x_l.squeeze(2)
            for e in range(self.num_experts):
                gate_scores.append(self.gating[e](flat))
                v_x = torch.tanh(torch.matmul(self.v_list[i][e].t(), x_l))  # (B, r, 1)
                v_x = torch.tanh(torch.matmul(self.c_list[i][e], v_x))
                uv_x = torch.matmul(self.u_list[i][e], v_x)  # (B, width, 1)
                expert_out.append((x_0 * (uv_x + self.bias[i])).squeeze(2))
            expert_out = torch.stack(expert_out, 2)  # (B, width, E)
            gate_scores = torch.stack(gate_scores, 1)  # (B, E, 1)
            x_l = torch.matmul(expert_out, gate_scores.softmax(1)) + x_l
        return x_l.squeeze()  # NOTE reference quirk kept: squeeze() drops the batch axis when B == 1 (layers.py:505)
