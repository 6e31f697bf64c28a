"""DeepFM (mirror of reference ``torch_rechub/models/ranking/deepfm.py:14-43``).

``sigmoid(LR(flatten(e_fm)) + FM(e_fm) + MLP(e_deep))``.

On CUDA the whole front end is ONE launch: every table is gathered once (the reference gathers a table
twice when a feature is in both ``deep_features`` and ``fm_features``, deepfm.py:35,37), FM and LR are
reduced in registers, and the MLP's input tile is written in the same pass (``rh_fields_fwd``).
"""
import torch

from ...basic.features import SparseFeature
from ...basic.layers import FM, LR, MLP, EmbeddingLayer


class DeepFM(torch.nn.Module):
    """Deep Factorization Machine.

    Args:
        deep_features (list): features feeding the MLP tower.
        fm_features (list): features feeding the FM + linear terms.
        mlp_params (dict): ``{"dims": list, "activation": str, "dropout": float, "output_layer": bool}``.
    """

    def __init__(self, deep_features, fm_features, mlp_params):
        super(DeepFM, self).__init__()
        self.deep_features = deep_features
        self.fm_features = fm_features
        self.deep_dims = sum([fea.embed_dim for fea in deep_features])
        self.fm_dims = sum([fea.embed_dim for fea in fm_features])
        self.linear = LR(self.fm_dims)  # 1st-order term
        self.fm = FM(reduce_sum=True)  # 2nd-order term
        self.embedding = EmbeddingLayer(deep_features + fm_features)
        self.mlp = MLP(self.deep_dims, **mlp_params)

    def _fused_plan(self, x):
        """One TilePlan covering the deep tile and the FM/LR reduction, or None (-> layer-by-layer route)."""
        fm = self.fm_features
        if not fm or len(fm) > 64 or not all(isinstance(f, SparseFeature) for f in fm):
            return None
        dim = fm[0].embed_dim
        if dim % 4 != 0 or dim > 128 or any(f.embed_dim != dim for f in fm):
            return None
        if len({f.name for f in fm}) != len(fm):
            return None
        from ...b200 import ops
        plan = self.embedding.build_plan(x, self.deep_features, with_dense=True)
        if plan is None:
            return None
        for slot, fea in enumerate(fm):
            ids = x[fea.name]
            if ids.dim() != 1:
                return None
            tbl = self.embedding.table_of(fea)
            if tbl.weight.shape[1] != dim:
                return None
            ref = plan.by_name.get(fea.name)
            if ref is not None and ref.weight is tbl.weight:
                ref.fm_slot = slot  # gathered once, used twice
            else:
                plan.fields.append(ops.FieldRef(tbl.weight, ops._as_ids(ids), tbl.padding_idx, -1, slot))
        plan.n_fm, plan.fm_dim = len(fm), dim
        plan.want_fm = plan.want_lr = True
        return plan

    def _fm_from_deep(self, input_deep):
        """(B, n_fm, D) view/gather of the deep tile's columns when every FM feature is also a deep feature
        (sharded runs: ONE all-to-all exchange serves both uses)."""
        from ...basic.features import DenseFeature
        offsets, col = {}, 0
        for fea in self.deep_features:
            if not isinstance(fea, DenseFeature):
                offsets.setdefault(fea.name, (col, fea.embed_dim))
                col += fea.embed_dim
        dim = self.fm_features[0].embed_dim
        starts = []
        for fea in self.fm_features:
            if fea.name not in offsets or offsets[fea.name][1] != dim or fea.embed_dim != dim:
                return None
            starts.append(offsets[fea.name][0])
        n = len(starts)
        if all(starts[i] == starts[0] + i * dim for i in range(n)):
            return input_deep[:, starts[0]:starts[0] + n * dim].unflatten(1, (n, dim))
        idx = torch.tensor([c for s0 in starts for c in range(s0, s0 + dim)], device=input_deep.device)
        return input_deep.index_select(1, idx).unflatten(1, (n, dim))

    def _tail(self, input_deep, y_fm, y_linear):
        """sigmoid(y_linear + y_fm + mlp(input_deep)) with the (B,) per-sample terms of the fused front end (deepfm.py:41-43)."""
        p = self.mlp.forward_head(input_deep, (y_linear, y_fm), sigmoid=True)
        if p is not None:
            return p
        y = y_linear.unsqueeze(1) + y_fm.unsqueeze(1) + self.mlp(input_deep)
        return torch.sigmoid(y.squeeze(1))

    def forward(self, x):
        front = self.embedding._dist
        if front is not None:
            from ...basic.features import DenseFeature
            deep_names = {f.name for f in self.deep_features if not isinstance(f, DenseFeature)}
            dim = self.fm_features[0].embed_dim if self.fm_features else 0
            fusable = (front.device.type == "cuda" and self.fm_features and all(f.name in deep_names and f.embed_dim == dim for f in self.fm_features) and dim % 4 == 0 and
                       len({f.name for f in self.fm_features}) == len(self.fm_features) <= 64)
            if fusable:  # ONE exchange; the receiving kernel unpacks the rows and reduces FM + LR in the same pass
                input_deep, y_fm, y_linear = front.run(x, self.deep_features, fm_features=self.fm_features, lr=(self.linear.fc.weight, self.linear.fc.bias))
                return self._tail(input_deep, y_fm, y_linear)
            input_deep = self.embedding(x, self.deep_features, squeeze_dim=True)
            input_fm = self._fm_from_deep(input_deep) if self.fm_features else None
            if input_fm is None:
                input_fm = self.embedding(x, self.fm_features, squeeze_dim=False)
            y = self.linear(input_fm.flatten(start_dim=1)) + self.fm(input_fm) + self.mlp(input_deep)
            return torch.sigmoid(y.squeeze(1))
        plan = self._fused_plan(x) if self.embedding._on_cuda(x, self.deep_features + self.fm_features) else None
        if plan is not None:
            from ...b200 import ops
            input_deep, y_fm, y_linear = ops.fused_tile(plan, self.linear.fc.weight, self.linear.fc.bias)
            return self._tail(input_deep, y_fm, y_linear)
        input_deep = self.embedding(x, self.deep_features, squeeze_dim=True)  # (B, deep_dims)
        input_fm = self.embedding(x, self.fm_features, squeeze_dim=False)  # (B, n_fm, D)
        y_linear = self.linear(input_fm.flatten(start_dim=1))
        y_fm = self.fm(input_fm)
        y_deep = self.mlp(input_deep)
        y = y_linear + y_fm + y_deep
        return torch.sigmoid(y.squeeze(1))
