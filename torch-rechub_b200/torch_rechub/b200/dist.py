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
    """out[s] = what rank s sent me; equal splits.  x: (world, rows, width) -> (world, rows, width).  Backward = the same exchange."""

    @staticmethod
    def forward(ctx, x, group):
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
        return out, None


class ShardedFront(object):
    """Field-sharded replacement for ``EmbeddingLayer.forward`` (installed by :func:`attach`)."""

    def __init__(self, layer, group, device):
        self.layer = layer
        self.group = group
        self.device = device
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.owner = field_owners(list(layer.embed_dict.keys()), self.world)
        self._plans = {}

    def owner_of(self, fea):
        return self.owner[fea.name if fea.shared_with is None else fea.shared_with]

    def _plan(self, features):
        key = tuple(id(f) for f in features)
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
            by_owner = [[f for f in sparse if self.owner_of(f) == r] for r in range(self.world)]
            fmax = max(len(b) for b in by_owner) if sparse else 0
            # column permutation: position of each feature's block in the owner-grouped (padded) receive buffer
            dim = sparse[0].embed_dim if sparse else 0
            cols = []
            for f in sparse:
                r = self.owner_of(f)
                k = by_owner[r].index(f)
                base = (r * fmax + k) * dim
                cols.extend(range(base, base + dim))
            p = {"sparse": sparse, "dense": dense, "by_owner": by_owner, "fmax": fmax, "dim": dim, "cols": torch.tensor(cols, dtype=torch.long, device=self.device)}
            self._plans[key] = p
        return p

    def forward(self, x, features, squeeze_dim):
        p = self._plan(features)
        sparse, dense, W, fmax, dim = p["sparse"], p["dense"], self.world, p["fmax"], p["dim"]
        parts = []
        if sparse:
            b = x[sparse[0].name].shape[0]
            # 1) ids all-to-all: chunk r = my samples' ids of the fields rank r owns, padded to fmax columns
            send = torch.zeros((W, fmax, b), dtype=torch.long, device=self.device)
            for r in range(W):
                for k, f in enumerate(p["by_owner"][r]):
                    send[r, k] = x[f.name].long()
            recv = torch.empty_like(send)
            dist.all_to_all_single(recv, send, group=self.group)  # recv[s, k, i] = id of sample i of rank s for my k-th field
            mine = p["by_owner"][self.rank]
            # 2) owner-local fused gather over the global batch (W*b samples) of my fields only
            if mine:
                xg = {f.name: recv[:, k, :].reshape(W * b) for k, f in enumerate(mine)}
                rows = self.layer._forward_local(xg, mine, squeeze_dim=True)  # (W*b, len(mine)*dim)
                if len(mine) < fmax:
                    rows = torch.nn.functional.pad(rows, (0, (fmax - len(mine)) * dim))
            else:
                rows = torch.zeros((W * b, fmax * dim), dtype=torch.float32, device=self.device)
            # 3) rows all-to-all back to the samples' ranks (backward: tile gradients travel the other way)
            got = _AllToAllRows.apply(rows.view(W, b, fmax * dim), self.group)  # got[r] = rank r's fields for MY samples
            grouped = got.permute(1, 0, 2).reshape(b, W * fmax * dim)
            parts.append(grouped.index_select(1, p["cols"]))  # the model's column order
        if squeeze_dim:
            for f in dense:
                v = x[f.name].float()
                parts.append(v if v.dim() > 1 else v.unsqueeze(1))
            if not parts:
                raise ValueError("The input features can note be empty")
            return torch.cat(parts, dim=1) if len(parts) > 1 else parts[0]
        if not sparse:
            raise ValueError("If keep the original shape:[batch_size, num_features, embed_dim], expected %s in feature list, got %s" % ("SparseFeatures", features))
        return parts[0].unflatten(1, (len(sparse), dim))


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
        for mod in model.modules():
            if isinstance(mod, EmbeddingLayer):
                front = ShardedFront(mod, self.group, self.device)
                mod._dist = front
                self.fronts.append(front)
                for name, tbl in mod.embed_dict.items():
                    if front.owner[name] == self.rank:
                        self.owned.append(tbl.weight)
                    else:
                        self.foreign.append(tbl.weight)
                        tbl._dist_shape = tuple(tbl.weight.shape)
                        tbl.weight.data = torch.empty((0, tbl.weight.shape[1]), dtype=tbl.weight.dtype, device=tbl.weight.device)  # free it
                        tbl.weight.requires_grad_(False)
        skip = {id(p) for p in self.owned} | {id(p) for p in self.foreign}
        self.dense_params = [p for p in model.parameters() if id(p) not in skip and p.requires_grad]
        # replicated parameters start identical on every rank
        for p in self.dense_params:
            dist.broadcast(p.data, src=0, group=self.group)
        for b in model.buffers():
            if b.dtype.is_floating_point:
                dist.broadcast(b.data, src=0, group=self.group)
        n = sum(p.numel() for p in self.dense_params)
        self.flat = torch.zeros(n + 1, dtype=torch.float32, device=self.device)  # [+1]: the loss rides along
        off = 0
        for p in self.dense_params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    # -- one training step -------------------------------------------------------------------------------------
    def train_step(self, trainer, x_dict, y):
        loss = trainer._loss(x_dict, y)
        self.flat.zero_()  # dense grads are views of this bucket
        for p in self.owned:
            p.grad = None
        (loss / self.world).backward()
        self.flat[-1] = loss.detach() / self.world
        dist.all_reduce(self.flat, group=self.group)  # SUM over ranks of (local grad / world) = gradient of the global-batch mean
        trainer.optimizer.step()
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
