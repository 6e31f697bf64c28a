"""TEST / BENCH INFRASTRUCTURE ONLY — CPU port of the reference's DeepFM training step, built from stock torch modules.

Purpose: the ``cpu_baseline`` / ``--impl reference`` leg of ``bench.py`` on the GPU box, where ``/root/reference`` does
not exist.  It restates, with plain ``torch.nn`` modules and no code from this repo's package, what the reference runs:

* tables: one ``nn.Embedding(vocab, dim)`` per sparse field, default ``sparse=False`` so every lookup's backward
  materialises a dense ``(vocab, dim)`` gradient (reference ``basic/initializers.py:16-21``);
* ``EmbeddingLayer.forward``: per-field lookup -> ``unsqueeze(1)`` -> ``cat`` (reference ``basic/layers.py:77-127``);
* DeepFM: ``sigmoid(LR(flatten(e_fm)) + FM(e_fm) + MLP(e_deep))`` with the table looked up once per use
  (reference ``models/ranking/deepfm.py:34-43``; ``FM`` ``layers.py:313-319``; ``MLP`` ``layers.py:276-292``);
* step: forward, ``BCELoss``, ``zero_grad``, ``backward``, ``Adam(lr=1e-3, weight_decay=1e-5).step()``
  (reference ``trainers/ctr_trainer.py:60-68,87-99``).

Validated against the live reference by ``tests/test_oracle.py::test_ref_port_matches_live_reference`` (bitwise).
Nothing under ``torch-rechub_b200/`` may import this file.
"""
import time

import torch
import torch.nn as nn


class PortDeepFM(nn.Module):

    def __init__(self, n_dense, vocab_sizes, dim, mlp_dims=(256, 128), dropout=0.2, deep_includes_sparse=True, init_std=1e-4):
        super().__init__()
        self.n_dense = n_dense
        self.deep_includes_sparse = deep_includes_sparse
        n_sparse = len(vocab_sizes)
        self.linear = nn.Linear(n_sparse * dim, 1)  # LR (layers.py:183)
        self.tables = nn.ModuleList()
        for v in vocab_sizes:
            emb = nn.Embedding(v, dim)
            nn.init.normal_(emb.weight, 0.0, init_std)
            self.tables.append(emb)
        in_dim = n_dense + (n_sparse * dim if deep_includes_sparse else 0)
        layers = []
        for d in mlp_dims:
            layers += [nn.Linear(in_dim, d), nn.BatchNorm1d(d), nn.ReLU(inplace=True), nn.Dropout(p=dropout)]
            in_dim = d
        layers.append(nn.Linear(in_dim, 1))
        self.mlp = nn.Sequential(*layers)

    def _lookup(self, ids):  # ids: list of (B,) int64
        return torch.cat([t(i.long()).unsqueeze(1) for t, i in zip(self.tables, ids)], dim=1)  # (B, F, D)

    def forward(self, dense, ids):
        # dense: list of (B,) float columns; ids: list of (B,) int64 columns
        dense_values = torch.cat([d.float().unsqueeze(1) for d in dense], dim=1)
        if self.deep_includes_sparse:  # layers.py:120 — sparse block first, dense appended; a SECOND lookup of every table
            input_deep = torch.cat((self._lookup(ids).flatten(start_dim=1), dense_values), dim=1)
        else:
            input_deep = dense_values
        input_fm = self._lookup(ids)
        y_linear = self.linear(input_fm.flatten(start_dim=1))
        square_of_sum = torch.sum(input_fm, dim=1)**2
        sum_of_square = torch.sum(input_fm**2, dim=1)
        y_fm = 0.5 * torch.sum(square_of_sum - sum_of_square, dim=1, keepdim=True)
        y = y_linear + y_fm + self.mlp(input_deep)
        return torch.sigmoid(y.squeeze(1))


def time_train_steps(model, batches, steps, warmup, with_optimizer=True, lr=1e-3, weight_decay=1e-5):
    """Run ``warmup + steps`` CTRTrainer-style steps on CPU; returns seconds per timed step (list)."""
    opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=weight_decay) if with_optimizer else None
    crit = nn.BCELoss()
    model.train()
    times = []
    for i in range(warmup + steps):
        dense, ids, y = batches[i % len(batches)]
        t0 = time.perf_counter()
        y_pred = model(dense, ids)
        loss = crit(y_pred, y)
        model.zero_grad()
        loss.backward()
        if opt is not None:
            opt.step()
        loss_value = loss.item()  # ctr_trainer.py:100
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, loss_value
