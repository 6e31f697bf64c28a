"""Diagnostic: the tower (MLP 429-256-128 + Linear head) at batch 4096 on the engine vs float64 on the CPU, per-parameter gradient
error with the location of the worst element, under the engine switches given in the environment."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "torch-rechub_b200"))
import copy

import torch

from torch_rechub.basic.layers import MLP

torch.manual_seed(0)
B, K = 4096, 429
m = MLP(K, output_layer=True, dims=[256, 128], dropout=0.0, activation="relu")
x = torch.randn(B, K) * 0.05
x[:, 416:] = torch.rand(B, 13)
e0 = torch.randn(B) * 0.1
y = torch.randint(0, 2, (B,)).float()
md = copy.deepcopy(m).double()
xd = x.double().requires_grad_(True)
pd = torch.sigmoid(md(xd).squeeze(1) + e0.double())
torch.nn.BCELoss()(pd, y.double()).backward()
mg = copy.deepcopy(m).cuda().train()
buf = torch.empty(B, 432, device="cuda")
buf[:, :K] = x.cuda()
xg = buf[:, :K].requires_grad_(True)
head = os.environ.get("DIAG_HEAD", "1") == "1"
if head:
    pg = mg.forward_head(xg, (e0.cuda(),), sigmoid=True)
else:
    pg = torch.sigmoid(mg(xg).squeeze(1) + e0.cuda())
torch.nn.BCELoss()(pg, y.cuda()).backward()
print("switches:", {k: v for k, v in os.environ.items() if k.startswith("RECHUB_")}, "head", head)
print("prob err", float((pg.detach().cpu().double() - pd.detach()).abs().max()))
gx = xg.grad.cpu().double()
print("d_x err %.3e scale %.3e" % (float((gx - xd.grad).abs().max()), float(xd.grad.abs().max())))
for (n, p), q in zip(mg.named_parameters(), md.parameters()):
    d = (p.grad.cpu().double() - q.grad).abs()
    idx = int(d.argmax())
    loc = tuple(int(v) for v in torch.unravel_index(torch.tensor(idx), d.shape)) if d.dim() > 0 else ()
    nbad = int((d > 1e-3 * q.grad.abs().max()).sum())
    print("%-16s err %.3e scale %.3e rel %.2e worst at %s  n_bad(>1e-3) %d of %d" % (n, float(d.max()), float(q.grad.abs().max()), float(d.max() / q.grad.abs().max().clamp_min(1e-30)), loc, nbad, d.numel()))
    if n == "mlp.0.weight" and nbad > 0:
        bad = (d > 1e-3 * q.grad.abs().max()).nonzero()
        print("   bad rows:", sorted(set(int(r) for r in bad[:, 0]))[:20], " bad cols:", sorted(set(int(c) for c in bad[:, 1]))[:40])
