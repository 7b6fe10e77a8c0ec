"""Diagnostic: variance-ratio spread of the gradients between noise modes, against the null (same mode, other seeds)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_util import problem_from_case, run_cuda, synthetic_case
from pertrenderer_b200 import _cabi
N, H, W, K, S = 1, 8, 8, 12, 16
g = synthetic_case(N, H, W, K, S, S, kind="dense", seed=5)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
def runs(flags, off):
    acc = {k: [] for k in ("image", "grad_dists", "grad_zbuf", "counts", "rsum")}
    for r in range(reps):
        out = run_cuda(problem_from_case(g, explicit=False, seed_rast=31 * r + 7 + off, seed_agg=17 * r + 3 + 2 * off, flags=flags), g["grad_image"])
        for k in acc: acc[k].append(out[k].double())
    return {k: torch.stack(v) for k, v in acc.items()}
A = runs(0, 0); B = runs(_cabi.F_PER_SAMPLE_NOISE, 100000); C = runs(_cabi.F_PER_SAMPLE_NOISE, 300000); D = runs(0, 500000)
def cmp(x, y, name):
    for k in x:
        a, b = x[k], y[k]
        va, vb = a.var(0), b.var(0)
        ma, mb = a.mean(0), b.mean(0)
        m4a, m4b = ((a - ma) ** 4).mean(0), ((b - mb) ** 4).mean(0)
        sev = ((m4a - va * va).clamp(min=0) / reps + (m4b - vb * vb).clamp(min=0) / reps).sqrt()
        zv = (va - vb).abs() / (sev + 1e-3 * sev.max() + 1e-30)
        zm = (ma - mb).abs() / ((va / reps + vb / reps).sqrt() + 1e-3 * (va / reps + vb / reps).sqrt().max() + 1e-30)
        big = vb > 0.05 * vb.max()
        ratio = va[big] / vb[big]
        print(f"{name:12s} {k:10s} zmean {zm.max():.2f} zvar {zv.max():.2f} ratio [{ratio.min():.3f}, {ratio.max():.3f}] n={int(big.sum())}")
cmp(A, B, "cmp-vs-ps"); cmp(C, B, "ps-vs-ps"); cmp(A, D, "cmp-vs-cmp")
# where is the worst ratio?
a, b = A["grad_dists"], B["grad_dists"]
va, vb = a.var(0), b.var(0)
big = vb > 0.05 * vb.max()
r = torch.where(big, va / vb, torch.ones_like(va))
idx = (r.log().abs()).flatten().topk(5).indices
x = (-g["dists"]).flatten() / 1e-3
for i in idx.tolist():
    print("entry", i, "x/sigma %.3f" % x[i].item(), "ratio %.3f" % r.flatten()[i].item(), "va %.3e vb %.3e" % (va.flatten()[i].item(), vb.flatten()[i].item()),
          "max|a| %.3e max|b| %.3e" % (a.reshape(reps, -1)[:, i].abs().max().item(), b.reshape(reps, -1)[:, i].abs().max().item()),
          "nz a %d b %d" % ((a.reshape(reps, -1)[:, i] != 0).sum().item(), (b.reshape(reps, -1)[:, i] != 0).sum().item()))
