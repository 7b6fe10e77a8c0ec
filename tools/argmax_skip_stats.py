"""How many (logit, sample-quad) evaluations of the perturbed argmax could be skipped exactly: with the live logits of a
pixel sorted by descending zeta, a lane can stop at the first logit with zeta_l + gamma*Umax <= min over its 4 samples of
the running best (bounded noise: it and every later logit cannot win any of the 4).
    python tools/argmax_skip_stats.py [rasterised|dense|realistic]"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pertrenderer_b200 import ops, synthetic_fragments  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "rasterised"
N, HW, K, S = 8, 256, 50, 64
sigma, gamma, UMAX = 1e-3, 1e-2, 5.66
dev = torch.device("cuda:0")
if kind == "rasterised":
    fr, col = bench.rasterised_fragments(types.SimpleNamespace(views=N, image_size=HW, faces_per_pixel=K, nb_samples=S), dev)
else:
    fr, col = synthetic_fragments(N, HW, HW, K, kind=kind, sigma=sigma, seed=0, device=dev)
pr = ops.ShadeProblem(pix_to_face=fr.pix_to_face, zbuf=fr.zbuf, dists=fr.dists, colors=col, znear=1.0, zfar=100.0,
                      background=(1.0, 1.0, 1.0), sigma=sigma, gamma=gamma, alpha=1.0, eps=1e-10, S_rast=S, S_agg=S,
                      seed_rast=1, seed_agg=2)
image, saved = ops.shade_forward(pr)
torch.cuda.synchronize()
P = N * HW * HW
valid = (fr.pix_to_face >= 0).reshape(P, K)
cov = valid.any(-1).nonzero()[:, 0]
sel = cov[torch.randperm(cov.numel(), generator=torch.Generator().manual_seed(0))[:8192].to(dev)].sort().values
p2f = fr.pix_to_face.reshape(P, K)[sel].cpu()[None, None]
zb = fr.zbuf.reshape(P, K)[sel].cpu()[None, None]
cnt = ((saved.counts.to(torch.int32) & 0xFFFF).reshape(P, K)[sel].cpu() * valid[sel].cpu())[None, None]
# logits of smoothagg.py:198-202 from the kernel's own hit counts (alpha = 1, znear = 1, zfar = 100, eps = 1e-10)
m = (p2f >= 0)[0, 0]
prob = (cnt[0, 0].float() / S) * m
zi = (100.0 - zb[0, 0]) / 99.0 * m
zmx = zi.max(-1, keepdim=True).values.clamp(min=1e-10)
zeta = torch.cat((gamma * torch.log(prob) + zi - zmx, 1e-10 - zmx), dim=-1)  # (n, K+1); log 0 = -inf: can never win
zmax = zeta.max(-1, keepdim=True).values
live = torch.isfinite(zeta) & (zeta >= zmax - 2 * gamma * UMAX)
nlive = live.sum(-1)
print(f"{kind}: pixels {zeta.shape[0]}, live logits per pixel mean {nlive.float().mean():.2f}; multi-logit pixels {(nlive > 1).float().mean():.3f}")
gap = ((zmax - zeta) / gamma)[live]
print("  (zmax - zeta)/gamma of live logits, deciles:", [round(float(q), 2) for q in torch.quantile(gap, torch.linspace(0.1, 1.0, 10))])
# simulate
g = torch.Generator().manual_seed(1)
zs = torch.where(live, zeta, torch.full_like(zeta, -float("inf")))
zsort, _ = zs.sort(-1, descending=True)
n, K1 = zsort.shape
Q = S // 4
V = torch.randn((n, Q, 4, K1), generator=g).clamp(-5.647, 5.647)
v = zsort[:, None, None, :] + gamma * V  # -inf where dead
run = torch.cummax(v, dim=-1).values  # best after logits 0..l
minbest = run.min(2).values  # (n, Q, K1): min over the 4 samples
bound = zsort + gamma * UMAX
# logit l is evaluated iff l == 0 or bound[l] > minbest after l-1
need = torch.ones((n, Q, K1), dtype=torch.bool)
need[:, :, 1:] = bound[:, None, 1:] > minbest[:, :, :-1]
need &= torch.isfinite(zsort)[:, None, :]
exit_at = need.sum(-1)  # sorted => prefix
multi = nlive > 1
tot_full = (nlive[multi] * Q).sum().item()
tot_lane = exit_at[multi].sum().item()
print(f"  per-lane evaluations with early exit: {tot_lane / tot_full:.3f} of the full loop")
# warp model: pixels ordered by nlive (as the kernel orders a tile's pixels), 8 pixels x 4 lanes, lane q handles quads q, q+4, ...
order = torch.argsort(nlive[multi], descending=True)
ex = exit_at[multi][order]
nl = nlive[multi][order]
m = (ex.shape[0] // 8) * 8
ex = ex[:m].reshape(-1, 8, Q // 4, 4)  # (warp, pixel, iteration, lane)
nl = nl[:m].reshape(-1, 8)
warp_cost = ex.permute(0, 2, 1, 3).reshape(ex.shape[0], Q // 4, 32).max(-1).values.sum().item()
warp_full = (nl.max(-1).values * (Q // 4)).sum().item()
print(f"  warp-level (8 pixels x 4 lanes, max over lanes): {warp_cost / warp_full:.3f} of the full loop")
