"""Statistics of a fragment set that decide the cost of the perturbed shader: how many entries can flip a
coverage sample, how many logits can win, how many pixels are active.
    python tools/frag_stats.py [realistic|dense|rasterised]"""
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
sigma, gamma = 1e-3, 1e-2
dev = torch.device("cuda:0")
args = types.SimpleNamespace(views=N, image_size=HW, faces_per_pixel=K, nb_samples=S)
if kind == "rasterised":
    fr, col = bench.rasterised_fragments(args, dev)
else:
    fr, col = synthetic_fragments(N, HW, HW, K, kind=kind, sigma=sigma, seed=0, device=dev)
valid = fr.pix_to_face >= 0
cov = valid.any(-1)
print(f"{kind}: coverage {cov.float().mean():.3f} valid/covered {valid.sum(-1)[cov].float().mean():.2f}")
t = (fr.dists.abs() / sigma)[valid]
edges = [0, 0.5, 1, 1.5, 2, 3, 4, 5.66, 1e30]
h = torch.histc(t.clamp(max=20), bins=200, min=0, max=20)
for lo, hi in zip(edges[:-1], edges[1:]):
    print(f"  |x|/sigma in [{lo},{hi}): {((t >= lo) & (t < hi)).float().mean():.4f}")
print("  inside (x>0):", ((-fr.dists)[valid] > 0).float().mean().item())
pr = ops.ShadeProblem(pix_to_face=fr.pix_to_face, zbuf=fr.zbuf, dists=fr.dists, colors=col, znear=1.0, zfar=100.0,
                      background=(1.0, 1.0, 1.0), sigma=sigma, gamma=gamma, alpha=1.0, eps=1e-10, S_rast=S, S_agg=S,
                      seed_rast=1, seed_agg=2)
image, saved = ops.shade_forward(pr, want_hist=True, phases=0x70)
cnt = (saved.counts.to(torch.int32) & 0xFFFF) * valid
live = (cnt > 0).sum(-1)
print(f"  logits with P>0 per covered pixel: mean {live[cov].float().mean():.2f} max {live.max().item()}")
print("  hist of P>0 logits per covered pixel:", torch.bincount(live[cov].clamp(max=50), minlength=51)[:40].tolist())
st = saved.pixstate.to(torch.int32) & 0xFFFF
act = (st & 0x8000) != 0
print(f"  active pixels: {act.float().mean():.3f} of all, {act[cov].float().mean():.3f} of covered")
hist = saved.hist
nw = (hist > 0).sum(-1)
print(f"  distinct winners per active pixel: mean {nw[act].float().mean():.2f}")
print("  hist of distinct winners (active):", torch.bincount(nw[act], minlength=20)[:30].tolist())
frac_nz = 1.0 - hist.gather(-1, (st & 0x7FFF).long()[..., None])[..., 0].float() / S
print(f"  samples not won by a0 (active pixels): {frac_nz[act].mean():.3f}")
