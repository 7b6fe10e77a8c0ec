"""Small driver for ncu: a few fused forward+backward steps of the BASELINE config-2 workload.
    python tools/prof_driver.py [realistic|dense] [steps] [flags]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pertrenderer_b200 import ops, synthetic_fragments  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "realistic"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
N, HW, K, S = 8, 256, 50, 64
dev = "cuda:0"
if kind == "rasterised":
    import types
    import bench
    fr, col = bench.rasterised_fragments(types.SimpleNamespace(views=N, image_size=HW, faces_per_pixel=K, nb_samples=S), torch.device(dev))
else:
    fr, col = synthetic_fragments(N, HW, HW, K, kind=kind, sigma=1e-3, seed=0, device=dev)
G = torch.randn((N, HW, HW, 4), device=dev, generator=torch.Generator(device=dev).manual_seed(1))
for i in range(steps):
    pr = ops.ShadeProblem(pix_to_face=fr.pix_to_face, zbuf=fr.zbuf, dists=fr.dists, colors=col, znear=1.0, zfar=100.0,
                          background=(1.0, 1.0, 1.0), sigma=1e-3, gamma=1e-2, alpha=1.0, eps=1e-10, S_rast=S, S_agg=S,
                          seed_rast=100 + i, seed_agg=200 + i, flags=flags)
    image, saved = ops.shade_forward(pr)
    out = ops.shade_backward(pr, saved, G)
torch.cuda.synchronize()
print("ok", float(image.sum()), float(out[0].sum()))
