"""Device-timed fused forward / backward at BASELINE config 2 (or --cfg N,HW,K,S) on the named fragment sets.
    python tools/time_shade.py [realistic dense rasterised] [--flags F] [--steps n] [--cfg 8,256,50,64]"""
import argparse
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pertrenderer_b200 import ops, synthetic_fragments  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("kinds", nargs="*", default=["realistic", "dense", "rasterised"])
ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--cfg", default="8,256,50,64")
ap.add_argument("--tag", default="")
ap.add_argument("--lib", default=None, help="alternate libpertshade.so (A/B runs of two builds in one GPU call)")
a = ap.parse_args()
if a.lib:
    from pertrenderer_b200 import _cabi
    _cabi.LIB_PATH = os.path.abspath(a.lib)
N, HW, K, S = (int(v) for v in a.cfg.split(","))
dev = torch.device("cuda:0")
peak = bench.peaks()[0]
for kind in a.kinds:
    if kind == "rasterised":
        fr, col = bench.rasterised_fragments(types.SimpleNamespace(views=N, image_size=HW, faces_per_pixel=K, nb_samples=S), dev)
    else:
        fr, col = synthetic_fragments(N, HW, HW, K, kind=kind, sigma=1e-3, seed=0, device=dev)
    G = torch.randn((N, HW, HW, 4), device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(a.steps)]
    for i in range(-3, a.steps):
        pr = ops.ShadeProblem(pix_to_face=fr.pix_to_face, zbuf=fr.zbuf, dists=fr.dists, colors=col, znear=1.0, zfar=100.0,
                              background=(1.0, 1.0, 1.0), sigma=1e-3, gamma=1e-2, alpha=1.0, eps=1e-10, S_rast=S, S_agg=S,
                              seed_rast=100 + i, seed_agg=200 + i, flags=a.flags)
        if i >= 0:
            ev[i][0].record()
        image, saved = ops.shade_forward(pr)
        if i >= 0:
            ev[i][1].record()
        out = ops.shade_backward(pr, saved, G)
        if i >= 0:
            ev[i][2].record()
    torch.cuda.synchronize()
    f = sorted(e[0].elapsed_time(e[1]) for e in ev)[len(ev) // 2]
    b = sorted(e[1].elapsed_time(e[2]) for e in ev)[len(ev) // 2]
    P = N * HW * HW
    fb, bb = bench.alg_bytes(P, K)
    print(f"{a.tag} {kind:10s} fwd {f:.4f} ms ({fb / f / 1e6 / peak:.3f})  bwd {b:.4f} ms ({bb / b / 1e6 / peak:.3f})  "
          f"fwd+bwd {f + b:.4f} ms ({(fb + bb) / (f + b) / 1e6 / peak:.3f})  img {float(image.sum()):.1f}", flush=True)
