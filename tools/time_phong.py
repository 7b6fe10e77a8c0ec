"""Kernel-only timing of the Phong passes at BASELINE config 2 (CUDA events around back-to-back launches; outputs
pre-allocated by the launcher each call, so the torch allocations are inside).  python tools/time_phong.py [n_faces]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pertrenderer_b200 as pb  # noqa: E402
from pertrenderer_b200 import ops, shading  # noqa: E402

N, HW, K, S = 8, 256, 50, 64
dev = "cuda:0"


def setup(F, kind="realistic"):
    fr, _ = pb.synthetic_fragments(N, HW, HW, K, kind=kind, sigma=1e-3, n_faces=F, seed=0, device=dev)
    verts, faces = pb.synthetic_mesh(F, device=dev)
    F = faces.shape[0]
    p2f = fr.pix_to_face.clamp(max=F - 1)
    bary = pb.synthetic_bary(p2f)
    mesh = pb.TriMeshes(verts, faces)
    fv, fn = verts[faces].contiguous(), mesh.verts_normals_packed()[faces].contiguous()
    fc = torch.rand((F, 3), device=dev)
    lighting = shading.pack_lighting(pb.PointLights(location=[[0.0, 2.0, -2.0]], device=dev), pb.Materials(device=dev),
                                     pb.ViewCameras(R=torch.eye(3)[None], T=[[0.0, 0.0, 6.7]], device=dev), N, dev)
    G = torch.randn((N, HW, HW, 4), device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    colors = shading.phong_forward(p2f, bary, fv, fn, None, fc, lighting, sparse=True)
    pr = ops.ShadeProblem(pix_to_face=p2f, zbuf=fr.zbuf, dists=fr.dists, colors=colors, znear=1.0, zfar=100.0,
                          background=(1.0, 1.0, 1.0), sigma=1e-3, gamma=1e-2, alpha=1.0, eps=1e-10, S_rast=S, S_agg=S,
                          seed_rast=100, seed_agg=200)
    image, saved = ops.shade_forward(pr)
    gc = ops.shade_backward(pr, saved, G)[2]
    mask = p2f >= 0
    print(f"F={F} valid {mask.float().mean().item():.4f} heavy {((gc != 0).any(-1) & mask).float().mean().item():.4f}")
    return p2f, bary, fv, fn, fc, lighting, gc


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for F in [int(a) for a in sys.argv[1:]] or [1280, 300, 100000]:
    p2f, bary, fv, fn, fc, lighting, gc = setup(F)
    print("  fwd sparse      %7.1f us" % timeit(lambda: shading.phong_forward(p2f, bary, fv, fn, None, fc, lighting, sparse=True)))
    print("  fwd dense       %7.1f us" % timeit(lambda: shading.phong_forward(p2f, bary, fv, fn, None, fc, lighting)))
    for name, kw in [("bwd all", {}), ("bwd no face grads", dict(need_verts=False, need_normals=False)),
                     ("bwd no bary", dict(need_bary=False)), ("bwd verts only", dict(need_bary=False, need_normals=False))]:
        print("  %-18s %7.1f us" % (name, timeit(lambda: shading.phong_backward(p2f, bary, fv, fn, None, fc, lighting, gc,
                                                                                need_texels=False, sparse=True, **kw))))
