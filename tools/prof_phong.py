"""Small driver for ncu: a few RandomPhongShader steps (Phong fwd -> fused shade fwd/bwd -> Phong bwd) of the
BASELINE config-2 workload.    python tools/prof_phong.py [realistic|dense] [steps] [n_faces]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pertrenderer_b200 as pb  # noqa: E402
from pertrenderer_b200 import ops, shading  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "realistic"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
F = int(sys.argv[3]) if len(sys.argv) > 3 else 1280
N, HW, K, S = 8, 256, 50, 64
dev = "cuda:0"
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
for i in range(steps):
    colors = shading.phong_forward(p2f, bary, fv, fn, None, fc, lighting, sparse=True)
    pr = ops.ShadeProblem(pix_to_face=p2f, zbuf=fr.zbuf, dists=fr.dists, colors=colors, znear=1.0, zfar=100.0,
                          background=(1.0, 1.0, 1.0), sigma=1e-3, gamma=1e-2, alpha=1.0, eps=1e-10, S_rast=S, S_agg=S,
                          seed_rast=100 + i, seed_agg=200 + i)
    image, saved = ops.shade_forward(pr)
    gd, gz, gc, scal = ops.shade_backward(pr, saved, G)
    out = shading.phong_backward(p2f, bary, fv, fn, None, fc, lighting, gc, need_texels=False, sparse=True)
torch.cuda.synchronize()
print("ok", float(image.sum()), float(out[1].sum()), float(out[2].sum()))
