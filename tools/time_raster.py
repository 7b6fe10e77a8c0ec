"""Kernel-level timing of the fragment producer at BASELINE config 2 shapes (8 views x 256x256, K = 50, icosphere of
1280 faces = sphere_642.obj, blur = log(1/1e-4 - 1) * sigma), and of the whole renderer step.
    python tools/time_raster.py [n_faces] [image_size] [views] [sigma]"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pertrenderer_b200 as pb  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 1280
HW = int(sys.argv[2]) if len(sys.argv) > 2 else 256
N = int(sys.argv[3]) if len(sys.argv) > 3 else 8
sigma = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-3
K, S = 50, 64
dev = "cuda:0"
verts, faces = pb.synthetic_mesh(F, device=dev)
R, T = pb.look_at_view_transform(dist=2.7, elev=30.0, azim=torch.linspace(0, 315, N), device=dev)
cam = pb.OpenGLPerspectiveCameras(R=R, T=T, device=dev)
blur = math.log(1.0 / 1e-4 - 1.0) * sigma
mesh = pb.TriMeshes(verts, faces, face_colors=torch.rand(faces.shape[0], 3, device=dev)).extend(N)
ndc = cam.transform_points_ndc(mesh.verts_padded())
fv = ndc[:, faces].reshape(-1, 3, 3).contiguous()
start = torch.arange(N + 1, device=dev) * faces.shape[0]


def timeit(fn, n=10):
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


p2f, zbuf, bary, dists = pb.rasterize_meshes(fv, start, HW, blur, K)
valid = (p2f >= 0)
print(f"F={faces.shape[0]} {N}x{HW}x{HW} K={K} blur={blur:.2e}: coverage {valid.any(-1).float().mean():.3f} "
      f"mean valid/covered pixel {valid.sum(-1)[valid.any(-1)].float().mean():.2f} max {valid.sum(-1).max().item()}")
print("  rasterize fwd  %8.1f us" % timeit(lambda: pb.rasterize_meshes(fv, start, HW, blur, K)))
fvg = fv.clone().requires_grad_(True)
out = pb.rasterize_meshes(fvg, start, HW, blur, K)
gz, gb, gd = torch.randn_like(out[1]), torch.randn_like(out[2]), torch.randn_like(out[3])


def bwd():
    fvg.grad = None
    torch.autograd.backward([out[1], out[2], out[3]], [gz, gb, gd], retain_graph=True)


print("  rasterize bwd  %8.1f us" % timeit(bwd))

renderer = pb.MeshRenderer(
    pb.MeshRasterizer(cam, pb.RasterizationSettings(image_size=HW, blur_radius=blur, faces_per_pixel=K)),
    pb.RandomPhongShader(device=dev, cameras=cam, lights=pb.PointLights(location=[[0.0, 2.0, -2.0]], device=dev),
                         smoothrast=pb.GaussianRast(nb_samples=S, sigma=sigma), smoothagg=pb.GaussianAgg(nb_samples=S, gamma=1e-2)))
G = torch.randn(N, HW, HW, 4, device=dev)
v = mesh.verts_padded().clone().requires_grad_(True)


def step():
    v.grad = None
    img = renderer(mesh.update_padded(v))
    (img * G).sum().backward()


print("  renderer fwd+bwd (public API, autograd, incl. host sync of the scalar grads) %8.1f us" % timeit(step))
