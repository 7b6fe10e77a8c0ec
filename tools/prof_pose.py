"""cProfile of the pose-optimisation loop (host side): python tools/prof_pose.py [niter]"""
import cProfile, pstats, io, math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "examples"))
import pertrenderer_b200 as pb
import pose_optimisation as po
niter = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = "cuda:0"
verts, faces, colors = po.cube_mesh(dev)
mesh = pb.TriMeshes(verts, faces, face_colors=colors)
R, T = pb.look_at_view_transform(dist=6.7, elev=30.0, azim=120.0, device=dev)
cameras = pb.OpenGLPerspectiveCameras(R=R, T=T, fov=60, device=dev)
lights = pb.PointLights(location=[[0.0, 2.0, -2.0]], device=dev)
hard = po.make_renderer("hard", cameras, lights, 1e-4, 1e-4, 1, 128, dev)
gen = torch.Generator().manual_seed(0)
R_true = po.random_rotation(gen, dev)
with torch.no_grad():
    target = hard(mesh.update_padded(verts @ R_true))[..., :3]
w0 = po.so3_log(R_true @ po.so3_exp(torch.tensor([0.3, 0.2, 0.1], device=dev)))
renderer = po.make_renderer("gaussian", cameras, lights, 1e-3, 1e-2, 16, 128, dev)
po.optimize_pose(mesh, verts, renderer, target, w0, 20, 5e-2, False)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
po.optimize_pose(mesh, verts, renderer, target, w0, niter, 5e-2, False)
torch.cuda.synchronize()
print("ms per iteration (no profiler): %.3f" % ((time.perf_counter() - t0) / niter * 1e3))
pr = cProfile.Profile()
pr.enable()
po.optimize_pose(mesh, verts, renderer, target, w0, niter, 5e-2, False)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print("\n".join(l[:150] for l in s.getvalue().splitlines()[:75]))
