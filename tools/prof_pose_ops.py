"""torch.profiler op counts of the pose-optimisation loop: python tools/prof_pose_ops.py"""
import os, sys
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "examples"))
import pertrenderer_b200 as pb
import pose_optimisation as po
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
po.optimize_pose(mesh, verts, renderer, target, w0, 10, 5e-2, False)
torch.cuda.synchronize()
n = 20
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    po.optimize_pose(mesh, verts, renderer, target, w0, n, 5e-2, False)
    torch.cuda.synchronize()
ev = prof.key_averages()
rows = sorted(ev, key=lambda e: -e.self_cpu_time_total)[:40]
tot_calls = sum(e.count for e in ev if e.key.startswith("aten::")) / n
print(f"aten op calls per iteration: {tot_calls:.0f}; cuda kernel launches per iteration: {sum(e.count for e in ev if e.key == 'cudaLaunchKernel') / n:.0f}")
for e in rows:
    print(f"{e.key[:60]:60s} calls/it {e.count / n:6.1f} self cpu us/it {e.self_cpu_time_total / n:8.1f} cuda us/it {getattr(e, 'self_device_time_total', 0) / n:8.1f}")
