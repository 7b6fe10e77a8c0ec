import cProfile, pstats, sys, math, torch, io
sys.path.insert(0, '.'); sys.path.insert(0, 'examples')
import pertrenderer_b200 as pb
import pose_optimisation as po
dev = "cuda:0"
verts, faces, colors = po.cube_mesh(dev)
mesh = pb.TriMeshes(verts, faces, face_colors=colors)
R, T = pb.look_at_view_transform(dist=6.7, elev=30.0, azim=120.0, device=dev)
cameras = pb.OpenGLPerspectiveCameras(R=R, T=T, fov=60, device=dev)
lights = pb.PointLights(location=[[0.0, 2.0, -2.0]], device=dev)
renderer = po.make_renderer("gaussian", cameras, lights, 1e-3, 1e-2, 16, 128, dev)
target = torch.rand(1, 128, 128, 3, device=dev)
w0 = torch.tensor([0.3, -0.2, 0.5], device=dev)
po.optimize_pose(mesh, verts, renderer, target, w0, 20, 5e-2, False)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
po.optimize_pose(mesh, verts, renderer, target, w0, 100, 5e-2, False)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
