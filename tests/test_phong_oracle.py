"""CPU checks of the Phong row: closed-form known answers of the oracle (oracle/phong_oracle.py restates
pytorch3d 0.4.0's phong_shading; no reference-side fixture exists at that boundary), the shim objects,
and argument validation of the C ABI without a GPU."""

import ctypes
import math

import pytest
import torch

import pertrenderer_b200 as pb
from oracle import phong_oracle as PO
from pertrenderer_b200 import _cabi


def _one_entry(light_location=None, light_direction=None, camera=(0.0, 0.0, 5.0), shininess=8.0, bary=(0.2, 0.3, 0.5),
               valid=True):
    """One pixel, one face in the plane z = 0 with vertex normals +z, the shaded point inside the face."""
    fv = torch.tensor([[[-1.0, -1.0, 0.0], [3.0, -1.0, 0.0], [-1.0, 3.0, 0.0]]])
    fn = torch.tensor([[[0.0, 0.0, 1.0]] * 3])
    p2f = torch.tensor([[[[0 if valid else -1]]]])
    b = torch.tensor(bary).reshape(1, 1, 1, 1, 3)
    tex = torch.tensor([0.2, 0.5, 0.9]).reshape(1, 1, 1, 1, 3)
    kw = dict(light_ambient=torch.tensor([[0.5, 0.4, 0.3]]), light_diffuse=torch.tensor([[0.3, 0.2, 0.1]]),
              light_specular=torch.tensor([[0.2, 0.3, 0.4]]), mat_ambient=torch.tensor([[1.0, 0.9, 0.8]]),
              mat_diffuse=torch.tensor([[0.7, 0.8, 0.9]]), mat_specular=torch.tensor([[0.6, 0.5, 0.4]]),
              shininess=torch.tensor([shininess]), camera_center=torch.tensor([camera]))
    if light_location is not None:
        kw["light_location"] = torch.tensor([light_location])
    else:
        kw["light_direction"] = torch.tensor([light_direction])
    point = (b.reshape(3, 1) * fv[0]).sum(0)
    return PO.phong_colors(p2f, b, fv, fn, tex, **kw).reshape(3), tex.reshape(3), kw, point


def test_head_on_light_and_viewer():
    # light and camera straight above the shaded point: n.d = 1, reflection = n, v.r = 1
    bary = (0.25, 0.25, 0.5)
    pt = torch.tensor([-1.0, -1.0, 0.0]) * 0.25 + torch.tensor([3.0, -1.0, 0.0]) * 0.25 + torch.tensor([-1.0, 3.0, 0.0]) * 0.5
    above = (pt[0].item(), pt[1].item(), 5.0)
    col, tex, kw, _ = _one_entry(light_location=above, camera=above, bary=bary)
    amb = kw["mat_ambient"][0] * kw["light_ambient"][0]
    dif = kw["mat_diffuse"][0] * kw["light_diffuse"][0]
    spc = kw["mat_specular"][0] * kw["light_specular"][0]
    assert torch.allclose(col, (amb + dif) * tex + spc, atol=1e-6)


def test_back_facing_light_gives_ambient_only():
    col, tex, kw, _ = _one_entry(light_location=(0.3, 0.2, -4.0))
    amb = kw["mat_ambient"][0] * kw["light_ambient"][0]
    assert torch.allclose(col, amb * tex, atol=1e-7)


def test_directional_light_45_degrees_mirror_and_off_mirror_viewer():
    s = 1.0 / math.sqrt(2.0)
    _, _, _, pt = _one_entry(light_direction=(1.0, 0.0, 1.0))
    mirror_cam = (pt[0].item() - 3.0, pt[1].item(), 3.0)  # viewer along the reflected ray (-s, 0, s)
    col, tex, kw, _ = _one_entry(light_direction=(1.0, 0.0, 1.0), camera=mirror_cam)
    amb = kw["mat_ambient"][0] * kw["light_ambient"][0]
    dif = kw["mat_diffuse"][0] * kw["light_diffuse"][0]
    spc = kw["mat_specular"][0] * kw["light_specular"][0]
    assert torch.allclose(col, (amb + dif * s) * tex + spc, atol=1e-5)
    above = (pt[0].item(), pt[1].item(), 7.0)  # viewer along the normal: v.r = cos 45
    col2, _, _, _ = _one_entry(light_direction=(1.0, 0.0, 1.0), camera=above, shininess=8.0)
    assert torch.allclose(col2, (amb + dif * s) * tex + spc * s ** 8, atol=1e-5)


def test_padded_entry_keeps_only_ambient_times_texel():
    # masked interpolation gives a zero point and a zero normal: no diffuse, no specular; pytorch3d's
    # sample_textures returns zero texels there, so the colour of a padded entry is black in practice
    col, tex, kw, _ = _one_entry(light_location=(0.0, 0.0, 4.0), valid=False)
    assert torch.allclose(col, kw["mat_ambient"][0] * kw["light_ambient"][0] * tex, atol=1e-7)


def test_oracle_gradients_match_finite_differences():
    torch.manual_seed(0)
    F_, N, H, W, K = 6, 1, 2, 2, 3
    fv = torch.randn(F_, 3, 3, dtype=torch.float64)
    fn = torch.randn(F_, 3, 3, dtype=torch.float64)
    p2f = torch.randint(-1, F_, (N, H, W, K))
    bary = torch.rand(N, H, W, K, 3, dtype=torch.float64)
    tex = torch.rand(N, H, W, K, 3, dtype=torch.float64)
    kw = dict(light_location=torch.tensor([[0.0, 2.0, -2.0]], dtype=torch.float64),
              light_ambient=torch.full((1, 3), 0.5, dtype=torch.float64), light_diffuse=torch.full((1, 3), 0.3, dtype=torch.float64),
              light_specular=torch.full((1, 3), 0.2, dtype=torch.float64), mat_ambient=torch.ones(1, 3, dtype=torch.float64),
              mat_diffuse=torch.ones(1, 3, dtype=torch.float64), mat_specular=torch.ones(1, 3, dtype=torch.float64),
              shininess=torch.tensor([4.0], dtype=torch.float64), camera_center=torch.tensor([[0.0, 0.0, 6.7]], dtype=torch.float64))
    f = lambda a, b, c, d: PO.phong_colors(p2f, b, a, c, d, **kw)  # noqa: E731
    assert torch.autograd.gradcheck(f, (fv.requires_grad_(), bary.requires_grad_(), fn.requires_grad_(), tex.requires_grad_()),
                                    eps=1e-6, atol=1e-5)


def test_shims_expose_the_pytorch3d_attributes():
    lights, mats = pb.PointLights(location=[[0.0, 2.0, -2.0]]), pb.Materials()
    assert lights.location.shape == (1, 3) and lights.ambient_color.shape == (1, 3)
    assert mats.shininess.item() == 64 and mats.specular_color.shape == (1, 3)
    assert not hasattr(pb.DirectionalLights(), "location")
    # camera at distance d on the +z axis looking at the origin: R = diag(-1, 1, -1), T = (0, 0, d)
    cam = pb.ViewCameras(R=torch.diag(torch.tensor([-1.0, 1.0, -1.0]))[None], T=[[0.0, 0.0, 6.7]])
    assert torch.allclose(cam.get_camera_center(), torch.tensor([[0.0, 0.0, 6.7]]), atol=1e-6)
    verts, faces = pb.synthetic_mesh(200, device="cpu")
    mesh = pb.TriMeshes(verts, faces, face_colors=torch.rand(faces.shape[0], 3))
    n = mesh.verts_normals_packed()
    assert torch.allclose(n.norm(dim=1), torch.ones(n.shape[0]), atol=1e-5)
    # a sphere's vertex normals point outwards
    assert (torch.sum(n * verts, dim=1) > 0.9).float().mean() > 0.95
    sh = pb.RandomPhongShader(cameras=None, smoothrast=pb.GaussianRast(), smoothagg=pb.GaussianAgg())
    with pytest.raises(ValueError, match="Cameras"):
        sh(None, None)


def test_phong_struct_layout_and_validation_without_gpu():
    assert ctypes.sizeof(_cabi.PertPhong) == 144
    assert _cabi.PertPhong.pix_to_face.offset == 40 and _cabi.PertPhong.faces_per_mesh.offset == 136
    lib = _cabi.load()
    ph = _cabi.PertPhong()
    assert lib.pert_phong_fwd(None, None, None) == -1
    ph.P, ph.HW, ph.K, ph.num_faces, ph.light_rows = 8, 3, 2, 4, 1
    assert lib.pert_phong_fwd(ph, None, None) == -2  # P not a multiple of HW
    ph.HW = 4
    ph.light_rows = 3
    assert lib.pert_phong_fwd(ph, None, None) == -2  # rows must be 1 or N
    ph.light_rows = 2
    assert lib.pert_phong_fwd(ph, None, None) == -1  # null inputs
    ph.flags = _cabi.PHONG_UNLIT
    ph.light_rows = 0
    assert lib.pert_phong_fwd(ph, None, None) == -1  # unlit: no lighting rows needed, inputs still NULL
    assert lib.pert_phong_bwd(ph, None, None, None, None, None, None, None) == -1


def test_phong_cpu_tensors_fail_loudly():
    verts, faces = pb.synthetic_mesh(20, device="cpu")
    mesh = pb.TriMeshes(verts, faces, face_colors=torch.rand(faces.shape[0], 3))
    p2f = torch.zeros(1, 2, 2, 3, dtype=torch.int64)
    frag = pb.Fragments(p2f, torch.ones(1, 2, 2, 3), torch.rand(1, 2, 2, 3, 3), torch.zeros(1, 2, 2, 3))
    cam = pb.ViewCameras(R=torch.eye(3)[None], T=[[0.0, 0.0, 3.0]])
    with pytest.raises(RuntimeError, match="CUDA"):
        pb.phong_shading(mesh, frag, pb.PointLights(), cam, pb.Materials(), mesh.sample_textures(frag))


def test_uv_texels_on_a_striped_map_are_constant_per_face():
    """The reference's cube (eval.py:727-757): six colour strips in one map, ONE UV point per side, so every face samples a
    constant colour whatever the bilinear details; and a UV outside [0,1] clamps to the border texel."""
    strip = torch.tensor([[0.9, 0.1, 0.1], [0.1, 0.7, 0.1], [0.1, 0.2, 0.9]])
    cmap = strip.repeat_interleave(4, dim=0)[None].expand(5, 12, 3).contiguous()  # (Hm=5, Wm=12, 3)
    vt = torch.tensor([[1.5 / 11, 0.5], [5.5 / 11, 0.5], [9.5 / 11, 0.5], [-3.0, 0.2], [7.0, 0.9]])
    fuv = torch.tensor([[0, 0, 0], [1, 1, 1], [2, 2, 2], [3, 3, 3], [4, 4, 4]])
    p2f = torch.tensor([0, 1, 2, 3, 4, -1]).reshape(1, 1, 1, 6)
    bary = torch.tensor([0.2, 0.3, 0.5]).expand(1, 1, 1, 6, 3).contiguous()
    tex = pb.UVTexels(cmap, vt, fuv).materialize(p2f, bary)[0, 0, 0]
    assert torch.allclose(tex[:3], strip, atol=1e-6)
    assert torch.allclose(tex[3], strip[0], atol=1e-6) and torch.allclose(tex[4], strip[2], atol=1e-6)  # border clamp
    assert (tex[5] == 0).all()  # padded entry


def test_lighting_table_is_differentiable_and_batches():
    from pertrenderer_b200 import shading
    loc = torch.tensor([[0.0, 2.0, -2.0]], requires_grad=True)
    lights = pb.PointLights(location=loc)
    mats = pb.Materials(shininess=torch.tensor([8.0, 16.0]))
    R, T = pb.look_at_view_transform(dist=2.7, elev=[10.0, 20.0], azim=[0.0, 90.0])
    T.requires_grad_(True)
    cams = pb.FoVPerspectiveCameras(R=R, T=T)
    table = shading.pack_lighting(lights, mats, cams, 2, "cpu")
    assert table.shape == (2, 20) and table[:, 12].tolist() == [8.0, 16.0] and (table[:, 16] == 0).all()
    assert torch.allclose(table[:, 13:16].norm(dim=1), torch.full((2,), 2.7), atol=1e-5)  # camera centres
    table[:, :3].sum().backward(retain_graph=True)
    assert torch.allclose(loc.grad, torch.full((1, 3), 2.0))  # one light row broadcast over two images
    table[:, 13:16].sum().backward()
    assert T.grad is not None and T.grad.abs().sum() > 0
    d = shading.pack_lighting(pb.DirectionalLights(), pb.Materials(), cams, 2, "cpu")
    assert (d[:, 16] == 1).all()


def test_trimeshes_batches_and_textures():
    verts, faces = pb.synthetic_mesh(20, device="cpu")
    m = pb.TriMeshes(verts, faces, face_colors=torch.rand(20, 3))
    assert len(m) == 1 and m.verts_padded().shape == (1, 12, 3)
    m3 = m.extend(3)
    assert len(m3) == 3 and m3.verts_packed().shape == (36, 3) and m3.faces_packed().max().item() == 35
    assert isinstance(m3.sample_textures(None), pb.FaceTexels) and m3.sample_textures(None).face_colors.shape == (60, 3)
    mv = pb.TriMeshes(verts, faces, verts_colors=torch.rand(12, 3)).extend(2)
    t = mv.sample_textures(None)
    assert isinstance(t, pb.VertexTexels) and t.face_vert_colors().shape == (40, 3, 3)
    mu = pb.TriMeshes(verts, faces, uv=(torch.rand(1, 4, 4, 3), torch.rand(7, 2), torch.randint(0, 7, (20, 3)))).extend(2)
    assert isinstance(mu.sample_textures(None), pb.UVTexels) and mu.sample_textures(None).face_uvs().shape == (40, 3, 2)
    moved = m3.update_padded(m3.verts_padded() + 1.0)
    assert torch.allclose(moved.verts_packed(), m3.verts_packed() + 1.0) and moved.face_colors is m3.face_colors
    n = m3.verts_normals_packed()
    assert n.shape == (36, 3) and torch.allclose(n[:12], n[12:24], atol=1e-6)  # same pose, same normals
