"""Parity of the Phong kernels (pert_phong_fwd / pert_phong_bwd, through the C ABI) with the CPU oracle
(oracle/phong_oracle.py, pytorch3d 0.4.0's phong_shading restated), and of RandomPhongShader end to end
(random_rasterizer.py:60-130) with the oracle chain Phong -> perturbed blend fed the same noise.
Tolerance: 1e-5 relative in fp32 (north_star); face-table gradients are summed by atomics in no fixed order."""

import pytest
import torch

from conftest import rel_err
from oracle import pert_oracle as O
from oracle import phong_oracle as PO

pytestmark = pytest.mark.gpu

RTOL = 1e-5
DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _scene(N, H, W, K, n_faces, light="point", per_batch=False, seed=0, shininess=8.0, kind="realistic"):
    """Everything on the CPU: fragments with bary, a sphere mesh, lights / materials / cameras shims."""
    import pertrenderer_b200 as pb
    fr, _ = pb.synthetic_fragments(N, H, W, K, kind=kind, n_faces=n_faces, seed=seed, device="cpu", mean_valid=3.0)
    verts, faces = pb.synthetic_mesh(n_faces, device="cpu")
    fr = pb.Fragments(fr.pix_to_face.clamp(max=faces.shape[0] - 1), fr.zbuf, pb.synthetic_bary(fr.pix_to_face, seed), fr.dists)
    g = torch.Generator().manual_seed(seed + 3)
    rows = N if per_batch else 1
    rnd = lambda *s: torch.rand(*s, generator=g)  # noqa: E731
    col = dict(ambient_color=0.2 + 0.5 * rnd(rows, 3), diffuse_color=0.2 + 0.5 * rnd(rows, 3), specular_color=0.2 + 0.5 * rnd(rows, 3))
    if light == "point":
        lights = pb.PointLights(location=torch.tensor([[0.0, 2.0, -2.0]]) + rnd(rows, 3), **col)
    else:
        lights = pb.DirectionalLights(direction=torch.tensor([[0.3, 1.0, -0.5]]) + 0.3 * rnd(rows, 3), **col)
    mats = pb.Materials(ambient_color=0.5 + 0.5 * rnd(rows, 3), diffuse_color=0.5 + 0.5 * rnd(rows, 3),
                        specular_color=0.5 + 0.5 * rnd(rows, 3), shininess=torch.full((rows,), shininess))
    cams = pb.ViewCameras(R=torch.eye(3).expand(N, 3, 3).clone(), T=torch.tensor([[0.0, 0.0, 2.7]]) + 0.2 * rnd(N, 3))
    face_colors = rnd(faces.shape[0], 3)
    texels = rnd(N, H, W, K, 3) * (fr.pix_to_face >= 0)[..., None]
    return fr, verts, faces, lights, mats, cams, face_colors, texels


def _to(obj, dev):
    import copy
    return copy.deepcopy(obj).to(dev)


def _frag_to(fr, dev, bary_grad=False):
    import pertrenderer_b200 as pb
    b = fr.bary_coords.to(dev)
    if bary_grad:
        b.requires_grad_(True)
    return pb.Fragments(fr.pix_to_face.to(dev), fr.zbuf.to(dev), b, fr.dists.to(dev))


@pytest.mark.parametrize("cfg", [
    dict(n_faces=40, light="point", per_batch=False, face_mode=False),      # small shared-memory gradient table
    dict(n_faces=1200, light="point", per_batch=True, face_mode=True),      # one-CTA-per-SM gradient table
    dict(n_faces=1200, light="directional", per_batch=False, face_mode=False),
    dict(n_faces=40, light="directional", per_batch=True, face_mode=True),  # per-batch rows, face colours, table
    dict(n_faces=3000, light="point", per_batch=True, face_mode=False),     # global atomics
    dict(n_faces=3000, light="directional", per_batch=False, face_mode=True),
])
def test_phong_kernels_match_oracle(cfg):
    import pertrenderer_b200 as pb
    N, H, W, K = 3, 9, 11, 7
    fr, verts, faces, lights, mats, cams, face_colors, texels = _scene(N, H, W, K, cfg["n_faces"], cfg["light"], cfg["per_batch"],
                                                                       seed=cfg["n_faces"])
    gen = torch.Generator().manual_seed(1)
    grad_colors = torch.randn(N, H, W, K, 3, generator=gen)
    grad_colors[torch.rand(N, H, W, K, generator=gen) < 0.5] = 0.0  # entries no sample picked: the kernel's shortcut

    # oracle (autograd over the restatement)
    v_o, t_o = verts.clone().requires_grad_(True), (face_colors if cfg["face_mode"] else texels).clone().requires_grad_(True)
    fr_o = pb.Fragments(fr.pix_to_face, fr.zbuf, fr.bary_coords.clone().requires_grad_(True), fr.dists)
    mesh_o = pb.TriMeshes(v_o, faces, face_colors=t_o) if cfg["face_mode"] else pb.TriMeshes(v_o, faces, texels=t_o)
    tex_o = pb.FaceTexels(t_o).materialize(fr.pix_to_face) if cfg["face_mode"] else t_o
    col_o = PO.phong_colors_from(mesh_o, fr_o, lights, cams, mats, tex_o)
    (col_o * grad_colors).sum().backward()

    # CUDA
    v_c = verts.to(DEV).requires_grad_(True)
    t_c = (face_colors if cfg["face_mode"] else texels).to(DEV).requires_grad_(True)
    fr_c = _frag_to(fr, DEV, bary_grad=True)
    mesh_c = pb.TriMeshes(v_c, faces.to(DEV), face_colors=t_c) if cfg["face_mode"] else pb.TriMeshes(v_c, faces.to(DEV), texels=t_c)
    col_c = pb.phong_shading(mesh_c, fr_c, _to(lights, DEV), _to(cams, DEV), _to(mats, DEV), mesh_c.sample_textures(fr_c))
    (col_c * grad_colors.to(DEV)).sum().backward()

    assert (col_c.detach().cpu() - col_o.detach()).abs().max() <= 2e-6
    assert rel_err(col_c.detach().cpu(), col_o.detach()) <= RTOL
    assert rel_err(t_c.grad.cpu(), t_o.grad) <= RTOL
    # the vertex normals come from torch's index_add on the GPU (summation order not fixed, last-bit differences):
    # a specular entry with alpha^(shininess-1) amplification moves by 5e-5 relative between runs (300 repetitions)
    assert rel_err(fr_c.bary_coords.grad.cpu(), fr_o.bary_coords.grad) <= 5 * RTOL
    assert rel_err(v_c.grad.cpu(), v_o.grad) <= 5 * RTOL  # and thousands of atomic adds per vertex, order not fixed
    mask = fr.pix_to_face >= 0
    assert (fr_c.bary_coords.grad.cpu()[~mask] == 0).all()


def test_phong_sparse_flag_leaves_padding_alone_and_keeps_valid_entries():
    import pertrenderer_b200 as pb
    from pertrenderer_b200 import shading
    N, H, W, K = 2, 8, 8, 6
    fr, verts, faces, lights, mats, cams, face_colors, texels = _scene(N, H, W, K, 60, seed=4)
    mesh = pb.TriMeshes(verts.to(DEV), faces.to(DEV))
    # one set of face tables for both launches (torch's index_add behind verts_normals_packed sums in no fixed order)
    fv, fn = mesh.verts_packed()[mesh.faces_packed()].contiguous(), mesh.verts_normals_packed()[mesh.faces_packed()].contiguous()
    lighting = shading.pack_lighting(_to(lights, DEV), _to(mats, DEV), _to(cams, DEV), N, DEV)
    p2f, bary, tex = fr.pix_to_face.to(DEV), fr.bary_coords.to(DEV), texels.to(DEV)
    dense = shading.phong_forward(p2f, bary, fv, fn, tex, None, lighting)
    sparse = shading.phong_forward(p2f, bary, fv, fn, tex, None, lighting, sparse=True)
    mask = p2f >= 0
    assert torch.equal(dense[mask], sparse[mask])
    assert (dense[~mask] == 0).all()  # zero texels at padded entries -> black
    g = torch.randn(N, H, W, K, 3, device=DEV) * mask[..., None]
    a = shading.phong_backward(p2f, bary, fv, fn, tex, None, lighting, g)
    b = shading.phong_backward(p2f, bary, fv, fn, tex, None, lighting, g, sparse=True)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert rel_err(b[2], a[2]) <= 1e-6 and rel_err(b[3], a[3]) <= 1e-6


def test_degenerate_normals_and_shininess_zero():
    """Zero interpolated normal (clamped normalisation), light at the shaded point, shininess 0 (0^0 = 1)."""
    import pertrenderer_b200 as pb
    fv = torch.tensor([[[-1.0, -1.0, 0.0], [3.0, -1.0, 0.0], [-1.0, 3.0, 0.0]]])
    verts = fv.reshape(3, 3)
    faces = torch.tensor([[0, 1, 2]])
    p2f = torch.zeros(1, 1, 2, 2, dtype=torch.int64)
    p2f[0, 0, 1, 1] = -1
    bary = torch.tensor([0.25, 0.25, 0.5]).expand(1, 1, 2, 2, 3).contiguous()
    fr = pb.Fragments(p2f, torch.ones(1, 1, 2, 2), bary, torch.zeros(1, 1, 2, 2))
    texels = torch.rand(1, 1, 2, 2, 3)

    class ZeroNormalMesh(pb.TriMeshes):
        def verts_normals_packed(self):
            return torch.zeros_like(self.verts_packed())

    for sh in (0.0, 1.0, 64.0):
        lights = pb.PointLights(location=[[0.0, 0.5, 0.0]])  # ON the shaded point (0, 0.5, 0): zero light direction
        mats = pb.Materials(shininess=sh)
        cams = pb.ViewCameras(R=torch.eye(3)[None], T=[[0.0, 0.0, 2.7]])
        for mesh_cls in (pb.TriMeshes, ZeroNormalMesh):
            ref = PO.phong_colors_from(mesh_cls(verts, faces, texels=texels), fr, lights, cams, mats, texels)
            got = pb.phong_shading(mesh_cls(verts.to(DEV), faces.to(DEV), texels=texels.to(DEV)), _frag_to(fr, DEV), _to(lights, DEV),
                                   _to(cams, DEV), _to(mats, DEV), texels.to(DEV))
            assert torch.isfinite(got).all()
            assert (got.cpu() - ref).abs().max() <= 2e-6, (sh, mesh_cls.__name__)


@pytest.mark.parametrize("pair", ["gaussian", "softras"])
def test_random_phong_shader_matches_oracle_chain(pair):
    """RandomPhongShader.forward + autograd vs oracle Phong colours -> oracle blend, same noise: image, and the
    gradient that reaches the mesh vertices (the quantity pose optimisation uses, eval.py:343-369)."""
    import pertrenderer_b200 as pb
    N, H, W, K, S = 2, 10, 9, 12, 16
    fr, verts, faces, lights, mats, cams, face_colors, _ = _scene(N, H, W, K, 80, per_batch=True, seed=9)
    sigma, gamma, alpha = 1e-3, 1e-2, 1.0
    background = (0.2, 0.4, 0.6)
    gen = torch.Generator().manual_seed(2)
    grad_image = torch.randn(N, H, W, 4, generator=gen)

    v_o = verts.clone().requires_grad_(True)
    mesh_o = pb.TriMeshes(v_o, faces, face_colors=face_colors)
    col_o = PO.phong_colors_from(mesh_o, fr, lights, cams, mats, pb.FaceTexels(face_colors).materialize(fr.pix_to_face))
    zn, zf = cams.znear.reshape(-1, 1, 1, 1), cams.zfar.reshape(-1, 1, 1, 1)
    if pair == "gaussian":
        U, V = O.draw_noise((N, H, W, K), S, S, generator=torch.Generator().manual_seed(5))
        st, gr = O.shade_fwd_bwd(fr.pix_to_face, fr.zbuf, fr.dists, col_o.detach(), torch.tensor(background), zn, zf, sigma, gamma,
                                 alpha, 1e-10, U, V, grad_image)
        image_o = st.image
        rast, agg = pb.GaussianRast(nb_samples=S, sigma=sigma), pb.GaussianAgg(nb_samples=S, gamma=gamma, alpha=alpha)
    else:
        image_o, _, _, gr = O.soft_shade_fwd_bwd(fr.pix_to_face, fr.zbuf, fr.dists, col_o.detach(), torch.tensor(background), zn, zf, sigma,
                                           gamma, alpha, 1e-10, grad_image)
        rast, agg = pb.SoftRast(sigma=sigma), pb.SoftAgg(gamma=gamma, alpha=alpha)
    col_o.backward(gr["colors"])

    v_c = verts.to(DEV).requires_grad_(True)
    mesh_c = pb.TriMeshes(v_c, faces.to(DEV), face_colors=face_colors.to(DEV))
    shader = pb.RandomPhongShader(device=DEV, cameras=_to(cams, DEV), lights=_to(lights, DEV), materials=_to(mats, DEV),
                                  smoothrast=rast, smoothagg=agg, blend_params=pb.BlendParams(background_color=background))
    fr_c = _frag_to(fr, DEV)
    if pair == "gaussian":
        with pb.explicit_noise(U.to(DEV), V.to(DEV)):
            img = shader(fr_c, mesh_c)
            (img * grad_image.to(DEV)).sum().backward()
    else:
        img = shader(fr_c, mesh_c)
        (img * grad_image.to(DEV)).sum().backward()
    assert (img.detach().cpu() - image_o).abs().max() <= 3e-6
    assert rel_err(v_c.grad.cpu(), v_o.grad) <= 5 * RTOL


@pytest.mark.parametrize("n_faces", [40, 1200, 3000])
def test_vertex_colour_textures_match_oracle(n_faces):
    """TexturesVertex (experiments/eval.py:450): texel = barycentric interpolation of the vertex colours.
    (a) stand-alone sampling (texture-only mode of the kernel) against interpolate_face_attributes,
    (b) inside the Phong kernel, (c) through RandomSimpleShader with the SoftRas pair: gradients reach the
    vertex colours."""
    import pertrenderer_b200 as pb
    N, H, W, K = 2, 9, 8, 6
    fr, verts, faces, lights, mats, cams, _, _ = _scene(N, H, W, K, n_faces, seed=n_faces + 1)
    gen = torch.Generator().manual_seed(3)
    vcol = torch.rand(verts.shape[0], 3, generator=gen)
    grad = torch.randn(N, H, W, K, 3, generator=gen)
    grad[torch.rand(N, H, W, K, generator=gen) < 0.4] = 0.0

    # (a) sampling only
    vc_o = vcol.clone().requires_grad_(True)
    b_o = fr.bary_coords.clone().requires_grad_(True)
    tex_o = PO.interpolate_face_attributes(fr.pix_to_face, b_o, vc_o[faces])
    (tex_o * grad).sum().backward()
    vc_c = vcol.to(DEV).requires_grad_(True)
    fr_c = _frag_to(fr, DEV, bary_grad=True)
    tex_c = pb.sample_lazy_textures(pb.VertexTexels(vc_c, faces.to(DEV)), fr_c)
    (tex_c * grad.to(DEV)).sum().backward()
    assert (tex_c.detach().cpu() - tex_o.detach()).abs().max() <= 1e-6
    assert rel_err(vc_c.grad.cpu(), vc_o.grad) <= 5 * RTOL
    assert rel_err(fr_c.bary_coords.grad.cpu(), b_o.grad) <= RTOL

    # (b) Phong with vertex colours
    vc_o2, v_o = vcol.clone().requires_grad_(True), verts.clone().requires_grad_(True)
    mesh_o = pb.TriMeshes(v_o, faces, verts_colors=vc_o2)
    col_o = PO.phong_colors_from(mesh_o, fr, lights, cams, mats, PO.interpolate_face_attributes(fr.pix_to_face, fr.bary_coords, vc_o2[faces]))
    (col_o * grad).sum().backward()
    vc_c2, v_c = vcol.to(DEV).requires_grad_(True), verts.to(DEV).requires_grad_(True)
    mesh_c = pb.TriMeshes(v_c, faces.to(DEV), verts_colors=vc_c2)
    fr_c2 = _frag_to(fr, DEV)
    col_c = pb.phong_shading(mesh_c, fr_c2, _to(lights, DEV), _to(cams, DEV), _to(mats, DEV), mesh_c.sample_textures(fr_c2))
    (col_c * grad.to(DEV)).sum().backward()
    assert (col_c.detach().cpu() - col_o.detach()).abs().max() <= 2e-6
    assert rel_err(vc_c2.grad.cpu(), vc_o2.grad) <= 5 * RTOL
    assert rel_err(v_c.grad.cpu(), v_o.grad) <= 5 * RTOL

    # (c) RandomSimpleShader + SoftRas pair on vertex colours
    gi = torch.randn(N, H, W, 4, generator=gen)
    vc_o3 = vcol.clone().requires_grad_(True)
    tex3 = PO.interpolate_face_attributes(fr.pix_to_face, fr.bary_coords, vc_o3[faces])
    image_o, _, _, gr = O.soft_shade_fwd_bwd(fr.pix_to_face, fr.zbuf, fr.dists, tex3.detach(), torch.tensor((1.0, 1.0, 1.0)),
                                             cams.znear.reshape(-1, 1, 1, 1), cams.zfar.reshape(-1, 1, 1, 1), 1e-3, 1e-2, 1.0, 1e-10, gi)
    tex3.backward(gr["colors"])
    vc_c3 = vcol.to(DEV).requires_grad_(True)
    shader = pb.RandomSimpleShader(device=DEV, cameras=_to(cams, DEV), smoothrast=pb.SoftRast(sigma=1e-3),
                                   smoothagg=pb.SoftAgg(gamma=1e-2, alpha=1.0))
    img = shader(_frag_to(fr, DEV), pb.TriMeshes(verts.to(DEV), faces.to(DEV), verts_colors=vc_c3))
    (img * gi.to(DEV)).sum().backward()
    assert (img.detach().cpu() - image_o).abs().max() <= 3e-6
    assert rel_err(vc_c3.grad.cpu(), vc_o3.grad) <= 5 * RTOL


def test_phong_batch_of_poses_uses_per_image_tables_and_matches_oracle():
    """N poses of one topology (TriMeshes.extend / update_padded): image n sees the packed faces [n F, (n+1) F); backward
    keeps one shared-memory gradient table per image (faces_per_mesh hint).  Same gradients as the oracle, and as the
    kernels without the hint."""
    import pertrenderer_b200 as pb
    from pertrenderer_b200 import shading
    N, H, W, K, F_ = 3, 64, 64, 24, 80  # per-image tables are used for meshes of up to 256 faces
    fr, verts, faces, lights, mats, cams, face_colors, _ = _scene(N, H, W, K, F_, per_batch=True, seed=21, kind="realistic")
    gen = torch.Generator().manual_seed(5)
    poses = verts[None] + 0.05 * torch.randn(N, verts.shape[0], 3, generator=gen)
    off = (torch.arange(N) * F_).view(N, 1, 1, 1)
    p2f = torch.where(fr.pix_to_face >= 0, fr.pix_to_face + off, fr.pix_to_face)  # image n -> faces of pose n
    fr = pb.Fragments(p2f, fr.zbuf, fr.bary_coords, fr.dists)
    grad = torch.randn(N, H, W, K, 3, generator=gen)
    grad[torch.rand(N, H, W, K, generator=gen) < 0.5] = 0.0
    v_o = poses.clone().requires_grad_(True)
    mesh_o = pb.TriMeshes(v_o, faces, face_colors=face_colors)
    tex_o = mesh_o.sample_textures(fr).materialize(p2f)
    (PO.phong_colors_from(mesh_o, fr, lights, cams, mats, tex_o) * grad).sum().backward()
    v_c = poses.to(DEV).requires_grad_(True)
    mesh_c = pb.TriMeshes(v_c, faces.to(DEV), face_colors=face_colors.to(DEV))
    assert len(mesh_c) == N and mesh_c.faces_packed().shape[0] == N * F_
    fr_c = _frag_to(fr, DEV)
    col = pb.phong_shading(mesh_c, fr_c, _to(lights, DEV), _to(cams, DEV), _to(mats, DEV), mesh_c.sample_textures(fr_c))
    (col * grad.to(DEV)).sum().backward()
    assert rel_err(v_c.grad.cpu(), v_o.grad) <= 5 * RTOL
    # the hint changes where the atomics land, not the result
    fv = v_c.detach().reshape(-1, 3)[mesh_c.faces_packed()].contiguous()
    fn = mesh_c.verts_normals_packed().detach()[mesh_c.faces_packed()].contiguous()
    lighting = shading.pack_lighting(_to(lights, DEV), _to(mats, DEV), _to(cams, DEV), N, DEV)
    fc = face_colors.to(DEV).repeat(N, 1)
    args = (fr_c.pix_to_face, fr_c.bary_coords, fv, fn, None, fc, lighting, grad.to(DEV))
    a = shading.phong_backward(*args, faces_per_mesh=F_)
    b = shading.phong_backward(*args, faces_per_mesh=0)
    for x, y in zip(a, b):
        assert rel_err(x, y) <= 1e-5


@pytest.mark.parametrize("light", ["point", "directional"])
def test_gradients_reach_lights_materials_and_camera(light):
    """eval.py:411-470 / :693-725 optimise the light position and the camera: the lighting table is a differentiable
    input (second sparse pass of pert_phong_bwd), torch carries its gradient back to whichever tensor requires it."""
    import pertrenderer_b200 as pb
    N, H, W, K = 2, 12, 10, 6
    fr, verts, faces, lights, mats, cams, face_colors, _ = _scene(N, H, W, K, 80, light=light, per_batch=(light == "point"), seed=33,
                                                                  shininess=6.0)
    gen = torch.Generator().manual_seed(4)
    grad = torch.randn(N, H, W, K, 3, generator=gen)
    grad[torch.rand(N, H, W, K, generator=gen) < 0.4] = 0.0

    def leaves(dev):
        li, ma, ca = _to(lights, dev), _to(mats, dev), _to(cams, dev)
        t = [getattr(li, "location" if light == "point" else "direction"), li.diffuse_color, li.ambient_color, ma.specular_color,
             ma.shininess, ca.T]
        for x in t:
            x.requires_grad_(True)
        return li, ma, ca, t

    li_o, ma_o, ca_o, t_o = leaves("cpu")
    mesh_o = pb.TriMeshes(verts, faces, face_colors=face_colors)
    col_o = PO.phong_colors_from(mesh_o, fr, li_o, ca_o, ma_o, pb.FaceTexels(face_colors).materialize(fr.pix_to_face))
    (col_o * grad).sum().backward()
    li_c, ma_c, ca_c, t_c = leaves(DEV)
    mesh_c = pb.TriMeshes(verts.to(DEV), faces.to(DEV), face_colors=face_colors.to(DEV))
    fr_c = _frag_to(fr, DEV)
    col_c = pb.phong_shading(mesh_c, fr_c, li_c, ca_c, ma_c, mesh_c.sample_textures(fr_c))
    (col_c * grad.to(DEV)).sum().backward()
    for name, a, b in zip(("light", "diffuse", "ambient", "specular", "shininess", "camera T"), t_c, t_o):
        assert a.grad is not None and rel_err(a.grad.cpu(), b.grad) <= 1e-4, (name, a.grad.cpu(), b.grad)


@pytest.mark.parametrize("n_maps", [1, 2])
def test_uv_textures_match_grid_sample(n_maps):
    """TexturesUV / the legacy Textures(verts_uvs, faces_uvs, maps) of eval.py:750-756: bilinear tap of the map at the
    interpolated corner UVs.  Oracle: UVTexels.materialize = pytorch3d 0.4.0's sample_textures restated on
    torch.nn.functional.grid_sample (align_corners, border padding, flipped map).  UVs partly outside [0,1]: border."""
    import pertrenderer_b200 as pb
    N, H, W, K = 2, 10, 9, 5
    fr, verts, faces, lights, mats, cams, _, _ = _scene(N, H, W, K, 80, seed=17)
    gen = torch.Generator().manual_seed(6)
    maps = torch.rand(n_maps, 7, 11, 3, generator=gen)
    verts_uvs = torch.rand(40, 2, generator=gen) * 1.3 - 0.15
    faces_uvs = torch.randint(0, 40, (faces.shape[0], 3), generator=gen)
    grad = torch.randn(N, H, W, K, 3, generator=gen)
    grad[torch.rand(N, H, W, K, generator=gen) < 0.3] = 0.0
    # (a) sampling alone
    m_o, b_o = maps.clone().requires_grad_(True), fr.bary_coords.clone().requires_grad_(True)
    tex_o = pb.UVTexels(m_o, verts_uvs, faces_uvs).materialize(fr.pix_to_face, b_o)
    (tex_o * grad).sum().backward()
    m_c = maps.to(DEV).requires_grad_(True)
    fr_c = _frag_to(fr, DEV, bary_grad=True)
    tex_c = pb.sample_lazy_textures(pb.UVTexels(m_c, verts_uvs.to(DEV), faces_uvs.to(DEV)), fr_c)
    (tex_c * grad.to(DEV)).sum().backward()
    assert (tex_c.detach().cpu() - tex_o.detach()).abs().max() <= 2e-6
    assert rel_err(m_c.grad.cpu(), m_o.grad) <= 5 * RTOL
    assert rel_err(fr_c.bary_coords.grad.cpu(), b_o.grad) <= 1e-4  # differences of neighbouring texels times (Wm - 1)
    # (b) inside the Phong kernel, through RandomPhongShader with the SoftRas pair
    m_o2, v_o = maps.clone().requires_grad_(True), verts.clone().requires_grad_(True)
    mesh_o = pb.TriMeshes(v_o, faces, uv=(m_o2, verts_uvs, faces_uvs))
    col_o = PO.phong_colors_from(mesh_o, fr, lights, cams, mats, pb.UVTexels(m_o2, verts_uvs, faces_uvs).materialize(fr.pix_to_face, fr.bary_coords))
    (col_o * grad).sum().backward()
    m_c2, v_c = maps.to(DEV).requires_grad_(True), verts.to(DEV).requires_grad_(True)
    mesh_c = pb.TriMeshes(v_c, faces.to(DEV), uv=(m_c2, verts_uvs.to(DEV), faces_uvs.to(DEV)))
    fr_c2 = _frag_to(fr, DEV)
    col_c = pb.phong_shading(mesh_c, fr_c2, _to(lights, DEV), _to(cams, DEV), _to(mats, DEV), mesh_c.sample_textures(fr_c2))
    (col_c * grad.to(DEV)).sum().backward()
    assert (col_c.detach().cpu() - col_o.detach()).abs().max() <= 2e-6
    assert rel_err(m_c2.grad.cpu(), m_o2.grad) <= 5 * RTOL
    assert rel_err(v_c.grad.cpu(), v_o.grad) <= 5 * RTOL
    # (c) the reference's cube: one UV point per side -> a constant colour per face, whatever the sampling details
    strip = torch.tensor([[0.9, 0.1, 0.1], [0.1, 0.7, 0.1], [0.1, 0.2, 0.9]])
    cmap = strip.repeat_interleave(4, dim=0)[None, None].expand(1, 5, 12, 3).contiguous()  # three vertical strips
    vt = torch.tensor([[1.5 / 12 * 12 / 11, 0.5], [5.5 / 11, 0.5], [9.5 / 11, 0.5]])  # centres of the strips (align_corners)
    fuv = (torch.arange(faces.shape[0]) % 3)[:, None].expand(-1, 3).contiguous()
    tex = pb.sample_lazy_textures(pb.UVTexels(cmap.to(DEV), vt.to(DEV), fuv.to(DEV)), _frag_to(fr, DEV)).cpu()
    mask = fr.pix_to_face >= 0
    expect = strip[(fr.pix_to_face.clamp(min=0) % 3)]
    assert (tex[mask] - expect[mask]).abs().max() <= 1e-6


def test_phong_full_size_properties_config2():
    """BASELINE config 2 shapes (8 x 256 x 256, K = 50): size-independent properties of the Phong pass.
    Linearity of backward in grad_colors; padded entries untouched by sparse mode; the sum of the face-table
    gradient equals the sum over entries of b_i * g_p (checked through grad_bary's identity
    sum_i b_i * grad_b_i = g_p . p + g_n . n_raw, i.e. two launches agree on a scalar checksum)."""
    import pertrenderer_b200 as pb
    from pertrenderer_b200 import shading
    N, H, W, K = 8, 256, 256, 50
    fr, _ = pb.synthetic_fragments(N, H, W, K, kind="realistic", n_faces=1280, device=DEV)
    verts, faces = pb.synthetic_mesh(1280, device=DEV)
    p2f = fr.pix_to_face.clamp(max=faces.shape[0] - 1)
    bary = pb.synthetic_bary(p2f)
    mesh = pb.TriMeshes(verts, faces)
    fv, fn = verts[faces].contiguous(), mesh.verts_normals_packed()[faces].contiguous()
    fc = torch.rand(faces.shape[0], 3, device=DEV)
    lighting = shading.pack_lighting(pb.PointLights(location=[[0.0, 2.0, -2.0]], device=DEV), pb.Materials(device=DEV),
                                     pb.ViewCameras(R=torch.eye(3)[None], T=[[0.0, 0.0, 2.7]], device=DEV), N, DEV)
    colors = shading.phong_forward(p2f, bary, fv, fn, None, fc, lighting)
    mask = p2f >= 0
    assert torch.isfinite(colors).all() and (colors[~mask] == 0).all() and (colors[mask] >= 0).all()
    g1 = torch.randn(N, H, W, K, 3, device=DEV) * mask[..., None]
    g2 = torch.randn(N, H, W, K, 3, device=DEV) * mask[..., None]
    a = shading.phong_backward(p2f, bary, fv, fn, None, fc, lighting, g1)
    b = shading.phong_backward(p2f, bary, fv, fn, None, fc, lighting, g2)
    c = shading.phong_backward(p2f, bary, fv, fn, None, fc, lighting, 2.0 * g1 - 0.5 * g2)
    for x, y, z in zip(a, b, c):
        assert rel_err(z, 2.0 * x - 0.5 * y) <= 1e-4
    # Euler identity of the interpolation: sum_i b_i * dL/db_i = <dL/dV, V> + <dL/dN, N> entry by entry, summed
    lhs = (bary * a[1])[mask].double().sum()
    rhs = (a[2].double() * fv.double()).sum() + (a[3].double() * fn.double()).sum()
    assert abs(lhs - rhs) <= 1e-4 * max(abs(rhs), 1.0)


def test_integration_md_phong_stub_runs():
    """The Phong ctypes stub of INTEGRATION.md section 2b is executable as written and reproduces phong_shading."""
    import os
    import re
    import pertrenderer_b200 as pb
    from conftest import ROOT
    from pertrenderer_b200 import _cabi
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    base = [b for b in blocks if "class PerturbedShade" in b][0].replace('"libpertshade.so"', repr(_cabi.LIB_PATH))
    stub = [b for b in blocks if "class PhongColors" in b][0]
    ns = {}
    exec(base, ns)
    exec(stub, ns)
    N, H, W, K = 1, 10, 10, 8
    fr, verts, faces, lights, mats, cams, _, texels = _scene(N, H, W, K, 60, seed=11)
    lights, mats, cams = _to(lights, DEV), _to(mats, DEV), _to(cams, DEV)
    G = torch.randn(N, H, W, K, 3, device=DEV, generator=torch.Generator(device=DEV).manual_seed(3))

    def run(fn):
        v = verts.to(DEV).requires_grad_(True)
        t = texels.to(DEV).requires_grad_(True)
        frc = _frag_to(fr, DEV, bary_grad=True)
        mesh = pb.TriMeshes(v, faces.to(DEV), texels=t)
        col = fn(mesh, frc, t)
        (col * G).sum().backward()
        return col.detach(), v.grad, t.grad, frc.bary_coords.grad

    ours = run(lambda mesh, frc, t: pb.phong_shading(mesh, frc, lights, cams, mats, t))
    theirs = run(lambda mesh, frc, t: ns["PhongColors"].apply(
        mesh.verts_packed()[mesh.faces_packed()], mesh.verts_normals_packed()[mesh.faces_packed()], t.contiguous(),
        frc.bary_coords.contiguous(), frc.pix_to_face, ns["lighting_rows"](lights, mats, cams)))
    for a, b in zip(ours, theirs):
        assert rel_err(a, b) <= 1e-5
