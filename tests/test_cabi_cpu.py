"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/pertshade.h
declares; the reference-facing host classes keep the reference's surface and error behaviour."""

import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT

import pertrenderer_b200 as pb
from pertrenderer_b200 import _cabi, ops


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "pertshade.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pert_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _cabi.load()
    declared = _declared_functions()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in pertshade.h but not exported"
    assert sorted(_cabi.EXPORTS) == declared
    assert lib.pert_version() == _cabi.ABI_VERSION
    assert lib.pert_strerror(0) == b"ok"
    assert b"NULL" in lib.pert_strerror(-1)


def test_struct_layout_matches_header():
    """ctypes mirror of pert_problem: field order / sizes as in the header (LP64)."""
    assert ctypes.sizeof(_cabi.PertProblem) == 200
    assert _cabi.PertProblem.pix_to_face.offset == 112
    assert _cabi.PertProblem.seed_rast.offset == 80
    assert _cabi.PertProblem.seed_device.offset == 192


def test_argument_validation_without_gpu():
    """Validation happens before any launch, so it is testable on a CPU-only box."""
    lib = _cabi.load()
    p = _cabi.PertProblem()
    assert lib.pert_shade_fwd(None, None, None, None, None, None, None, None, None, None) == -1
    p.N, p.H, p.W, p.K = 1, 2, 2, 0
    assert lib.pert_shade_fwd(p, None, None, None, None, None, None, None, None, None) == -2  # bad shape
    p.K = 5000
    assert lib.pert_shade_fwd(p, None, None, None, None, None, None, None, None, None) == -3  # unsupported K
    p.K = 4
    p.S_rast = p.S_agg = 8
    p.depth_len = 1
    p.sigma, p.gamma, p.alpha = 1e-3, 0.0, 1.0
    assert lib.pert_shade_fwd(p, None, None, None, None, None, None, None, None, None) == -7  # gamma must be > 0
    p.gamma = 1e-2
    p.s_rast_begin, p.s_rast_end, p.s_agg_begin, p.s_agg_end = 2, 8, 0, 8
    assert lib.pert_shade_fwd(p, None, None, None, None, None, None, None, None, None) == -5  # shard begin % 4
    p.s_rast_begin = 0
    assert lib.pert_shade_fwd(p, None, None, None, None, None, None, None, None, None) == -1  # null inputs
    assert lib.pert_winner_bytes(50) == 1 and lib.pert_winner_bytes(255) == 1 and lib.pert_winner_bytes(256) == 2
    p.N, p.H, p.W, p.K = 8, 256, 256, 50
    assert lib.pert_num_tiles(p) == 8 * 256 * 256 // 8  # finest geometry of this problem: 8-pixel tiles
    assert lib.pert_rast_fwd(None, 1, 1, 1, 0, 1, 1.0, 0, 0, None, 0, None, None, None) == -1
    assert lib.pert_noise_fill(0, 0, 4, 4, 0, 4, 0, None, None) == -1


def test_cpu_tensors_fail_loudly():
    """No CPU fallback: the product path refuses CPU tensors instead of computing something."""
    frag = pb.Fragments(torch.zeros(1, 2, 2, 3, dtype=torch.int64), torch.ones(1, 2, 2, 3), None,
                        torch.zeros(1, 2, 2, 3))
    colors = torch.rand(1, 2, 2, 3, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        pb.smooth_rgb_blend(colors, frag, pb.GaussianRast(), pb.GaussianAgg(), pb.BlendParams())
    with pytest.raises(RuntimeError, match="CUDA"):
        pb.GaussianRast().rasterize(torch.zeros(1, 2, 2, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        pb.randomArgmax.apply(torch.zeros(1, 2, 2, 4), 4, torch.tensor(1e-2))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_cabi.PertLibraryError, match="no CPU or PyTorch fallback"):
        _cabi.load()


def test_operator_surface_matches_reference():
    """constructor defaults, attributes and setters."""
    r = pb.GaussianRast()
    assert r.nb_samples == 16 and abs(r.sigma.item() - 2e-4) < 1e-9
    assert r.sigma.requires_grad and r.sigma.dim() == 0 and r.sigma.device.type == "cpu"
    assert not isinstance(r.sigma, torch.nn.Parameter)
    old = r.sigma
    r.update_smoothing(5e-4)
    assert r.sigma is not old and abs(r.sigma.item() - 5e-4) < 1e-9
    r.update_nb_samples(64)
    assert r.nb_samples == 64
    a = pb.GaussianAgg()
    assert (a.nb_samples, a.eps, a.fixed_noise) == (16, 1e-10, False)
    assert abs(a.gamma.item() - 4e-2) < 1e-9 and a.alpha.item() == 1.0
    a.update_smoothing(gamma=1e-3, alpha=2.0)
    assert abs(a.gamma.item() - 1e-3) < 1e-9 and a.alpha.item() == 2.0
    a.update_nb_samples(8)
    assert a.nb_samples == 8
    sh = pb.RandomSimpleShader(smoothrast=r, smoothagg=a, cameras=pb.DepthCameras())
    assert sh.get_nb_samples() == 8
    s, g, al = sh.get_smoothing()
    assert s is r.sigma and g is a.gamma and al is a.alpha
    sh.update_smoothing(sigma=1e-3, gamma=1e-2, alpha=1.0)
    sh.update_nb_samples(32)
    assert r.nb_samples == 32 and a.nb_samples == 32 and abs(r.sigma.item() - 1e-3) < 1e-9
    assert sh.to("cpu") is None  # reference quirk B7
    sh.cameras = None
    with pytest.raises(ValueError):
        sh.forward(None, None)


def test_unsupported_noise_type_raises():
    with pytest.raises(ValueError, match="not implemented"):
        pb.randomHeaviside.apply(torch.zeros(1, 1, 1, 2), 4, torch.tensor(1e-3), "logistic")
    with pytest.raises(ValueError, match="not implemented"):
        pb.randomArgmax.apply(torch.zeros(1, 1, 1, 2), 4, torch.tensor(1e-3), "logistic")
    # gumbel / uniform exist forward-only (as in the reference): on CPU tensors they hit the CUDA-only guard
    with pytest.raises(RuntimeError, match="CUDA"):
        pb.randomArgmax.apply(torch.zeros(1, 1, 1, 2), 4, torch.tensor(1e-3), "gumbel")
    lib = _cabi.load()
    assert lib.pert_argmax_bwd(None, None, None, 1, 2, 4, 0, 4, 1e-2, 0, 0, None, _cabi.F_UNIFORM, None, None, None, None) == -3


def test_soft_operators_cpu():
    """SoftRast / SoftAgg (the shaders' default arguments) are plain torch and run anywhere."""
    d = torch.tensor([[[[-1e-4, 0.0, 2e-4]]]])
    p = pb.SoftRast(sigma=1e-4).rasterize(d)
    assert torch.allclose(p, torch.sigmoid(-d / 1e-4))
    agg = pb.SoftAgg(gamma=1e-2)
    z = torch.tensor([[[[5.0, 6.0, -1.0]]]])
    mask = torch.tensor([[[[True, True, False]]]])
    prob = torch.tensor([[[[1.0, 0.5, 0.0]]]], requires_grad=True)
    w = agg.aggregate(z, 100.0, 1.0, prob, mask)
    assert w.shape == (1, 1, 1, 4) and abs(w.sum().item() - 1) < 1e-6 and w[0, 0, 0, 2] == 0
    w[..., 0].sum().backward()
    assert torch.isfinite(prob.grad).all() and prob.grad[0, 0, 0, 2] == 0


def test_seed_follows_torch_generator():
    torch.manual_seed(123)
    a = ops.draw_seed()
    b = ops.draw_seed()
    torch.manual_seed(123)
    assert ops.draw_seed() == a and ops.draw_seed() == b and a != b


def test_device_seed_context_nests_and_restores():
    """ops.device_seeds: the seed pair a caller hands to every fused op inside a CUDA-graph capture (pert_problem.seed_device)
    is thread-local state that nests and is restored on exit, also when the block raises."""
    import threading
    from pertrenderer_b200 import ops
    assert ops.current_seed_device() is None
    a, b = torch.zeros(2, dtype=torch.int64), torch.ones(2, dtype=torch.int64)
    with ops.device_seeds(a):
        assert ops.current_seed_device() is a
        with ops.device_seeds(b):
            assert ops.current_seed_device() is b
        assert ops.current_seed_device() is a
        seen = []
        t = threading.Thread(target=lambda: seen.append(ops.current_seed_device()))
        t.start()
        t.join()
        assert seen == [None]  # another thread's ops are not affected
        with pytest.raises(RuntimeError):
            with ops.device_seeds(b):
                raise RuntimeError("boom")
        assert ops.current_seed_device() is a
        with pytest.raises(RuntimeError):  # stand-alone operators would replay frozen noise inside a graph: refused
            ops.refuse_device_seeds("randomArgmax")
    assert ops.current_seed_device() is None
    ops.refuse_device_seeds("randomArgmax")


def test_synthetic_fragments_contract():
    for kind in ("dense", "realistic"):
        fr, col = pb.synthetic_fragments(2, 16, 16, 10, kind=kind, device="cpu", seed=3)
        valid = fr.pix_to_face >= 0
        assert fr.pix_to_face.dtype == torch.int64 and fr.zbuf.dtype == torch.float32
        assert (fr.zbuf[~valid] == -1).all() and (fr.dists[~valid] == -1).all() and (col[~valid] == 0).all()
        # padding last, depth ascending among valid entries
        assert (valid[..., 1:] <= valid[..., :-1]).all()
        dz = fr.zbuf[..., 1:] - fr.zbuf[..., :-1]
        assert (dz[valid[..., 1:]] >= 0).all()
    assert valid.float().mean() < 0.5


def test_reference_package_exports_are_present():
    """Every name randomras/__init__.py:1-3 exports exists here with the reference's constructor arguments."""
    for name in ("SimpleShader", "RandomSimpleShader", "CauchyAgg", "GaussianAgg", "SoftAgg", "SoftRast", "ArctanRast",
                 "AffineRast", "GaussianRast"):
        assert hasattr(pb, name), name
    assert pb.ArctanRast(nb_samples=8, sigma=1e-3).nb_samples == 8
    assert pb.CauchyAgg(nb_samples=8, gamma=1e-2, alpha=1.0, eps=1e-10, fixed_noise=True).fixed_noise
    d = torch.tensor([[[[-1e-3, 0.0, 2e-4, 1e-3]]]])
    assert torch.allclose(pb.AffineRast(sigma=1e-3).rasterize(d), torch.tensor([[[[1.0, 0.5, 0.3, 0.0]]]]))
    assert torch.equal(pb.HardRast().rasterize(d), torch.tensor([[[[1.0, 1.0, 0.0, 0.0]]]]))
    frag = pb.Fragments(torch.tensor([[[[2, -1], [-1, -1]]]]), torch.zeros(1, 1, 2, 2), None, torch.zeros(1, 1, 2, 2))
    tex = torch.tensor([[[[[0.1, 0.2, 0.3], [0.0, 0.0, 0.0]], [[0.0, 0.0, 0.0], [0.0, 0.0, 0.0]]]]])
    img = pb.SimpleShader(blend_params=pb.BlendParams(background_color=(1.0, 0.5, 0.0)))(frag, pb.TexelMeshes(tex))
    assert torch.allclose(img, torch.tensor([[[[0.1, 0.2, 0.3, 1.0], [1.0, 0.5, 0.0, 0.0]]]]))


def test_atlas_texels_known_answers():
    """AtlasTexels (pytorch3d TexturesAtlas.sample_textures restated; eval.py:216-238): R = 1 is a per-face colour; the cell
    of a point is floor(bary * R) below the grid's diagonal and reflected above it; padded entries are zero; the gradient
    is the scatter of the texel gradients into the cells that were hit."""
    import torch
    from pertrenderer_b200.structures import AtlasTexels, Fragments, TriMeshes
    F, R = 3, 4
    atlas = torch.arange(F * R * R * 3, dtype=torch.float32).reshape(F, R, R, 3).requires_grad_(True)
    p2f = torch.tensor([[[[0, 1, -1], [2, 2, -1]]]])  # (1,1,2,3)
    bary = torch.tensor([[[[[0.1, 0.1, 0.8], [0.60, 0.30, 0.10], [0.3, 0.3, 0.4]],
                           [[0.49, 0.49, 0.02], [0.0, 0.0, 1.0], [0.5, 0.5, 0.0]]]]])
    tex = AtlasTexels(atlas).materialize(p2f, bary)
    assert tex.shape == (1, 1, 2, 3, 3)
    a = atlas.detach()
    assert torch.equal(tex[0, 0, 0, 0], a[0, 0, 0])           # (0.1, 0.1) * 4 -> cell (0, 0), below the diagonal
    # (0.6, 0.3) * 4 = (2.4, 1.2): cell (2, 1), 3.6 - 3 = 0.6 <= 1 -> below: atlas[face, w_y = 1, w_x = 2]
    assert torch.equal(tex[0, 0, 0, 1], a[1, 1, 2])
    assert (tex[0, 0, 0, 2] == 0).all() and (tex[0, 0, 1, 2] == 0).all()  # padding
    # (0.49, 0.49) * 4 = (1.96, 1.96): cell (1, 1), 3.92 - 2 = 1.92 > 1 -> above: reflected to (R-1-1, R-1-1) = (2, 2)
    assert torch.equal(tex[0, 0, 1, 0], a[2, 2, 2])
    assert torch.equal(tex[0, 0, 1, 1], a[2, 0, 0])
    g = torch.ones_like(tex)
    tex.backward(g)
    hit = atlas.grad.sum(-1) > 0
    assert hit.sum().item() == 4 and atlas.grad[2, 2, 2, 0].item() == 1.0 and atlas.grad.sum().item() == 12.0
    # R = 1: the atlas is a per-face colour table
    one = torch.rand(F, 1, 1, 3)
    t1 = AtlasTexels(one).materialize(p2f, bary)
    assert torch.equal(t1[0, 0, 0, 1], one[1, 0, 0]) and torch.equal(t1[0, 0, 1, 0], one[2, 0, 0])
    # through the mesh container, extended to a batch of poses (atlas repeated per mesh like the packed faces)
    verts, faces = torch.zeros(4, 3), torch.tensor([[0, 1, 2], [0, 2, 3], [0, 1, 3]])
    m = TriMeshes(verts, faces, atlas=atlas.detach()).extend(2)
    p2 = torch.tensor([[[[0, -1]]], [[[4, -1]]]])  # second image: packed face 3 + 1
    b2 = torch.tensor([[[[[0.1, 0.1, 0.8], [0, 0, 0]]]], [[[[0.1, 0.1, 0.8], [0, 0, 0]]]]], dtype=torch.float32)
    t2 = m.sample_textures(Fragments(p2, None, b2, None))
    assert torch.equal(t2[0, 0, 0, 0], a[0, 0, 0]) and torch.equal(t2[1, 0, 0, 0], a[1, 0, 0])
