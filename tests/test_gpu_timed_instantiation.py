"""Oracle checks of the kernel instantiations bench.py times (VERDICT r1, weak #1a): BASELINE config 2 at full size with
the default flags -- the sparse-first main pass on 16-pixel tiles with tile blobs for sparse fragments, the fallback pass
for dense / rasterised ones -- and BASELINE config 1 at full size.

With default flags the coverage stage is drawn in law (compound sampler), so what the oracle can check sample by sample is
everything DOWNSTREAM of the kernel's own hit counts: logits, the S_agg perturbed argmax draws (pert_noise_fill with the
tile's global pixel offset materialises exactly the noise the kernel drew in registers), histogram, image, grad_colors, and
the gradients of every logit that can win (never-winning logits get their in-law draw in this mode; the argmax of zi also
collects them through -sum_j grad_zeta_j, so those entries are excluded).  The per-sample flag is compared end to end.
"""

import math
import os
import sys
import types

import pytest
import torch

from conftest import ROOT, elementwise_close, rel_err
from oracle import pert_oracle as O

pytestmark = pytest.mark.gpu

SIGMA, GAMMA, ALPHA, EPS = 1e-3, 1e-2, 1.0, 1e-10
BG = (1.0, 1.0, 1.0)


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _fragments(kind, N, HW, K, S):
    import pertrenderer_b200 as pb
    if kind == "rasterised":
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        import bench
        return bench.rasterised_fragments(types.SimpleNamespace(views=N, image_size=HW, faces_per_pixel=K, nb_samples=S),
                                          torch.device("cuda"))
    return pb.synthetic_fragments(N, HW, HW, K, kind=kind, sigma=SIGMA, seed=0, device="cuda")


def _downstream_oracle(sub, counts, rsum, V, S, G):
    """The oracle from given hit counts on: (a_s, image, grads dict, zeta)."""
    zeta, prob, aux = O.logits_from_counts(sub["pix_to_face"], sub["zbuf"], counts.float(), S, 1.0, 100.0, GAMMA, ALPHA, EPS)
    hist, a_s, a_0 = O.argmax_shard(zeta, V, GAMMA)
    image, weights = O.blend_from_hist(hist, S, prob, sub["colors"], BG)
    grad_w = O._grad_weights(sub["colors"], G, O._as_background(BG))
    packed = O.argmax_score_sums(grad_w, a_s, a_0, V)
    gr = O.shade_backward_from_sums(prob, weights, dict(aux, bg=BG), G, sub["zbuf"], sub["colors"], SIGMA, GAMMA, ALPHA, EPS,
                                    S, S, rsum, packed)
    return a_s, image, gr, zeta


def _check_tiles(kind, N, HW, K, S, n_tiles, seed):
    from pertrenderer_b200 import ops
    fr, col = _fragments(kind, N, HW, K, S)
    dev = fr.pix_to_face.device
    P = N * HW * HW
    G = torch.randn((N, HW, HW, 4), device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    seed_r, seed_a = 0x5EED0001 + seed, 0x5EED0002 + seed
    pr = ops.ShadeProblem(pix_to_face=fr.pix_to_face, zbuf=fr.zbuf, dists=fr.dists, colors=col, znear=1.0, zfar=100.0,
                          background=BG, sigma=SIGMA, gamma=GAMMA, alpha=ALPHA, eps=EPS, S_rast=S, S_agg=S,
                          seed_rast=seed_r, seed_agg=seed_a)
    image, saved = ops.shade_forward(pr)
    assert saved.blob is not None and saved.worklist is not None  # the production path: sparse-first tiles, tile blobs
    gd, gz, gc, scal = ops.shade_backward(pr, saved, G)
    torch.cuda.synchronize()
    n_over = int(saved.worklist[0].item())
    tp = 16
    tiles = torch.randperm(P // tp, generator=torch.Generator().manual_seed(seed))[:n_tiles].sort().values
    pix = (tiles[:, None] * tp + torch.arange(tp)[None, :]).reshape(-1).to(dev)

    def rows(t, tail):
        return t.reshape((P,) + tail)[pix].reshape((1, n_tiles, tp) + tail).cpu()

    sub = dict(pix_to_face=rows(fr.pix_to_face, (K,)), zbuf=rows(fr.zbuf, (K,)), colors=rows(col, (K, 3)))
    mask = sub["pix_to_face"] >= 0
    counts = rows(saved.counts.to(torch.int32) & 0xFFFF, (K,)) * mask
    rsum = rows(saved.rsum, (K,)) * mask
    V = torch.cat([ops.noise_fill(seed_a, 1, (1, 1, tp, K), S, dev, pixel_offset=int(t) * tp) for t in tiles], dim=2).cpu()
    a_s, img_o, gr, zeta = _downstream_oracle(sub, counts, rsum, V, S, rows(G, (4,)))
    # indices bit-exact, image to rounding, colour gradients (weights * G) to rounding
    win = rows(saved.winners_full(), (S,)).long().permute(3, 0, 1, 2)
    assert torch.equal(win, a_s), f"{kind}: argmax winners differ from the oracle"
    assert (rows(image, (4,)) - img_o).abs().max().item() <= 2e-6
    assert rel_err(rows(gc, (K, 3)), gr["colors"]) <= 1e-6
    # gradients of the logits that can win, away from the argmax of zi
    zmax = zeta.max(-1, keepdim=True).values
    cut = 2.0 * GAMMA * 5.66 * 1.0001 + 4e-7 * zmax.abs().clamp(min=1.0)
    live = (torch.isfinite(zeta) & (zeta >= zmax - cut + 1e-6))[..., :-1]
    zi = torch.where(mask, (100.0 - sub["zbuf"]) / 99.0, torch.zeros_like(sub["zbuf"]))
    sel = mask & live & (zi < zi.max(-1, keepdim=True).values)
    assert sel.sum().item() > n_tiles  # the comparison is not vacuous
    for name, got, ref in (("grad_zbuf", rows(gz, (K,)), gr["zbuf"]), ("grad_dists", rows(gd, (K,)), gr["dists"])):
        scale = ref.abs().max().item()
        err = (got - ref)[sel].abs().max().item()
        assert err <= 1e-5 * scale, (kind, name, err, scale)
        assert elementwise_close(got[sel], ref[sel], 1e-5), (kind, name)
    assert (rows(gz, (K,))[~mask] == 0).all() and (rows(gd, (K,))[~mask] == 0).all()
    return n_over, P // tp


@pytest.mark.parametrize("kind", ["realistic", "dense", "rasterised"])
def test_config2_default_flags_tiles_match_the_oracle(kind):
    """BASELINE config 2 (8 x 256^2, K = 50, S = 64), default flags, 64 random 16-pixel tiles against the oracle."""
    n_over, n_tiles = _check_tiles(kind, 8, 256, 50, 64, 64, seed=3)
    if kind == "realistic":
        assert n_over < 0.02 * n_tiles  # the main pass (16-pixel tiles, tile blobs) is what ran
    else:
        assert n_over > 0.3 * n_tiles   # the fallback pass is what ran


def test_config2_per_sample_flag_tiles_match_the_oracle_end_to_end():
    """Same job with PERT_F_PER_SAMPLE_NOISE: coverage counts, winners, image and ALL gradients of 48 random tiles."""
    from pertrenderer_b200 import _cabi, ops
    import pertrenderer_b200 as pb
    N, HW, K, S, tp, n_tiles = 8, 256, 50, 64, 16, 48
    fr, col = pb.synthetic_fragments(N, HW, HW, K, kind="realistic", sigma=SIGMA, seed=0, device="cuda")
    dev = fr.pix_to_face.device
    P = N * HW * HW
    G = torch.randn((N, HW, HW, 4), device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    seed_r, seed_a = 77, 78
    pr = ops.ShadeProblem(pix_to_face=fr.pix_to_face, zbuf=fr.zbuf, dists=fr.dists, colors=col, znear=1.0, zfar=100.0,
                          background=BG, sigma=SIGMA, gamma=GAMMA, alpha=ALPHA, eps=EPS, S_rast=S, S_agg=S,
                          seed_rast=seed_r, seed_agg=seed_a, flags=_cabi.F_PER_SAMPLE_NOISE)
    image, saved = ops.shade_forward(pr)
    gd, gz, gc, scal = ops.shade_backward(pr, saved, G)
    tiles = torch.randperm(P // tp, generator=torch.Generator().manual_seed(9))[:n_tiles].sort().values
    pix = (tiles[:, None] * tp + torch.arange(tp)[None, :]).reshape(-1).to(dev)

    def rows(t, tail):
        return t.reshape((P,) + tail)[pix].reshape((1, n_tiles, tp) + tail).cpu()

    U = torch.cat([ops.noise_fill(seed_r, 0, (1, 1, tp, K), S, dev, pixel_offset=int(t) * tp) for t in tiles], dim=2).cpu()
    V = torch.cat([ops.noise_fill(seed_a, 1, (1, 1, tp, K), S, dev, pixel_offset=int(t) * tp) for t in tiles], dim=2).cpu()
    st, gr = O.shade_fwd_bwd(rows(fr.pix_to_face, (K,)), rows(fr.zbuf, (K,)), rows(fr.dists, (K,)), rows(col, (K, 3)), BG, 1.0,
                             100.0, SIGMA, GAMMA, ALPHA, EPS, U, V, rows(G, (4,)))
    mask = rows(fr.pix_to_face, (K,)) >= 0
    assert torch.equal((rows(saved.counts.to(torch.int32) & 0xFFFF, (K,)))[mask], st.counts[mask])
    assert torch.equal(rows(saved.winners_full(), (S,)).long().permute(3, 0, 1, 2), st.a_s)
    assert (rows(image, (4,)) - st.image).abs().max().item() <= 2e-6
    for name, got, ref in (("colors", rows(gc, (K, 3)), gr["colors"]), ("zbuf", rows(gz, (K,)), gr["zbuf"]),
                           ("dists", rows(gd, (K,)), gr["dists"])):
        assert rel_err(got, ref) <= 1e-5, name
        assert elementwise_close(got, ref, 1e-5), name


def test_config1_full_size_matches_the_oracle():
    """BASELINE config 1 at full size (1 x 64^2, K = 50, S = 16): per-sample flag end to end against the oracle on the
    whole image; default flags through the downstream oracle on every pixel."""
    from pertrenderer_b200 import _cabi, ops
    import pertrenderer_b200 as pb
    N, HW, K, S = 1, 64, 50, 16
    for kind in ("realistic", "dense"):
        fr, col = pb.synthetic_fragments(N, HW, HW, K, kind=kind, sigma=SIGMA, seed=4, device="cuda")
        dev = fr.pix_to_face.device
        G = torch.randn((N, HW, HW, 4), device=dev, generator=torch.Generator(device=dev).manual_seed(2))
        cpu = dict(pix_to_face=fr.pix_to_face.cpu(), zbuf=fr.zbuf.cpu(), dists=fr.dists.cpu(), colors=col.cpu())
        mask = cpu["pix_to_face"] >= 0

        def problem(flags):
            return ops.ShadeProblem(pix_to_face=fr.pix_to_face, zbuf=fr.zbuf, dists=fr.dists, colors=col, znear=1.0, zfar=100.0,
                                    background=BG, sigma=SIGMA, gamma=GAMMA, alpha=ALPHA, eps=EPS, S_rast=S, S_agg=S,
                                    seed_rast=5, seed_agg=6, flags=flags)
        U = ops.noise_fill(5, 0, (N, HW, HW, K), S, dev).cpu()
        V = ops.noise_fill(6, 1, (N, HW, HW, K), S, dev).cpu()
        # per-sample path, everything
        pr = problem(_cabi.F_PER_SAMPLE_NOISE)
        image, saved = ops.shade_forward(pr)
        gd, gz, gc, scal = ops.shade_backward(pr, saved, G)
        st, gr = O.shade_fwd_bwd(cpu["pix_to_face"], cpu["zbuf"], cpu["dists"], cpu["colors"], BG, 1.0, 100.0, SIGMA, GAMMA,
                                 ALPHA, EPS, U, V, G.cpu())
        assert torch.equal((saved.counts.to(torch.int32) & 0xFFFF).cpu()[mask], st.counts[mask])
        assert torch.equal(saved.winners_full().cpu().long().permute(3, 0, 1, 2), st.a_s)
        assert (image.cpu() - st.image).abs().max().item() <= 2e-6
        for name, got, ref in (("colors", gc, gr["colors"]), ("zbuf", gz, gr["zbuf"]), ("dists", gd, gr["dists"])):
            assert rel_err(got.cpu(), ref) <= 1e-5, (kind, name)
            assert elementwise_close(got.cpu(), ref, 1e-5), (kind, name)
        for i, k in enumerate(("sigma", "gamma", "alpha")):
            assert abs(scal[i].item() - gr[k].item()) <= 2e-5 * abs(gr[k].item()) + 1e-7, (kind, k)
        # default flags: downstream of the kernel's own counts
        pr = problem(0)
        image, saved = ops.shade_forward(pr)
        gd, gz, gc, scal = ops.shade_backward(pr, saved, G)
        counts = (saved.counts.to(torch.int32) & 0xFFFF).cpu() * mask
        a_s, img_o, gr2, zeta = _downstream_oracle(cpu, counts, saved.rsum.cpu() * mask, V, S, G.cpu())
        assert torch.equal(saved.winners_full().cpu().long().permute(3, 0, 1, 2), a_s)
        assert (image.cpu() - img_o).abs().max().item() <= 2e-6
        assert rel_err(gc.cpu(), gr2["colors"]) <= 1e-6
        assert torch.isfinite(gd).all() and torch.isfinite(gz).all() and torch.isfinite(scal).all()


def test_graph_captured_step_matches_eager_and_redraws_noise():
    """Small-problem path: forward + backward captured in ONE CUDA graph with device-side seeds (pert_problem.seed_device,
    pert_seed_advance).  A replay equals the eager run with the effective seeds (seed ^ device value) bit for bit; the next
    replay draws other noise; inputs written into the static buffers between replays are picked up."""
    import pertrenderer_b200 as pb
    from pertrenderer_b200 import ops
    N, HW, K, S = 1, 64, 50, 16
    fr, col = pb.synthetic_fragments(N, HW, HW, K, kind="realistic", sigma=SIGMA, seed=3, device="cuda")
    dev = fr.pix_to_face.device
    G = torch.randn((N, HW, HW, 4), device=dev, generator=torch.Generator(device=dev).manual_seed(2))
    zbuf, dists, colors = fr.zbuf.clone(), fr.dists.clone(), col.clone()
    step = ops.GraphedShadeStep(fr.pix_to_face, zbuf, dists, colors, G, sigma=SIGMA, gamma=GAMMA, alpha=ALPHA, eps=EPS,
                                S_rast=S, S_agg=S, background=BG, seed=1234)
    seeds_before = step.seed_device.clone()
    out1 = [t.clone() for t in step.replay()]
    torch.cuda.synchronize()
    seeds_used = step.seed_device.clone()  # advance runs first inside the graph: these are the seeds replay 1 drew with
    assert not torch.equal(seeds_before, seeds_used)

    def eager(seeds):
        pr = ops.ShadeProblem(pix_to_face=fr.pix_to_face, zbuf=zbuf, dists=dists, colors=colors, znear=1.0, zfar=100.0,
                              background=BG, sigma=SIGMA, gamma=GAMMA, alpha=ALPHA, eps=EPS, S_rast=S, S_agg=S,
                              seed_rast=int(seeds[0].item()), seed_agg=int(seeds[1].item()))
        image, saved = ops.shade_forward(pr)
        return (image,) + ops.shade_backward(pr, saved, G)
    ref1 = eager(seeds_used)
    for a, b in zip(out1, ref1):
        assert torch.equal(a, b)
    out2 = [t.clone() for t in step.replay()]
    assert not torch.equal(out1[0], out2[0])  # fresh noise
    for a, b in zip(out2, eager(step.seed_device)):
        assert torch.equal(a, b)
    dists.mul_(0.5)  # new inputs through the static buffers
    out3 = [t.clone() for t in step.replay()]
    for a, b in zip(out3, eager(step.seed_device)):
        assert torch.equal(a, b)


def test_in_place_edit_of_an_input_before_backward_is_detected():
    """The backward kernels re-read dists / zbuf / colours: the autograd Function registers them, so an in-place edit
    between forward and backward raises instead of changing the gradients silently."""
    import pertrenderer_b200 as pb
    fr, col = pb.synthetic_fragments(1, 8, 8, 6, kind="dense", device="cuda")
    d = fr.dists.clone().requires_grad_(True)
    dd = d * 1.0  # non-leaf, so that an in-place op is legal
    img = pb.smooth_rgb_blend(col, pb.Fragments(fr.pix_to_face, fr.zbuf, None, dd), pb.GaussianRast(nb_samples=8, sigma=1e-3),
                              pb.GaussianAgg(nb_samples=8, gamma=1e-2), pb.BlendParams())
    dd.mul_(2.0)
    with pytest.raises(RuntimeError, match="modified by an inplace operation"):
        img.sum().backward()


def test_atlas_textured_mesh_through_the_fused_shader():
    """TexturesAtlas meshes (eval.py:216-238): TriMeshes(atlas=...) -> sample_textures -> the fused pair; the gradient reaches
    the atlas cells that were hit, and the image equals the one rendered from the materialised texel tensor."""
    import pertrenderer_b200 as pb
    dev = "cuda"
    N, HW, K, F, R = 2, 16, 8, 20, 4
    fr, _ = pb.synthetic_fragments(N, HW, HW, K, kind="realistic", sigma=SIGMA, n_faces=F, seed=2, device=dev)
    p2f = fr.pix_to_face.clamp(max=F - 1)
    bary = pb.synthetic_bary(p2f, seed=1)
    frag = pb.Fragments(p2f, fr.zbuf, bary, fr.dists)
    verts, faces = pb.synthetic_mesh(F, device=dev)
    atlas = torch.rand((F, R, R, 3), device=dev).requires_grad_(True)
    mesh = pb.TriMeshes(verts, faces, atlas=atlas)
    shader = pb.RandomSimpleShader(device=dev, cameras=pb.DepthCameras(n=N, device=dev),
                                   smoothrast=pb.GaussianRast(nb_samples=16, sigma=SIGMA),
                                   smoothagg=pb.GaussianAgg(nb_samples=16, gamma=GAMMA), blend_params=pb.BlendParams())
    torch.manual_seed(5)
    img = shader(frag, mesh)
    img[..., :3].sum().backward()
    assert atlas.grad is not None and torch.isfinite(atlas.grad).all() and atlas.grad.abs().sum() > 0
    tex = pb.AtlasTexels(atlas.detach()).materialize(p2f, bary)
    torch.manual_seed(5)
    img2 = shader(frag, pb.TexelMeshes(tex))
    assert torch.equal(img.detach(), img2.detach())
