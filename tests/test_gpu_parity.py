"""Parity of the CUDA path (through the C ABI) with the CPU oracle and the reference's golden
vectors.  Tolerances (north_star): images and gradients within 1e-5 relative in fp32 when fed the
reference's exact noise; indices (hit counts, argmax winners) bit-exact."""

import pytest
import torch

from conftest import elementwise_close, load_golden, rel_err
from oracle import pert_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-5


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _scalars_close(got, ref, what):
    assert abs(got - ref) <= 2e-5 * max(abs(ref), 1e-12) + 1e-8, (what, got, ref)


def test_explicit_noise_matches_reference_golden(shade_case):
    """Tier 2: the reference's own noise tensors -> reference's outputs."""
    from gpu_util import problem_from_case, run_cuda
    g = shade_case
    out = run_cuda(problem_from_case(g), g["grad_image"])
    mask = g["pix_to_face"] >= 0
    S_a = int(g["S_a"])
    # indices: bit-exact
    assert torch.equal(out["counts"][mask], g["counts"][mask])
    assert torch.equal(out["winners"].permute(3, 0, 1, 2).long(), g["a_s"])
    assert torch.equal(out["hist"].float() / S_a, g["weights"])
    # image and gradients
    assert rel_err(out["image"], g["image"]) <= RTOL
    assert (out["image"] - g["image"]).abs().max() <= 1e-6
    assert rel_err(out["grad_colors"], g["grad_colors"]) <= RTOL
    assert rel_err(out["grad_dists"], g["grad_dists"]) <= RTOL
    assert rel_err(out["grad_zbuf"], g["grad_zbuf"]) <= RTOL
    # ... and entry by entry (north_star: 1e-5 relative on images and gradients), with the floor of conftest.elementwise_close
    for k in ("image", "grad_colors", "grad_dists", "grad_zbuf"):
        assert elementwise_close(out[k], g[k], RTOL), k
    _scalars_close(out["scalars"][0].item(), g["grad_sigma"], "sigma")
    _scalars_close(out["scalars"][1].item(), g["grad_gamma"], "gamma")
    _scalars_close(out["scalars"][2].item(), g["grad_alpha"], "alpha")
    # padded entries: exactly zero gradients (Appendix A.4)
    assert (out["grad_dists"][~mask] == 0).all() and (out["grad_zbuf"][~mask] == 0).all()


def test_public_api_autograd_matches_golden(shade_case):
    """Same comparison through smooth_rgb_blend + autograd (what MeshRenderer calls)."""
    import pertrenderer_b200 as pb
    g = shade_case
    dev = "cuda"
    rast = pb.GaussianRast(nb_samples=int(g["S_r"]), sigma=float(g["sigma"]))
    agg = pb.GaussianAgg(nb_samples=int(g["S_a"]), gamma=float(g["gamma"]), alpha=float(g["alpha"]))
    d = g["dists"].to(dev).requires_grad_(True)
    z = g["zbuf"].to(dev).requires_grad_(True)
    c = g["colors"].to(dev).requires_grad_(True)
    frag = pb.Fragments(g["pix_to_face"].to(dev), z, None, d)
    blend = pb.BlendParams(background_color=tuple(g["background"].tolist()))
    with pb.explicit_noise(g["U"].to(dev), g["V"].to(dev)):
        img = pb.smooth_rgb_blend(c, frag, rast, agg, blend, znear=g["znear_t"].to(dev), zfar=g["zfar_t"].to(dev))
        (img * g["grad_image"].to(dev)).sum().backward()
    assert rel_err(img.detach().cpu(), g["image"]) <= RTOL
    assert rel_err(d.grad.cpu(), g["grad_dists"]) <= RTOL
    assert rel_err(z.grad.cpu(), g["grad_zbuf"]) <= RTOL
    assert rel_err(c.grad.cpu(), g["grad_colors"]) <= RTOL
    assert rast.sigma.grad.device.type == "cpu" and agg.gamma.grad.dim() == 0
    _scalars_close(rast.sigma.grad.item(), g["grad_sigma"], "sigma")
    _scalars_close(agg.gamma.grad.item(), g["grad_gamma"], "gamma")
    _scalars_close(agg.alpha.grad.item(), g["grad_alpha"], "alpha")


def test_unfused_operator_composition_matches_golden(shade_case):
    """GaussianRast.rasterize and GaussianAgg.aggregate used one by one (mixed-pair path): the
    stand-alone kernels + torch glue reproduce the same golden outputs."""
    import pertrenderer_b200 as pb
    g = shade_case
    dev = "cuda"
    rast = pb.GaussianRast(nb_samples=int(g["S_r"]), sigma=float(g["sigma"]))
    agg = pb.GaussianAgg(nb_samples=int(g["S_a"]), gamma=float(g["gamma"]), alpha=float(g["alpha"]))
    d = g["dists"].to(dev).requires_grad_(True)
    z = g["zbuf"].to(dev).requires_grad_(True)
    mask = (g["pix_to_face"] >= 0).to(dev)
    with pb.explicit_noise(g["U"].to(dev), g["V"].to(dev)):
        prob = rast.rasterize(d) * mask
        w = agg.aggregate(z, g["zfar_t"].to(dev), g["znear_t"].to(dev), prob, mask)
    assert torch.equal(prob.detach().cpu(), g["prob"])
    assert torch.equal(w.detach().cpu(), g["weights"])
    colors = g["colors"].to(dev)
    bg = g["background"].to(dev)
    rgb = (w[..., :-1, None] * colors).sum(-2) + w[..., -1:] * bg
    img = torch.cat((rgb, (1 - torch.prod(1 - prob, -1))[..., None]), -1)
    (img * g["grad_image"].to(dev)).sum().backward()
    assert rel_err(d.grad.cpu(), g["grad_dists"]) <= RTOL
    assert rel_err(z.grad.cpu(), g["grad_zbuf"]) <= RTOL
    _scalars_close(rast.sigma.grad.item(), g["grad_sigma"], "sigma")
    _scalars_close(agg.gamma.grad.item(), g["grad_gamma"], "gamma")
    _scalars_close(agg.alpha.grad.item(), g["grad_alpha"], "alpha")


def test_standalone_functions_match_golden():
    """randomHeaviside / randomArgmax autograd Functions with arbitrary upstream gradients."""
    import pertrenderer_b200 as pb
    g = load_golden("ops_small")
    dev = "cuda"
    x = g["x"].to(dev).requires_grad_(True)
    sig = torch.tensor(float(g["sigma"]), requires_grad=True)
    with pb.explicit_noise(g["U"].to(dev), None):
        y = pb.randomHeaviside.apply(x, int(g["S"]), sig)
    (y * g["grad_l"].to(dev)).sum().backward()
    assert torch.equal(y.detach().cpu(), g["prob"])
    assert rel_err(x.grad.cpu(), g["grad_x"]) <= RTOL
    _scalars_close(sig.grad.item(), g["grad_sigma"], "sigma")
    z = g["z"].to(dev).requires_grad_(True)
    gam = torch.tensor(float(g["gamma"]), requires_grad=True)
    with pb.explicit_noise(None, g["V"].to(dev)):
        w = pb.randomArgmax.apply(z, int(g["S"]), gam, "gaussian", False)
    (w * g["grad_w"].to(dev)).sum().backward()
    assert torch.equal(w.detach().cpu(), g["weights"])
    assert rel_err(z.grad.cpu(), g["grad_z"]) <= RTOL
    _scalars_close(gam.grad.item(), g["grad_gamma"], "gamma")


@pytest.mark.parametrize("shape", [(1, 16, 16, 50, 16, 16, "realistic"), (2, 8, 12, 50, 64, 64, "dense"),
                                   (1, 9, 7, 100, 12, 20, "realistic"), (1, 4, 4, 300, 8, 8, "dense")])
def test_explicit_noise_matches_oracle_seeded(shape):
    """Fresh seeded inputs at sizes the oracle finishes in seconds, incl. K=100, K>255 (16-bit
    winners), S not a multiple of 4 and S_rast != S_agg."""
    from gpu_util import problem_from_case, run_cuda, run_oracle, synthetic_case
    N, H, W, K, S_r, S_a, kind = shape
    g = synthetic_case(N, H, W, K, S_r, S_a, kind=kind, seed=K + S_r, znear=[1.0 + 0.1 * n for n in range(N)],
                       background=(0.3, 0.6, 0.9), alpha=1.2)
    U, V = O.draw_noise((N, H, W, K), S_r, S_a, generator=torch.Generator().manual_seed(5))
    g["U"], g["V"] = U, V
    st, gr = run_oracle(g, U, V)
    out = run_cuda(problem_from_case(g), g["grad_image"])
    mask = g["pix_to_face"] >= 0
    assert torch.equal(out["counts"][mask], st.counts[mask])
    assert torch.equal(out["winners"].permute(3, 0, 1, 2).long(), st.a_s)
    assert (out["image"] - st.image).abs().max() <= 2e-6
    assert rel_err(out["grad_colors"], gr["colors"]) <= RTOL
    assert rel_err(out["grad_dists"], gr["dists"]) <= RTOL
    assert rel_err(out["grad_zbuf"], gr["zbuf"]) <= RTOL
    assert elementwise_close(out["grad_colors"], gr["colors"], RTOL) and elementwise_close(out["grad_dists"], gr["dists"], RTOL)
    assert elementwise_close(out["grad_zbuf"], gr["zbuf"], RTOL) and elementwise_close(out["image"], st.image, RTOL)
    for i, k in enumerate(("sigma", "gamma", "alpha")):
        _scalars_close(out["scalars"][i].item(), gr[k].item(), k)


def _holey_case(N, H, W, K, S_r, S_a, seed):
    """Fragments whose valid entries are NOT a prefix of K (holes anywhere, some pixels empty, some
    full), unsorted depths, depths beyond zfar and coverage counts of 0 and S: the reference takes the
    mask from pix_to_face >= 0 wherever it is (random_rasterizer.py:46)."""
    from gpu_util import synthetic_case
    g = synthetic_case(N, H, W, K, S_r, S_a, kind="dense", seed=seed, background=(0.7, 0.1, 0.4),
                       znear=[0.5 + 0.25 * n for n in range(N)], zfar=[7.0 + n for n in range(N)], alpha=0.9)
    gen = torch.Generator().manual_seed(seed)
    keep = torch.rand((N, H, W, K), generator=gen) < 0.3
    keep[0, 0, 0] = False  # an empty pixel
    keep[0, 0, 1 % W] = True  # a full one
    g["pix_to_face"] = torch.where(keep, g["pix_to_face"], torch.full_like(g["pix_to_face"], -1))
    g["zbuf"] = torch.where(keep, 5.5 + 2.5 * torch.rand((N, H, W, K), generator=gen), torch.full_like(g["zbuf"], -1.0))
    g["dists"] = torch.where(keep, g["dists"] * 0.6, torch.full_like(g["dists"], -1.0))
    g["colors"] = g["colors"] * keep[..., None]
    return g


@pytest.mark.parametrize("shape", [
    (1, 5, 7, 1, 8, 8), (2, 3, 3, 2, 5, 3), (1, 6, 5, 16, 4, 4), (1, 5, 5, 17, 7, 9), (1, 3, 7, 33, 16, 16),
    (1, 3, 5, 64, 8, 8), (1, 3, 3, 65, 8, 8), (1, 2, 3, 255, 4, 4), (1, 2, 2, 256, 4, 4), (1, 1, 3, 1023, 4, 4),
    (1, 3, 3, 50, 1, 2), (1, 2, 2, 50, 600, 520)])
def test_edge_shapes_holes_and_sample_counts(shape):
    """Every tile geometry (K = 1 .. 1023: 32, 16, 8, 4 pixels per warp tile; 8- and 16-bit winners),
    pixel counts that are not a multiple of the tile, masks with holes, S not a multiple of 4, S = 1,
    S_rast != S_agg and S large enough for several backward sample chunks."""
    from gpu_util import problem_from_case, run_cuda, run_oracle
    N, H, W, K, S_r, S_a = shape
    g = _holey_case(N, H, W, K, S_r, S_a, seed=1000 + K + S_r)
    U, V = O.draw_noise((N, H, W, K), S_r, S_a, generator=torch.Generator().manual_seed(K))
    g["U"], g["V"] = U, V
    st, gr = run_oracle(g, U, V)
    out = run_cuda(problem_from_case(g), g["grad_image"])
    mask = g["pix_to_face"] >= 0
    assert torch.equal(out["counts"][mask], st.counts[mask])
    assert torch.equal(out["winners"].permute(3, 0, 1, 2).long(), st.a_s)
    assert (out["image"] - st.image).abs().max() <= 2e-6
    assert rel_err(out["grad_colors"], gr["colors"]) <= RTOL
    assert rel_err(out["grad_dists"], gr["dists"]) <= RTOL
    assert rel_err(out["grad_zbuf"], gr["zbuf"]) <= RTOL
    assert (out["grad_dists"][~mask] == 0).all() and (out["grad_zbuf"][~mask] == 0).all()
    assert (out["grad_colors"][~mask] == 0).all()
    for i, k in enumerate(("sigma", "gamma", "alpha")):
        _scalars_close(out["scalars"][i].item(), gr[k].item(), k)
    # the in-kernel noise path on the same inputs: Philox == explicit(noise_fill) on indices
    from pertrenderer_b200 import _cabi, ops
    Up = ops.noise_fill(5, 0, (N, H, W, K), S_r, "cuda")
    Vp = ops.noise_fill(6, 1, (N, H, W, K), S_a, "cuda")
    a = run_cuda(problem_from_case(g, explicit=False, seed_rast=5, seed_agg=6, flags=_cabi.F_PER_SAMPLE_NOISE), g["grad_image"])
    g2 = dict(g, U=Up.cpu(), V=Vp.cpu())
    b = run_cuda(problem_from_case(g2, explicit=True), g["grad_image"])
    assert torch.equal(a["counts"][mask], b["counts"][mask]) and torch.equal(a["winners"], b["winners"])
    # same weights; bounded noise prunes logits that cannot win, so the blend may sum in another lane order
    assert (a["image"] - b["image"]).abs().max() <= 1e-6
    assert rel_err(a["grad_zbuf"], b["grad_zbuf"]) <= RTOL and rel_err(a["grad_dists"], b["grad_dists"]) <= RTOL


def test_unaligned_views_take_the_scalar_path():
    """Tensors whose storage offset breaks the 16-byte alignment of the tile rows (a view into a larger
    buffer) give the same results as aligned copies."""
    from gpu_util import problem_from_case, run_cuda, synthetic_case
    g = synthetic_case(1, 5, 5, 7, 8, 8, kind="realistic", seed=71)
    a = run_cuda(problem_from_case(g, explicit=False, seed_rast=1, seed_agg=2), g["grad_image"])

    def shifted(t, elems):
        buf = torch.zeros(t.numel() + elems, dtype=t.dtype, device="cuda")
        v = buf[elems:].view(t.shape)
        v.copy_(t)
        return v

    pr = problem_from_case(g, explicit=False, seed_rast=1, seed_agg=2)
    pr.pix_to_face = shifted(pr.pix_to_face, 1)  # 8-byte offset
    pr.zbuf, pr.dists = shifted(pr.zbuf, 1), shifted(pr.dists, 3)
    assert pr.pix_to_face.data_ptr() % 16 == 8
    b = run_cuda(pr, g["grad_image"])
    for k in ("image", "winners", "grad_dists", "grad_zbuf", "grad_colors", "scalars"):
        assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("kind", ["realistic", "dense"])
def test_philox_equals_explicit_with_materialised_noise(kind):
    """The in-register Philox path is the explicit path fed with pert_noise_fill's tensor: bit-exact
    indices, and the CPU oracle on that same noise agrees (so Philox mode is oracle-checked too)."""
    from gpu_util import problem_from_case, run_cuda, run_oracle, synthetic_case
    from pertrenderer_b200 import ops
    N, H, W, K, S_r, S_a = 2, 10, 10, 50, 16, 32
    g = synthetic_case(N, H, W, K, S_r, S_a, kind=kind, seed=7)
    seed_r, seed_a = 0x1234567890ABCDEF, 0x0FEDCBA987654321
    U = ops.noise_fill(seed_r, 0, (N, H, W, K), S_r, "cuda")
    V = ops.noise_fill(seed_a, 1, (N, H, W, K), S_a, "cuda")
    g["U"], g["V"] = U.cpu(), V.cpu()
    from pertrenderer_b200 import _cabi
    a = run_cuda(problem_from_case(g, explicit=False, seed_rast=seed_r, seed_agg=seed_a,
                                   flags=_cabi.F_PER_SAMPLE_NOISE), g["grad_image"])
    b = run_cuda(problem_from_case(g, explicit=True), g["grad_image"])
    mask = g["pix_to_face"] >= 0
    assert torch.equal(a["counts"][mask], b["counts"][mask])
    assert torch.equal(a["winners"], b["winners"])
    assert torch.equal(a["hist"], b["hist"])
    assert torch.equal(a["image"], b["image"])
    assert rel_err(a["grad_dists"], b["grad_dists"]) <= 1e-6
    assert rel_err(a["grad_zbuf"], b["grad_zbuf"]) <= RTOL
    assert torch.equal(a["grad_colors"], b["grad_colors"])
    st, gr = run_oracle(g, g["U"], g["V"])
    assert torch.equal(a["counts"][mask], st.counts[mask])
    assert torch.equal(a["winners"].permute(3, 0, 1, 2).long(), st.a_s)
    assert (a["image"] - st.image).abs().max() <= 2e-6
    assert rel_err(a["grad_dists"], gr["dists"]) <= RTOL
    assert rel_err(a["grad_zbuf"], gr["zbuf"]) <= RTOL


@pytest.mark.parametrize("kind", ["realistic", "dense"])
def test_philox7_stream_is_oracle_checked_too(kind):
    """PERT_F_PHILOX7: the 7-round stream through the same kernels.  Per-sample mode == the explicit path fed with
    pert_noise_fill(stage | 16) == the CPU oracle on that tensor; the stream is standard normal and differs from the
    10-round one."""
    from gpu_util import problem_from_case, run_cuda, run_oracle, synthetic_case
    from pertrenderer_b200 import _cabi, ops
    N, H, W, K, S_r, S_a = 2, 10, 10, 50, 16, 32
    g = synthetic_case(N, H, W, K, S_r, S_a, kind=kind, seed=8)
    seed_r, seed_a = 0x1122334455667788, 0x0123456789ABCDEF
    U = ops.noise_fill(seed_r, 0 | 16, (N, H, W, K), S_r, "cuda")
    V = ops.noise_fill(seed_a, 1 | 16, (N, H, W, K), S_a, "cuda")
    assert not torch.equal(V, ops.noise_fill(seed_a, 1, (N, H, W, K), S_a, "cuda"))
    g["U"], g["V"] = U.cpu(), V.cpu()
    a = run_cuda(problem_from_case(g, explicit=False, seed_rast=seed_r, seed_agg=seed_a,
                                   flags=_cabi.F_PER_SAMPLE_NOISE | _cabi.F_PHILOX7), g["grad_image"])
    st, gr = run_oracle(g, g["U"], g["V"])
    mask = g["pix_to_face"] >= 0
    assert torch.equal(a["counts"][mask], st.counts[mask])
    assert torch.equal(a["winners"].permute(3, 0, 1, 2).long(), st.a_s)
    assert (a["image"] - st.image).abs().max() <= 2e-6
    assert rel_err(a["grad_dists"], gr["dists"]) <= RTOL
    assert rel_err(a["grad_zbuf"], gr["zbuf"]) <= RTOL
    # default mode with the flag: finite, deterministic, same scene
    d1 = run_cuda(problem_from_case(g, explicit=False, seed_rast=seed_r, seed_agg=seed_a, flags=_cabi.F_PHILOX7), g["grad_image"])
    d2 = run_cuda(problem_from_case(g, explicit=False, seed_rast=seed_r, seed_agg=seed_a, flags=_cabi.F_PHILOX7), g["grad_image"])
    for k in ("image", "grad_dists", "grad_zbuf", "scalars"):
        assert torch.isfinite(d1[k]).all() and torch.equal(d1[k], d2[k]), k
    n = ops.noise_fill(77, 1 | 16, (4, 32, 32, 50), 32, "cuda").double().flatten()
    m = n.numel()
    assert abs(n.mean().item()) < 5 / m ** 0.5 and abs(n.var().item() - 1) < 5 * (2 / m) ** 0.5
    assert abs((n ** 3).mean().item()) < 5 * (15 / m) ** 0.5 and abs((n ** 4).mean().item() - 3) < 5 * (96 / m) ** 0.5
    s = n[:: max(1, m // 200000)].sort().values
    emp = torch.arange(1, s.numel() + 1, dtype=torch.float64, device=s.device) / s.numel()
    assert (O.normal_cdf(s).to(s.device) - emp).abs().max().item() < 1.95 / s.numel() ** 0.5
    # neighbouring counters decorrelate (the avalanche of 7 rounds): samples 4q..4q+3 vs 4q+4..4q+7, slot k vs k+1
    v = ops.noise_fill(78, 1 | 16, (1, 64, 64, 31), 8, "cuda").double()
    assert abs(torch.corrcoef(torch.stack((v[0].flatten(), v[4].flatten())))[0, 1].item()) < 5 / v[0].numel() ** 0.5
    assert abs(torch.corrcoef(torch.stack((v[..., :-1].flatten(), v[..., 1:].flatten())))[0, 1].item()) < 5 / v[..., 1:].numel() ** 0.5


@pytest.mark.parametrize("kind", ["realistic", "dense"])
def test_skipping_is_exact(kind):
    """The work-skipping rules (|x| beyond the noise bound, the radius gate, logits that cannot win,
    samples with c_s = 0) change no output bit: per-sample noise == PERT_F_NO_SKIP brute force."""
    from gpu_util import problem_from_case, run_cuda, synthetic_case
    from pertrenderer_b200 import _cabi
    g = synthetic_case(2, 12, 12, 50, 16, 16, kind=kind, seed=11, gamma=1e-3)
    a = run_cuda(problem_from_case(g, explicit=False, seed_rast=3, seed_agg=4, flags=_cabi.F_PER_SAMPLE_NOISE),
                 g["grad_image"])
    b = run_cuda(problem_from_case(g, explicit=False, seed_rast=3, seed_agg=4, flags=_cabi.F_NO_SKIP), g["grad_image"])
    mask = g["pix_to_face"] >= 0
    assert torch.equal(a["counts"][mask], b["counts"][mask])
    for k in ("winners", "hist", "image", "grad_colors"):
        assert torch.equal(a[k], b[k]), k
    # float sums: same terms (skipped ones are exact zeros), but the number of lanes that share one
    # sum adapts to how much work a tile has, so the association order may differ
    assert rel_err(a["rsum"], b["rsum"]) <= 1e-6
    for k in ("grad_dists", "grad_zbuf", "scalars"):
        assert rel_err(a[k], b[k]) <= 2e-6, k


def test_noise_stream_is_standard_normal():
    from pertrenderer_b200 import ops
    n = ops.noise_fill(42, 1, (4, 32, 32, 50), 32, "cuda").double().flatten()
    m = n.numel()
    assert abs(n.mean().item()) < 5 / m ** 0.5
    assert abs(n.var().item() - 1) < 5 * (2 / m) ** 0.5
    assert abs((n ** 3).mean().item()) < 5 * (15 / m) ** 0.5
    assert abs((n ** 4).mean().item() - 3) < 5 * (96 / m) ** 0.5
    assert n.abs().max().item() <= 6.8
    # Kolmogorov-Smirnov on a subsample
    s = n[:: max(1, m // 200000)].sort().values
    cdf = O.normal_cdf(s)
    emp = torch.arange(1, s.numel() + 1, dtype=torch.float64, device=s.device) / s.numel()
    assert (cdf.to(s.device) - emp).abs().max().item() < 1.95 / s.numel() ** 0.5
    # different stages / seeds / samples decorrelate
    a = ops.noise_fill(42, 0, (1, 16, 16, 51), 8, "cuda").flatten()
    b = ops.noise_fill(42, 1, (1, 16, 16, 50), 8, "cuda").flatten()
    c = ops.noise_fill(43, 1, (1, 16, 16, 50), 8, "cuda").flatten()
    assert abs(torch.corrcoef(torch.stack((a, b)))[0, 1].item()) < 0.02
    assert abs(torch.corrcoef(torch.stack((b, c)))[0, 1].item()) < 0.02


def test_philox_closed_forms():
    """Tier 3 (Appendix A.4): with in-kernel noise, E[P_k] = Phi(-d/sigma), the coverage score is
    phi(x/sigma)/sigma, the two-way weight is Phi(dzeta/(gamma sqrt2)), within Monte-Carlo CIs."""
    import pertrenderer_b200 as pb
    dev = "cuda"
    sigma, gamma, S = 1e-3, 1e-2, 4096
    n = 21
    d = torch.linspace(-3e-3, 3e-3, n, device=dev).reshape(1, 1, n, 1).contiguous().requires_grad_(True)
    torch.manual_seed(0)
    p = pb.GaussianRast(nb_samples=S, sigma=sigma).rasterize(d)
    p.sum().backward()
    e = O.expected_coverage(d.detach().cpu(), sigma)
    se = (e * (1 - e) / S).sqrt() + 1e-4
    assert ((p.detach().cpu().double() - e).abs() <= 5 * se).all()
    eg = -O.expected_coverage_score(d.detach().cpu(), sigma)
    assert ((d.grad.cpu().double() - eg).abs() <= 6 * (1 / sigma) / S ** 0.5).all()
    zeta = torch.tensor([0.0, -0.7e-2, float("-inf")], device=dev).reshape(1, 1, 1, 3).repeat(1, 8, 8, 1)
    w = pb.randomArgmax.apply(zeta, S, torch.tensor(gamma))
    ew = O.expected_two_way_weight(0.0, -0.7e-2, gamma).item()
    assert (w[..., 2] == 0).all() and torch.allclose(w.sum(-1), torch.ones_like(w[..., 0]))
    assert ((w[..., 0].cpu().double() - ew).abs() <= 5 * (ew * (1 - ew) / S) ** 0.5).all()


def test_philox_statistics_match_reference_estimator():
    """Tier 3: the mean image / mean gradients over many Philox seeds agree with the mean of the
    reference estimator (oracle with torch.normal noise) within Monte-Carlo confidence intervals."""
    from gpu_util import problem_from_case, run_cuda, run_oracle, synthetic_case
    N, H, W, K, S = 1, 6, 6, 8, 16
    g = synthetic_case(N, H, W, K, S, S, kind="dense", seed=21)
    reps = 300
    acc = {k: [] for k in ("image", "grad_dists", "grad_zbuf")}
    ref = {k: [] for k in acc}
    gen = torch.Generator().manual_seed(1)
    for r in range(reps):
        out = run_cuda(problem_from_case(g, explicit=False, seed_rast=1000 + r, seed_agg=5000 + r), g["grad_image"])
        for k in acc:
            acc[k].append(out[k])
        U, V = O.draw_noise((N, H, W, K), S, S, generator=gen)
        st, gr = run_oracle(g, U, V)
        ref["image"].append(st.image)
        ref["grad_dists"].append(gr["dists"])
        ref["grad_zbuf"].append(gr["zbuf"])
    for k in acc:
        a, b = torch.stack(acc[k]).double(), torch.stack(ref[k]).double()
        diff = a.mean(0) - b.mean(0)
        se = (a.var(0) / reps + b.var(0) / reps).sqrt()
        scale = se.max()
        z = diff.abs() / (se + 1e-3 * scale)
        # ~600 comparisons: a 5.5 sigma bound keeps the family-wise false-alarm rate < 1e-4
        assert z.max().item() < 5.5, (k, z.max().item())


def test_batch_shards_reproduce_the_whole_job():
    """Row (e): pixels are independent and the Philox counter uses the global pixel index, so a job
    split by batch element over ranks gives bit-identical images / gradients; only the three scalar
    gradients need a sum."""
    from gpu_util import problem_from_case, run_cuda, synthetic_case
    g = synthetic_case(4, 8, 8, 50, 16, 16, kind="realistic", seed=31, znear=[1.0, 1.1, 1.2, 1.3])
    whole = run_cuda(problem_from_case(g, explicit=False, seed_rast=9, seed_agg=10), g["grad_image"])
    parts = []
    for r in range(2):
        sl = slice(2 * r, 2 * r + 2)
        gs = dict(g)
        for k in ("pix_to_face", "zbuf", "dists", "colors", "znear", "zfar", "grad_image"):
            gs[k] = g[k][sl].contiguous()
        pr = problem_from_case(gs, explicit=False, seed_rast=9, seed_agg=10, pixel_offset=2 * r * 64)
        parts.append(run_cuda(pr, gs["grad_image"]))
    for k in ("image", "grad_dists", "grad_zbuf", "grad_colors"):
        assert torch.equal(torch.cat([p[k] for p in parts]), whole[k]), k
    tot = parts[0]["scalars"] + parts[1]["scalars"]
    assert torch.allclose(tot, whole["scalars"], rtol=1e-5, atol=1e-7)


def test_sample_shards_reproduce_the_whole_job():
    """Row (e), noise-sample sharding: phases with sums in between (counts, rsum | hist | acc,
    pixstat) reproduce the single-call result: integer state exactly, gradients to rounding."""
    from gpu_util import counts_u16, problem_from_case, run_cuda, synthetic_case
    from pertrenderer_b200 import _cabi, ops
    g = synthetic_case(1, 8, 8, 50, 32, 32, kind="dense", seed=41)
    PS = _cabi.F_PER_SAMPLE_NOISE  # the once-per-logit draws of the default mode are per shard by construction
    whole = run_cuda(problem_from_case(g, explicit=False, seed_rast=9, seed_agg=10, flags=PS), g["grad_image"])
    R = 2
    prs = [problem_from_case(g, explicit=False, seed_rast=9, seed_agg=10, s_rast=(16 * r, 16 * r + 16),
                             s_agg=(16 * r, 16 * r + 16), flags=PS) for r in range(R)]
    # fwd phase 1 on every shard, then "all-reduce" counts and rsum
    saved = [ops.shade_forward(pr, want_hist=True, phases=_cabi.PH_RAST)[1] for pr in prs]
    counts = sum(counts_u16(s) for s in saved)
    rsum = sum(s.rsum for s in saved)
    for s in saved:
        s.counts.copy_(counts.to(torch.int16))
        s.rsum.copy_(rsum)
    for pr, s in zip(prs, saved):
        ops.shade_forward(pr, phases=_cabi.PH_AGG, saved=s)
    hist = sum(s.hist for s in saved)
    for s in saved:
        s.hist.copy_(hist)
    images = [ops.shade_forward(pr, phases=_cabi.PH_BLEND, saved=s)[0] for pr, s in zip(prs, saved)]
    assert torch.equal(counts.cpu(), whole["counts"])
    assert torch.equal(hist.cpu(), whole["hist"])
    assert torch.equal(torch.cat([s.winners_full() for s in saved], -1).cpu(), whole["winners"])
    for im in images:  # same weights; the blend sums them in a different lane order
        assert (im.cpu() - whole["image"]).abs().max() <= 1e-6
    assert torch.allclose(rsum.cpu(), whole["rsum"], rtol=1e-6, atol=1e-6)
    # bwd: sample phase per shard, sum, finish everywhere
    P, K1 = 64, 51
    gi = g["grad_image"].cuda()
    accs = [torch.empty((P, K1), device="cuda") for _ in range(R)]
    stats = [torch.empty((P, 2), device="cuda") for _ in range(R)]
    for pr, s, a, t in zip(prs, saved, accs, stats):
        ops.shade_backward(pr, s, gi, phases=_cabi.PH_BWD_SAMPLE, acc=a, pixstat=t, use_hist=True)
    acc, stat = sum(accs), sum(stats)
    for pr, s in zip(prs, saved):
        gd, gz, gc, scal = ops.shade_backward(pr, s, gi, phases=_cabi.PH_BWD_FINISH, acc=acc, pixstat=stat,
                                              use_hist=True)
        assert rel_err(gd.cpu(), whole["grad_dists"]) <= RTOL
        assert rel_err(gz.cpu(), whole["grad_zbuf"]) <= RTOL
        assert torch.equal(gc.cpu(), whole["grad_colors"])
        assert torch.allclose(scal.cpu(), whole["scalars"], rtol=1e-4, atol=1e-6)


def test_dead_noise_modes_share_forward_and_live_gradients():
    """Backward noise modes (include/pertshade.h): default (one draw per never-winning logit),
    PERT_F_PER_SAMPLE_NOISE (reference-like) and PERT_F_SKIP_DEAD_NOISE (expectation) share the
    forward pass bit for bit, grad_colors, and the score sums of every logit that can win; padded
    entries stay exactly zero in all of them."""
    from gpu_util import problem_from_case, run_cuda, synthetic_case
    from pertrenderer_b200 import _cabi
    from pertrenderer_b200 import ops
    g = synthetic_case(1, 12, 12, 20, 32, 32, kind="realistic", seed=51)
    # ONE forward (the per-sample flag also selects the per-sample coverage draws, which would change the sample
    # path: the compound sampler of the default mode is tested in test_gpu_compound.py), three backward modes
    a = run_cuda(problem_from_case(g, explicit=False, seed_rast=1, seed_agg=2, flags=0), g["grad_image"])
    runs = {0: a}
    for f in (_cabi.F_PER_SAMPLE_NOISE, _cabi.F_SKIP_DEAD_NOISE):
        pr = problem_from_case(g, explicit=False, seed_rast=1, seed_agg=2, flags=f)
        gd, gz, gc, scal = ops.shade_backward(pr, a["saved"], g["grad_image"].cuda())
        runs[f] = dict(grad_dists=gd.cpu(), grad_zbuf=gz.cpu(), grad_colors=gc.cpu(), scalars=scal.cpu())
    b, c = runs[_cabi.F_PER_SAMPLE_NOISE], runs[_cabi.F_SKIP_DEAD_NOISE]
    mask = g["pix_to_face"] >= 0
    for r in (b, c):
        assert torch.equal(a["grad_colors"], r["grad_colors"])
        assert (r["grad_zbuf"][~mask] == 0).all() and (r["grad_dists"][~mask] == 0).all()
        assert torch.isfinite(r["scalars"]).all()
    # entries whose logit was selected at least once are certainly "live": same score sums in every mode,
    # except at the per-pixel argmax of zi, which also collects -sum_j grad_zeta_j
    hist = a["hist"][..., :-1]
    zi = torch.where(mask, (100.0 - g["zbuf"]) / 99.0, torch.zeros_like(g["zbuf"]))
    not_argzi = zi < zi.max(-1, keepdim=True).values
    sel = mask & (hist > 0) & not_argzi
    assert sel.any()
    for r in (b, c):  # same terms; the lanes that share one sum (hence its association order) may differ
        assert rel_err(a["grad_zbuf"][sel], r["grad_zbuf"][sel]) <= 1e-5


@pytest.mark.parametrize("shape", [(1, 6, 6, 12, 16, 6.0, 400), (1, 5, 5, 50, 16, 9.0, 300)])
def test_once_per_logit_noise_has_the_reference_distribution(shape):
    """Default mode (compound coverage draws, one draw per never-winning logit in backward) vs PERT_F_PER_SAMPLE_NOISE
    over many seeds: the gradients have the same mean AND the same variance entry by entry (both shortcuts are the
    exact conditional laws of the per-sample sums), and so do the scalar gradients, the gamma gradient included (its
    squared-noise term keeps mean and variance).  K = 12 and the benchmark's K = 50, both with padded pixels."""
    from gpu_util import problem_from_case, run_cuda, synthetic_case
    from pertrenderer_b200 import _cabi
    N, H, W, K, S, mean_valid, reps = shape
    g = synthetic_case(N, H, W, K, S, S, kind="realistic", seed=61, mean_valid=mean_valid)
    assert (g["pix_to_face"] < 0).any() and (g["pix_to_face"] >= 0).any()
    keys = ("grad_zbuf", "grad_dists", "scalars")
    acc = {m: {k: [] for k in keys} for m in (0, 1)}
    for r in range(reps):
        for m, f in ((0, 0), (1, _cabi.F_PER_SAMPLE_NOISE)):
            # same coverage/argmax noise in both modes would correlate them; use disjoint seeds
            out = run_cuda(problem_from_case(g, explicit=False, seed_rast=100 + 2 * r + m, seed_agg=9000 + 2 * r + m,
                                             flags=f), g["grad_image"])
            for k in keys:
                acc[m][k].append(out[k].double())
    for k in keys:
        a, b = torch.stack(acc[0][k]), torch.stack(acc[1][k])
        se = (a.var(0) / reps + b.var(0) / reps).sqrt()
        z = (a.mean(0) - b.mean(0)).abs() / (se + 1e-3 * se.max() + 1e-30)
        assert z.max().item() < 5.5, (k, "mean", z.max().item())
        # variance: z-score with the standard error of a sample variance, sqrt((m4 - var^2) / reps), estimated from the
        # samples (gradients of rarely flipping coverage entries are heavy tailed: a plain ratio bound fails between two
        # runs of the SAME mode, tools/diag_law.py)
        va, vb = a.var(0), b.var(0)
        m4a, m4b = ((a - a.mean(0)) ** 4).mean(0), ((b - b.mean(0)) ** 4).mean(0)
        sev = ((m4a - va * va).clamp(min=0) / reps + (m4b - vb * vb).clamp(min=0) / reps).sqrt()
        zv = (va - vb).abs() / (sev + 1e-3 * sev.max() + 1e-30)
        assert zv.max().item() < 6.0, (k, "var", zv.max().item())
        # ... and the bulk of the entries (the well-sampled ones) agree closely
        big = vb > 0.2 * vb.max()
        if big.sum() >= 8:
            assert abs((va[big] / vb[big]).log().median().item()) < 0.25, (k, "var ratio")


def test_face_colour_gather_matches_texel_tensor():
    """north_star: the texel gather through pix_to_face inlined into the fused kernels.  Rendering with
    lazy FaceTexels equals rendering with the materialised (N,H,W,K,3) tensor (same forward bits), and the
    gradient scattered onto the (F,3) table equals the index_add of the dense texel gradient."""
    import pertrenderer_b200 as pb
    dev = "cuda"
    N, H, W, K, S, F = 2, 24, 24, 50, 16, 37
    fr, _ = pb.synthetic_fragments(N, H, W, K, kind="realistic", sigma=1e-3, n_faces=F, seed=9, device=dev)
    G = torch.randn((N, H, W, 4), device=dev, generator=torch.Generator(device=dev).manual_seed(2))
    table = torch.rand((F, 3), device=dev, generator=torch.Generator(device=dev).manual_seed(3))

    def run(use_face):
        fc = table.clone().requires_grad_(True)
        d = fr.dists.clone().requires_grad_(True)
        z = fr.zbuf.clone().requires_grad_(True)
        frag = pb.Fragments(fr.pix_to_face, z, None, d)
        shader = pb.RandomSimpleShader(device=dev, cameras=pb.DepthCameras(n=N, device=dev),
                                       smoothrast=pb.GaussianRast(nb_samples=S, sigma=1e-3),
                                       smoothagg=pb.GaussianAgg(nb_samples=S, gamma=1e-2),
                                       blend_params=pb.BlendParams(background_color=(0.1, 0.2, 0.3)))
        meshes = pb.FaceColorMeshes(fc) if use_face else pb.TexelMeshes(pb.FaceTexels(fc).materialize(fr.pix_to_face))
        torch.manual_seed(11)
        img = shader(frag, meshes)
        (img * G).sum().backward()
        return img.detach(), fc.grad, d.grad, z.grad

    img_f, gfc_f, gd_f, gz_f = run(True)
    img_t, gfc_t, gd_t, gz_t = run(False)
    assert torch.equal(img_f, img_t)
    assert torch.equal(gd_f, gd_t) and torch.equal(gz_f, gz_t)
    assert rel_err(gfc_f.cpu(), gfc_t.cpu()) <= 1e-5
    assert gfc_f.abs().sum() > 0


def test_full_size_properties_config2():
    """BASELINE config 2 (N=8, 256x256, K=50, S=64) through the public API: size-independent
    properties — determinism under a fixed torch seed, alpha in [0,1] and quantised, empty pixels
    render the background with zero gradients, weights sum to one (rgb of an all-ones colour
    tensor is exactly 1 where covered by a white background), finite gradients."""
    import pertrenderer_b200 as pb
    dev = "cuda"
    N, H, W, K, S = 8, 256, 256, 50, 64
    fr, col = pb.synthetic_fragments(N, H, W, K, kind="realistic", sigma=1e-3, seed=0, device=dev)
    col = torch.ones_like(col)
    shader = pb.RandomSimpleShader(device=dev, cameras=pb.DepthCameras(n=N, device=dev),
                                   smoothrast=pb.GaussianRast(nb_samples=S, sigma=1e-3),
                                   smoothagg=pb.GaussianAgg(nb_samples=S, gamma=1e-2),
                                   blend_params=pb.BlendParams(background_color=(1.0, 1.0, 1.0)))

    def render():
        d = fr.dists.clone().requires_grad_(True)
        z = fr.zbuf.clone().requires_grad_(True)
        c = col.clone().requires_grad_(True)
        torch.manual_seed(7)
        img = shader(pb.Fragments(fr.pix_to_face, z, None, d), pb.TexelMeshes(c))
        img.square().sum().backward()
        return img.detach(), d.grad, z.grad, c.grad

    img, gd, gz, gc = render()
    img2, gd2, gz2, gc2 = render()
    assert torch.equal(img, img2) and torch.equal(gd, gd2) and torch.equal(gz, gz2) and torch.equal(gc, gc2)
    assert img.shape == (N, H, W, 4)
    assert (img[..., 3] >= 0).all() and (img[..., 3] <= 1).all()
    assert (img[..., :3] - 1).abs().max() <= 1e-5  # weights sum to one
    empty = (fr.pix_to_face < 0).all(-1)
    assert (img[empty][:, 3] == 0).all()
    pad = fr.pix_to_face < 0
    assert (gd[pad] == 0).all() and (gz[pad] == 0).all()
    assert torch.isfinite(gd).all() and torch.isfinite(gz).all() and torch.isfinite(gc).all()
    # grad_colors = w_k * dL/drgb: per pixel the colour gradients sum to dL/drgb * (1 - w_bg)
    assert (gc.sum(-2) <= 2 * img[..., :3].abs() + 1e-5).all()
    s, gm, al = shader.get_smoothing()
    assert all(torch.isfinite(t.grad) for t in (s, gm, al))


def test_integration_md_ctypes_stub_runs():
    """The ctypes stub printed in INTEGRATION.md (what a maintainer of the reference would paste into
    randomras/random_rasterizer.py) is executable as written and reproduces the package's result."""
    import os
    import re
    import pertrenderer_b200 as pb
    from conftest import ROOT
    from pertrenderer_b200 import _cabi
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = [b for b in blocks if "class PerturbedShade" in b][0].replace('"libpertshade.so"', repr(_cabi.LIB_PATH))
    ns = {}
    exec(stub, ns)
    dev = "cuda"
    N, H, W, K, S = 2, 12, 12, 50, 16
    fr, col = pb.synthetic_fragments(N, H, W, K, kind="realistic", sigma=1e-3, seed=4, device=dev)
    G = torch.randn((N, H, W, 4), device=dev, generator=torch.Generator(device=dev).manual_seed(8))
    zn, zf = torch.ones(N, device=dev), torch.full((N,), 100.0, device=dev)

    def leaves():
        return (col.clone().requires_grad_(True), fr.dists.clone().requires_grad_(True), fr.zbuf.clone().requires_grad_(True))

    c, d, z = leaves()
    sig, gam, alp = (torch.tensor(v, requires_grad=True) for v in (1e-3, 1e-2, 1.0))
    torch.manual_seed(21)
    img = ns["PerturbedShade"].apply(c, d, z, sig, gam, alp, fr.pix_to_face, zn, zf, S, (1.0, 1.0, 1.0), 1e-10)
    (img * G).sum().backward()
    c2, d2, z2 = leaves()
    rast, agg = pb.GaussianRast(nb_samples=S, sigma=1e-3), pb.GaussianAgg(nb_samples=S, gamma=1e-2, alpha=1.0)
    torch.manual_seed(21)
    img2 = pb.smooth_rgb_blend(c2, pb.Fragments(fr.pix_to_face, z2, None, d2), rast, agg, pb.BlendParams(), znear=zn, zfar=zf)
    (img2 * G).sum().backward()
    assert torch.equal(img, img2)
    assert torch.equal(d.grad, d2.grad) and torch.equal(z.grad, z2.grad) and torch.equal(c.grad, c2.grad)
    assert torch.allclose(sig.grad, rast.sigma.grad) and torch.allclose(gam.grad, agg.gamma.grad)


def test_large_k_many_pixels_default_path():
    """K near the maximum with enough pixels for the production tile geometry: the sparse-first mode must
    switch itself off when its fallback pass would not fit in shared memory, and stay exact."""
    from gpu_util import problem_from_case, run_cuda
    from pertrenderer_b200 import _cabi
    N, H, W, K, S = 1, 96, 96, 1000, 4
    g = _holey_case(N, H, W, K, S, S, seed=5)
    a = run_cuda(problem_from_case(g, explicit=False, seed_rast=1, seed_agg=2), g["grad_image"])
    b = run_cuda(problem_from_case(g, explicit=False, seed_rast=1, seed_agg=2, flags=_cabi.F_NO_SKIP), g["grad_image"])
    c = run_cuda(problem_from_case(g, explicit=False, seed_rast=1, seed_agg=2, flags=_cabi.F_PER_SAMPLE_NOISE), g["grad_image"])
    mask = g["pix_to_face"] >= 0
    # the per-sample path is the brute-force run bit for bit
    assert torch.equal(c["counts"][mask], b["counts"][mask]) and torch.equal(c["winners"], b["winners"])
    assert (c["image"] - b["image"]).abs().max() <= 1e-6
    assert torch.equal(c["grad_colors"], b["grad_colors"])
    # the default mode draws the coverage flips with the compound sampler and the never-winning logits' noise once per
    # logit (another sample path than the brute-force run): same estimator, so here only sanity + determinism; the
    # distributions are compared in test_gpu_compound.py and test_once_per_logit_noise_has_the_reference_distribution
    a2 = run_cuda(problem_from_case(g, explicit=False, seed_rast=1, seed_agg=2), g["grad_image"])
    for k in ("image", "grad_dists", "grad_zbuf", "scalars"):
        assert torch.isfinite(a[k]).all() and torch.equal(a[k], a2[k]), k
    assert (a["grad_dists"][~mask] == 0).all() and (a["grad_zbuf"][~mask] == 0).all()
    assert (a["image"] - b["image"]).abs().mean() < 0.05  # same scene, other noise


@pytest.mark.parametrize("case", ["small", "k50", "empty"])
def test_softras_pair_matches_reference_golden(case):
    """SoftRast + SoftAgg (the shaders' default operators) through the fused soft kernels and the public
    API, against vectors produced by the unmodified reference."""
    import pertrenderer_b200 as pb
    g = load_golden("soft_" + case)
    dev = "cuda"
    rast = pb.SoftRast(sigma=float(g["sigma"]))
    agg = pb.SoftAgg(gamma=float(g["gamma"]), alpha=float(g["alpha"]))
    d = g["dists"].to(dev).requires_grad_(True)
    z = g["zbuf"].to(dev).requires_grad_(True)
    c = g["colors"].to(dev).requires_grad_(True)
    frag = pb.Fragments(g["pix_to_face"].to(dev), z, None, d)
    blend = pb.BlendParams(background_color=tuple(g["background"].tolist()))
    img = pb.smooth_rgb_blend(c, frag, rast, agg, blend, znear=g["znear"].reshape(-1, 1, 1, 1).to(dev),
                              zfar=g["zfar"].reshape(-1, 1, 1, 1).to(dev))
    (img * g["grad_image"].to(dev)).sum().backward()
    assert (img.detach().cpu() - g["image"]).abs().max() <= 2e-6
    assert rel_err(c.grad.cpu(), g["grad_colors"]) <= RTOL
    assert rel_err(d.grad.cpu(), g["grad_dists"]) <= RTOL
    assert rel_err(z.grad.cpu(), g["grad_zbuf"]) <= RTOL
    mask = g["pix_to_face"] >= 0
    assert (d.grad.cpu()[~mask] == 0).all() and (z.grad.cpu()[~mask] == 0).all() and (c.grad.cpu()[~mask] == 0).all()
    # d/dgamma and d/dalpha are sums of softmax gradients that cancel (sum_j grad y_j = 0): in fp32 the
    # reference itself is only within 2e-5 of the float64 value of these scalars (measured); allow 1e-4
    for got, key in ((rast.sigma.grad, "grad_sigma"), (agg.gamma.grad, "grad_gamma"), (agg.alpha.grad, "grad_alpha")):
        ref = float(g[key])
        assert abs(got.item() - ref) <= 1e-4 * abs(ref) + 1e-8, (key, got.item(), ref)


@pytest.mark.parametrize("shape", [(1, 5, 7, 1), (2, 3, 3, 2), (1, 6, 5, 17), (1, 3, 5, 64), (1, 3, 3, 100), (1, 2, 2, 300)])
def test_softras_pair_matches_oracle_holes(shape):
    """Fused soft kernels vs the CPU oracle on masks with holes, every tile geometry."""
    from gpu_util import problem_from_case
    from pertrenderer_b200 import ops
    N, H, W, K = shape
    g = _holey_case(N, H, W, K, 4, 4, seed=2000 + K)
    g["gamma"] = 4e-2
    zn, zf = g["znear"].reshape(-1, 1, 1, 1), g["zfar"].reshape(-1, 1, 1, 1)
    image, prob, weights, gr = O.soft_shade_fwd_bwd(g["pix_to_face"], g["zbuf"], g["dists"], g["colors"], g["background"], zn, zf,
                                                    g["sigma"], g["gamma"], g["alpha"], g["eps"], g["grad_image"])
    pr = problem_from_case(g, explicit=False)
    img = ops.soft_shade_forward(pr)
    gd, gz, gc, scal = ops.soft_shade_backward(pr, g["grad_image"].cuda())
    assert (img.cpu() - image).abs().max() <= 2e-6
    assert rel_err(gc.cpu(), gr["colors"]) <= RTOL and rel_err(gd.cpu(), gr["dists"]) <= RTOL
    assert rel_err(gz.cpu(), gr["zbuf"]) <= RTOL
    for i, k in enumerate(("sigma", "gamma", "alpha")):
        assert abs(scal[i].item() - gr[k].item()) <= 1e-4 * abs(gr[k].item()) + 1e-7, (k, scal[i].item(), gr[k].item())


def test_cauchy_operators_match_reference_golden():
    """ArctanRast / CauchyAgg (exported by randomras/__init__.py): randomHeaviside and randomArgmax with
    "cauchy" noise, fed the reference's own Cauchy draws."""
    import pertrenderer_b200 as pb
    g = load_golden("ops_cauchy")
    dev = "cuda"
    x = g["x"].to(dev).requires_grad_(True)
    sig = torch.tensor(float(g["sigma"]), requires_grad=True)
    with pb.explicit_noise(g["U"].to(dev), None):
        y = pb.randomHeaviside.apply(x, int(g["S"]), sig, "cauchy")
    (y * g["grad_l"].to(dev)).sum().backward()
    assert torch.equal(y.detach().cpu(), g["prob"])
    assert rel_err(x.grad.cpu(), g["grad_x"]) <= RTOL
    _scalars_close(sig.grad.item(), g["grad_sigma"], "sigma")
    z = g["z"].to(dev).requires_grad_(True)
    gam = torch.tensor(float(g["gamma"]), requires_grad=True)
    with pb.explicit_noise(None, g["V"].to(dev)):
        w = pb.randomArgmax.apply(z, int(g["S"]), gam, "cauchy", False)
    (w * g["grad_w"].to(dev)).sum().backward()
    assert torch.equal(w.detach().cpu(), g["weights"])
    assert rel_err(z.grad.cpu(), g["grad_z"]) <= RTOL
    _scalars_close(gam.grad.item(), g["grad_gamma"], "gamma")
    # in-kernel Cauchy stream: E[P] = arctan(x/sigma)/pi + 1/2 (the closed form the reference notes at smoothrast.py:172)
    S, sigma = 4096, 1e-3
    d = torch.linspace(-3e-3, 3e-3, 13, device=dev).reshape(1, 1, 13, 1).contiguous()
    torch.manual_seed(3)
    p = pb.ArctanRast(nb_samples=S, sigma=sigma).rasterize(d).cpu().double()
    e = torch.arctan(-d.cpu().double() / sigma) / torch.pi + 0.5
    assert ((p - e).abs() <= 5 * (e * (1 - e) / S).sqrt() + 1e-4).all()
    n = pb.ops.noise_fill(9, 1 | 2, (2, 16, 16, 7), 32, dev).flatten().double()
    assert abs(n.median().item()) < 0.02 and abs((n.abs() < 1).double().mean().item() - 0.5) < 0.01  # quartiles at +-1


def test_cauchy_wovr_heaviside_matches_reference_golden():
    """randomHeaviside_wovr with "cauchy" noise (smoothrast.py:99-101): no control variate in the Cauchy branch either."""
    import pertrenderer_b200 as pb
    g = load_golden("ops_cauchy_wovr")
    dev = "cuda"
    x = g["x"].to(dev).requires_grad_(True)
    sig = torch.tensor(float(g["sigma"]), requires_grad=True)
    with pb.explicit_noise(g["U"].to(dev), None):
        y = pb.randomHeaviside_wovr.apply(x, int(g["S"]), sig, "cauchy")
    (y * g["grad_l"].to(dev)).sum().backward()
    assert torch.equal(y.detach().cpu(), g["prob"])
    assert rel_err(x.grad.cpu(), g["grad_x"]) <= RTOL
    _scalars_close(sig.grad.item(), g["grad_sigma"], "sigma")
    # ... which differs from the variance-reduced Cauchy estimator on the far-inside entry
    x2 = g["x"].to(dev).requires_grad_(True)
    with pb.explicit_noise(g["U"].to(dev), None):
        y2 = pb.randomHeaviside.apply(x2, int(g["S"]), torch.tensor(float(g["sigma"])), "cauchy")
    (y2 * g["grad_l"].to(dev)).sum().backward()
    assert (x2.grad[..., -1] == 0).all() and (x.grad[..., -1] != 0).any()


def test_wovr_operators_match_reference_golden():
    """GaussianRast_wovr / GaussianAgg_wovr (eval.py:152-154 "gaussian_wovr"): the estimators without control
    variates, fed the reference's recorded noise; and through the shader with the operator-by-operator path."""
    import pertrenderer_b200 as pb
    g = load_golden("ops_wovr")
    dev = "cuda"
    x = g["x"].to(dev).requires_grad_(True)
    sig = torch.tensor(float(g["sigma"]), requires_grad=True)
    with pb.explicit_noise(g["U"].to(dev), None):
        y = pb.randomHeaviside_wovr.apply(x, int(g["S"]), sig)
    (y * g["grad_l"].to(dev)).sum().backward()
    assert torch.equal(y.detach().cpu(), g["prob"])
    assert rel_err(x.grad.cpu(), g["grad_x"]) <= RTOL
    _scalars_close(sig.grad.item(), g["grad_sigma"], "sigma")
    z = g["z"].to(dev).requires_grad_(True)
    gam = torch.tensor(float(g["gamma"]), requires_grad=True)
    with pb.explicit_noise(None, g["V"].to(dev)):
        w = pb.randomArgmax_wovr.apply(z, int(g["S"]), gam, "gaussian", False)
    (w * g["grad_w"].to(dev)).sum().backward()
    assert torch.equal(w.detach().cpu(), g["weights"])
    assert rel_err(z.grad.cpu(), g["grad_z"]) <= RTOL
    _scalars_close(gam.grad.item(), g["grad_gamma"], "gamma")
    # in-kernel noise: same expectation as the variance-reduced estimator, larger spread (the point of the ablation)
    S, sigma = 64, 1e-3
    d = torch.full((1, 64, 64, 4), -3e-3, device=dev)  # 3 sigma inside the face (h0 = 1): Var drops from ~0.98 to ~0.015
    gl = torch.ones_like(d)

    def grad_of(cls, seed):
        torch.manual_seed(seed)
        dd = d.clone().requires_grad_(True)
        (cls(nb_samples=S, sigma=sigma).rasterize(dd) * gl).sum().backward()
        return dd.grad.flatten().double()

    a = torch.cat([grad_of(pb.GaussianRast, s) for s in range(4)])
    b = torch.cat([grad_of(pb.GaussianRast_wovr, s) for s in range(4)])
    assert abs(a.mean() - b.mean()) <= 6 * (b.std() / b.numel() ** 0.5) + 1e-9
    assert b.std() > 3 * a.std()
    # the shader accepts the pair (operator-by-operator composition)
    fr, col = pb.synthetic_fragments(1, 8, 8, 6, kind="dense", device=dev)
    img = pb.smooth_rgb_blend(col, fr, pb.GaussianRast_wovr(nb_samples=8, sigma=1e-3), pb.GaussianAgg_wovr(nb_samples=8, gamma=1e-2),
                              pb.BlendParams())
    assert img.shape == (1, 8, 8, 4) and torch.isfinite(img).all()


def test_forward_only_noises_and_soft_simple_shader():
    """UniformAgg (smoothagg.py:252-272) and the "gumbel" branch of randomArgmax (smoothagg.py:22-24): forward only, as
    in the reference; SoftSimpleShader (random_rasterizer.py:205-215) on the fused soft kernels."""
    import math
    import pertrenderer_b200 as pb
    from pertrenderer_b200 import ops
    dev = "cuda"
    u = ops.noise_fill(5, 1 | 4, (2, 32, 32, 7), 64, dev).flatten().double()
    assert u.min() >= -0.5 and u.max() < 0.5 and abs(u.mean().item()) < 2e-3 and abs(u.var().item() - 1 / 12) < 2e-3
    gmb = ops.noise_fill(5, 1 | 8, (2, 32, 32, 7), 64, dev).flatten().double()
    assert abs(gmb.mean().item() - 0.5772) < 5e-3 and abs(gmb.var().item() - math.pi ** 2 / 6) < 2e-2
    # the in-kernel stream equals the explicit-noise kernel fed the materialised stream
    z = (torch.randn(2, 6, 5, 8, device=dev) * 0.05).requires_grad_(True)
    for name, bit in (("uniform", 4), ("gumbel", 8)):
        torch.manual_seed(11)
        w = pb.randomArgmax.apply(z, 16, torch.tensor(0.05), name, False)
        torch.manual_seed(11)
        seed = ops.draw_seed()
        V = ops.noise_fill(seed, 1 | bit, (2, 6, 5, 7), 16, dev)
        with pb.explicit_noise(None, V):
            w2 = pb.randomArgmax.apply(z.detach(), 16, torch.tensor(0.05), "gaussian", False)
        assert torch.equal(w, w2) and torch.allclose(w.sum(-1), torch.ones_like(w.sum(-1)))
        with pytest.raises(RuntimeError, match="no backward"):
            w.sum().backward()
    fr, col = pb.synthetic_fragments(2, 9, 7, 6, kind="dense", sigma=1e-3, device=dev)
    agg = pb.UniformAgg(nb_samples=8, gamma=1e-2)
    wts = agg.aggregate(fr.zbuf, 100.0, 1.0, torch.rand_like(fr.zbuf), fr.pix_to_face >= 0)
    assert wts.shape == (2, 9, 7, 7) and torch.allclose(wts.sum(-1), torch.ones(2, 9, 7, device=dev))
    # SoftSimpleShader = softmax_rgb_blend(blend_params.sigma, .gamma) = the SoftRas pair at alpha 1
    blend = pb.BlendParams(sigma=2e-3, gamma=5e-2, background_color=(0.2, 0.3, 0.4))
    img = pb.SoftSimpleShader(blend_params=blend)(fr, pb.TexelMeshes(col))
    ref, _, _ = O.soft_shade_fwd_bwd(fr.pix_to_face.cpu(), fr.zbuf.cpu(), fr.dists.cpu(), col.cpu(), torch.tensor(blend.background_color),
                                     1.0, 100.0, 2e-3, 5e-2, 1.0, 1e-10)
    assert (img.cpu() - ref).abs().max() <= 2e-6
