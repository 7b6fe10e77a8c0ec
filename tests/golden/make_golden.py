"""Generate the golden vectors under tests/golden/ by EXECUTING THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference operators (randomras/smoothrast.py, smoothagg.py, random_rasterizer.py) are imported
from /root/reference with the pytorch3d modules stubbed (pytorch3d is not installable here; the hot
path only touches pytorch3d objects by attribute access — SURVEY.md §0.2).  ``torch.normal`` is
wrapped while the reference runs so that the two noise tensors it draws internally
(smoothrast.py:21, smoothagg.py:21) are recorded; everything else is the reference's own code and
autograd.  Outputs are small ``.npz`` files; the script and the files are committed together.
"""

from __future__ import annotations

import os
import sys
import types
from collections import namedtuple

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _stub_pytorch3d():
    names = ["look_at_view_transform", "OpenGLPerspectiveCameras", "PointLights", "DirectionalLights",
             "Materials", "RasterizationSettings", "MeshRenderer", "MeshRasterizer", "SoftPhongShader",
             "HardPhongShader", "SoftSilhouetteShader", "hard_rgb_blend", "softmax_rgb_blend",
             "TexturesVertex", "BlendParams"]
    p3d = types.ModuleType("pytorch3d")
    rend = types.ModuleType("pytorch3d.renderer")
    mesh = types.ModuleType("pytorch3d.renderer.mesh")
    shading = types.ModuleType("pytorch3d.renderer.mesh.shading")
    for n in names:
        setattr(rend, n, type(n, (), {"__init__": lambda self, *a, **k: None}))
    rend.look_at_view_transform = lambda **k: (torch.eye(3)[None], torch.zeros(1, 3))
    shading.phong_shading = lambda **k: None
    p3d.renderer, rend.mesh, mesh.shading = rend, mesh, shading
    sys.modules.update({"pytorch3d": p3d, "pytorch3d.renderer": rend, "pytorch3d.renderer.mesh": mesh,
                        "pytorch3d.renderer.mesh.shading": shading})


def load_reference():
    _stub_pytorch3d()
    sys.path.insert(0, REF)
    import randomras.random_rasterizer as rr  # noqa
    import randomras.smoothagg as sa  # noqa
    import randomras.smoothrast as sr  # noqa
    return rr, sr, sa


Fragments = namedtuple("Fragments", ["pix_to_face", "zbuf", "bary_coords", "dists"])
Blend = namedtuple("Blend", ["sigma", "gamma", "background_color"])


class NoiseRecorder:
    """Context manager: records every tensor torch.normal returns."""

    def __enter__(self):
        self.drawn = []
        self._orig = torch.normal

        def wrapped(*a, **k):
            out = self._orig(*a, **k)
            self.drawn.append(out.detach().clone())
            return out

        torch.normal = wrapped
        return self

    def __exit__(self, *exc):
        torch.normal = self._orig


class NoiseReplayer:
    """Context manager: torch.normal returns pre-recorded tensors, in order."""

    def __init__(self, tensors):
        self.tensors = list(tensors)

    def __enter__(self):
        self._orig = torch.normal
        it = iter(self.tensors)
        torch.normal = lambda *a, **k: next(it).clone()
        return self

    def __exit__(self, *exc):
        torch.normal = self._orig


class FakeCtx:
    needs_input_grad = (True, False, True, False, False)

    def save_for_backward(self, *t):
        self.saved_tensors = t


def make_fragments(gen, N, H, W, K, sigma, n_faces, p_empty, mean_valid, frac_edge, p_outside=0.25):
    """Small synthetic Fragments with every edge case: empty pixels, padding last, faces whose
    coverage probability comes out exactly 0 (far outside), interior faces, ascending zbuf."""
    p2f = torch.full((N, H, W, K), -1, dtype=torch.int64)
    zbuf = torch.full((N, H, W, K), -1.0)
    dists = torch.full((N, H, W, K), -1.0)
    b = float(np.log(1.0 / 1e-4 - 1.0)) * sigma
    for n in range(N):
        for i in range(H):
            for j in range(W):
                if torch.rand((), generator=gen).item() < p_empty:
                    continue
                nv = int(min(K, 1 + torch.poisson(torch.tensor(float(mean_valid - 1)), generator=gen).item()))
                z = torch.sort(5.5 + 2.5 * torch.rand(nv, generator=gen)).values
                p2f[n, i, j, :nv] = torch.randint(0, n_faces, (nv,), generator=gen)
                zbuf[n, i, j, :nv] = z
                r = torch.rand(nv, generator=gen)
                edge = (torch.rand(nv, generator=gen) * 2 - 1) * b
                interior = -0.05 * torch.rand(nv, generator=gen)
                outside = 0.02 + 0.05 * torch.rand(nv, generator=gen)
                d = torch.where(r < frac_edge, edge, interior)
                d = torch.where(r > 1.0 - p_outside * (1 - frac_edge), outside, d)
                dists[n, i, j, :nv] = d
    return p2f, zbuf, dists


def run_case(rr, sr, sa, name, *, N, H, W, K, S_r, S_a, sigma, gamma, alpha, seed,
             znear, zfar, background, p_empty=0.25, mean_valid=3.0, frac_edge=0.6, all_empty_batch=None):
    gen = torch.Generator().manual_seed(seed)
    p2f, zbuf, dists = make_fragments(gen, N, H, W, K, sigma, 12, p_empty, mean_valid, frac_edge)
    if all_empty_batch is not None:
        p2f[all_empty_batch] = -1
        zbuf[all_empty_batch] = -1.0
        dists[all_empty_batch] = -1.0
    colors = torch.rand((N, H, W, K, 3), generator=gen)
    colors = colors * (p2f >= 0)[..., None]
    grad_image = torch.randn((N, H, W, 4), generator=gen)
    zn = torch.tensor(znear, dtype=torch.float32)[:, None, None, None]
    zf = torch.tensor(zfar, dtype=torch.float32)[:, None, None, None]

    rast = sr.GaussianRast(nb_samples=S_r, sigma=sigma)
    agg = sa.GaussianAgg(nb_samples=S_a, gamma=gamma, alpha=alpha)
    d_leaf = dists.clone().requires_grad_(True)
    z_leaf = zbuf.clone().requires_grad_(True)
    c_leaf = colors.clone().requires_grad_(True)
    frags = Fragments(p2f, z_leaf, None, d_leaf)
    torch.manual_seed(seed + 1000)
    with NoiseRecorder() as rec:
        image = rr.smooth_rgb_blend(c_leaf, frags, rast, agg, Blend(sigma, gamma, background), znear=zn, zfar=zf)
    U, V = rec.drawn
    assert U.shape == (S_r, N, H, W, K) and V.shape == (S_a, N, H, W, K + 1)
    (image * grad_image).sum().backward()

    # internals (indices / hit bits): re-run the reference's own stage functions on the same noise
    mask = p2f >= 0
    with NoiseReplayer([U]):
        ctx = FakeCtx()
        p_hat = sr.randomHeaviside.forward(ctx, -dists, S_r, rast.sigma.detach())
        h = ctx.saved_tensors[0]
    prob = p_hat * mask
    with torch.no_grad(), NoiseReplayer([V]):
        z_inv = (zf - zbuf) / (zf - zn) * mask
        z_inv_max = torch.max(z_inv, dim=-1).values[..., None].clamp(min=agg.eps)
        z_map = (agg.gamma / agg.alpha) * prob.log() + z_inv - z_inv_max
        z_map = torch.cat((z_map, torch.ones((N, H, W, 1)) * agg.eps - z_inv_max), dim=-1)
        ctx2 = FakeCtx()
        weights = sa.randomArgmax.forward(ctx2, z_map, S_a, agg.gamma.detach(), "gaussian", False)
        onehots, _, _, vr, _ = ctx2.saved_tensors
    a_s = onehots.argmax(dim=-1)
    a_0 = vr.argmax(dim=-1)

    out = dict(
        pix_to_face=p2f.numpy(), zbuf=zbuf.numpy(), dists=dists.numpy(), colors=colors.numpy(),
        grad_image=grad_image.numpy(), znear=np.asarray(znear, np.float32), zfar=np.asarray(zfar, np.float32),
        background=np.asarray(background, np.float32),
        sigma=np.float32(sigma), gamma=np.float32(gamma), alpha=np.float32(alpha), eps=np.float64(agg.eps),
        S_r=np.int32(S_r), S_a=np.int32(S_a),
        U=U.numpy(), V=V.numpy(),
        image=image.detach().numpy(), prob=prob.numpy(), weights=weights.numpy(),
        counts=h.sum(0).to(torch.int32).numpy(), a_s=a_s.to(torch.int16).numpy(), a_0=a_0.to(torch.int16).numpy(),
        zeta=z_map.numpy(),
        grad_dists=d_leaf.grad.numpy(), grad_zbuf=z_leaf.grad.numpy(), grad_colors=c_leaf.grad.numpy(),
        grad_sigma=rast.sigma.grad.numpy(), grad_gamma=agg.gamma.grad.numpy(), grad_alpha=agg.alpha.grad.numpy(),
    )
    path = os.path.join(OUT, f"shade_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)  "
          f"image mean {image.mean().item():.4f}  |gd| {d_leaf.grad.abs().sum().item():.3e}")


def run_soft_case(rr, sr, sa, name, *, N, H, W, K, sigma, gamma, alpha, seed, znear, zfar, background,
                  p_empty=0.25, mean_valid=3.0, frac_edge=0.6, all_empty_batch=None):
    """The SoftRas pair (the shaders' DEFAULT operators): SoftRast (smoothrast.py:126-134) + SoftAgg
    (smoothagg.py:165-182) through smooth_rgb_blend (random_rasterizer.py:34-56); deterministic."""
    gen = torch.Generator().manual_seed(seed)
    p2f, zbuf, dists = make_fragments(gen, N, H, W, K, sigma, 12, p_empty, mean_valid, frac_edge)
    if all_empty_batch is not None:
        p2f[all_empty_batch] = -1
        zbuf[all_empty_batch] = -1.0
        dists[all_empty_batch] = -1.0
    colors = torch.rand((N, H, W, K, 3), generator=gen) * (p2f >= 0)[..., None]
    grad_image = torch.randn((N, H, W, 4), generator=gen)
    zn = torch.tensor(znear, dtype=torch.float32)[:, None, None, None]
    zf = torch.tensor(zfar, dtype=torch.float32)[:, None, None, None]
    rast = sr.SoftRast(sigma=sigma)
    agg = sa.SoftAgg(gamma=gamma, alpha=alpha)
    d_leaf = dists.clone().requires_grad_(True)
    z_leaf = zbuf.clone().requires_grad_(True)
    c_leaf = colors.clone().requires_grad_(True)
    image = rr.smooth_rgb_blend(c_leaf, Fragments(p2f, z_leaf, None, d_leaf), rast, agg, Blend(sigma, gamma, background),
                                znear=zn, zfar=zf)
    (image * grad_image).sum().backward()
    with torch.no_grad():
        prob = rast.rasterize(dists) * (p2f >= 0)
        weights = agg.aggregate(zbuf, zf, zn, prob, p2f >= 0)
    out = dict(
        pix_to_face=p2f.numpy(), zbuf=zbuf.numpy(), dists=dists.numpy(), colors=colors.numpy(),
        grad_image=grad_image.numpy(), znear=np.asarray(znear, np.float32), zfar=np.asarray(zfar, np.float32),
        background=np.asarray(background, np.float32),
        sigma=np.float32(sigma), gamma=np.float32(gamma), alpha=np.float32(alpha), eps=np.float64(agg.eps),
        image=image.detach().numpy(), prob=prob.numpy(), weights=weights.numpy(),
        grad_dists=d_leaf.grad.numpy(), grad_zbuf=z_leaf.grad.numpy(), grad_colors=c_leaf.grad.numpy(),
        grad_sigma=rast.sigma.grad.numpy(), grad_gamma=agg.gamma.grad.numpy(), grad_alpha=agg.alpha.grad.numpy(),
    )
    path = os.path.join(OUT, f"soft_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"soft {name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)  image mean {image.mean().item():.4f}  "
          f"grads sigma/gamma/alpha {rast.sigma.grad.item():.4e} {agg.gamma.grad.item():.4e} {agg.alpha.grad.item():.4e}")


def run_ops_case(sr, sa, name, *, shape, S, sigma, gamma, seed):
    """Stand-alone autograd Functions with an arbitrary upstream gradient
    (randomHeaviside: smoothrast.py:12-59, randomArgmax: smoothagg.py:10-73)."""
    gen = torch.Generator().manual_seed(seed)
    N, H, W, K = shape
    x = (torch.rand(shape, generator=gen) * 2 - 1) * 3 * sigma
    x[..., -1] = 1.0  # a far-inside entry
    x.requires_grad_(True)
    sig = torch.tensor(sigma, requires_grad=True)
    gl = torch.randn(shape, generator=gen)
    torch.manual_seed(seed + 7)
    with NoiseRecorder() as rec:
        y = sr.randomHeaviside.apply(x, S, sig)
    (y * gl).sum().backward()
    U = rec.drawn[0]

    z = torch.randn((N, H, W, K + 1), generator=gen) * 2 * gamma
    z[..., 0] = float("-inf")
    z[0, 0, 0, :] = float("-inf")
    z[0, 0, 0, -1] = 0.0
    z.requires_grad_(True)
    gam = torch.tensor(gamma, requires_grad=True)
    gw = torch.randn((N, H, W, K + 1), generator=gen)
    with NoiseRecorder() as rec:
        w = sa.randomArgmax.apply(z, S, gam, "gaussian", False)
    (w * gw).sum().backward()
    V = rec.drawn[0]
    out = dict(x=x.detach().numpy(), sigma=np.float32(sigma), S=np.int32(S), U=U.numpy(), grad_l=gl.numpy(),
               prob=y.detach().numpy(), grad_x=x.grad.numpy(), grad_sigma=sig.grad.numpy(),
               z=z.detach().numpy(), gamma=np.float32(gamma), V=V.numpy(), grad_w=gw.numpy(),
               weights=w.detach().numpy(), grad_z=z.grad.numpy(), grad_gamma=gam.grad.numpy())
    path = os.path.join(OUT, f"ops_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def run_cauchy_ops_case(sr, sa, name, *, shape, S, sigma, gamma, seed):
    """Cauchy-perturbed stand-alone Functions (ArctanRast: smoothrast.py:162-173, CauchyAgg: smoothagg.py:230-250;
    branches of randomHeaviside smoothrast.py:22-24,49-50 and randomArgmax smoothagg.py:25-27,57-63).  The
    reference draws through torch.distributions.Cauchy: the same draw is repeated after the same manual_seed."""
    gen = torch.Generator().manual_seed(seed)
    N, H, W, K = shape

    def cauchy(shape5, s):
        torch.manual_seed(s)
        m = torch.distributions.cauchy.Cauchy(torch.tensor([0.]), torch.tensor([1.]))
        return torch.clamp(m.sample(shape5).squeeze(-1), min=-1e7, max=1e7)

    x = (torch.rand(shape, generator=gen) * 2 - 1) * 3 * sigma
    x[..., -1] = 1.0
    x.requires_grad_(True)
    sig = torch.tensor(sigma, requires_grad=True)
    gl = torch.randn(shape, generator=gen)
    U = cauchy((S, N, H, W, K), seed + 7)
    torch.manual_seed(seed + 7)
    y = sr.randomHeaviside.apply(x, S, sig, "cauchy")
    (y * gl).sum().backward()
    z = torch.randn((N, H, W, K + 1), generator=gen) * 2 * gamma
    z[..., 0] = float("-inf")
    z.requires_grad_(True)
    gam = torch.tensor(gamma, requires_grad=True)
    gw = torch.randn((N, H, W, K + 1), generator=gen)
    V = cauchy((S, N, H, W, K + 1), seed + 8)
    torch.manual_seed(seed + 8)
    w = sa.randomArgmax.apply(z, S, gam, "cauchy", False)
    (w * gw).sum().backward()
    # the recorded draws are the reference's: forward must reproduce from them
    assert torch.equal(y.detach(), ((x.detach() + sigma * U) >= 0).float().mean(0))
    out = dict(x=x.detach().numpy(), sigma=np.float32(sigma), S=np.int32(S), U=U.numpy(), grad_l=gl.numpy(),
               prob=y.detach().numpy(), grad_x=x.grad.numpy(), grad_sigma=sig.grad.numpy(),
               z=z.detach().numpy(), gamma=np.float32(gamma), V=V.numpy(), grad_w=gw.numpy(),
               weights=w.detach().numpy(), grad_z=z.grad.numpy(), grad_gamma=gam.grad.numpy())
    path = os.path.join(OUT, f"ops_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def run_wovr_ops_case(sr, sa, name, *, shape, S, sigma, gamma, seed):
    """The estimators without control variates (randomHeaviside_wovr: smoothrast.py:61-108, randomArgmax_wovr:
    smoothagg.py:75-141, Gaussian branches) with an arbitrary upstream gradient and recorded noise."""
    gen = torch.Generator().manual_seed(seed)
    N, H, W, K = shape
    x = (torch.rand(shape, generator=gen) * 2 - 1) * 3 * sigma
    x[..., -1] = 1.0  # a far-inside entry: without the control variate its score sum does not vanish
    x[..., 0] = -1.0  # a far-outside one
    x.requires_grad_(True)
    sig = torch.tensor(sigma, requires_grad=True)
    gl = torch.randn(shape, generator=gen)
    torch.manual_seed(seed + 7)
    with NoiseRecorder() as rec:
        y = sr.randomHeaviside_wovr.apply(x, S, sig)
    (y * gl).sum().backward()
    U = rec.drawn[0]
    z = torch.randn((N, H, W, K + 1), generator=gen) * 2 * gamma
    z[..., 0] = float("-inf")
    z.requires_grad_(True)
    gam = torch.tensor(gamma, requires_grad=True)
    gw = torch.randn((N, H, W, K + 1), generator=gen)
    with NoiseRecorder() as rec:
        w = sa.randomArgmax_wovr.apply(z, S, gam, "gaussian", False)
    (w * gw).sum().backward()
    V = rec.drawn[0]
    out = dict(x=x.detach().numpy(), sigma=np.float32(sigma), S=np.int32(S), U=U.numpy(), grad_l=gl.numpy(),
               prob=y.detach().numpy(), grad_x=x.grad.numpy(), grad_sigma=sig.grad.numpy(),
               z=z.detach().numpy(), gamma=np.float32(gamma), V=V.numpy(), grad_w=gw.numpy(),
               weights=w.detach().numpy(), grad_z=z.grad.numpy(), grad_gamma=gam.grad.numpy())
    path = os.path.join(OUT, f"ops_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def run_cauchy_wovr_case(sr, name, *, shape, S, sigma, seed):
    """randomHeaviside_wovr with Cauchy noise (smoothrast.py:99-101): the Cauchy branch drops the control variate too
    (`maps * score`), unlike randomArgmax_wovr's (smoothagg.py:125-128, which keeps it: SURVEY B11)."""
    gen = torch.Generator().manual_seed(seed)
    N, H, W, K = shape
    x = (torch.rand(shape, generator=gen) * 2 - 1) * 3 * sigma
    x[..., -1] = 1.0
    x[..., 0] = -1.0
    x.requires_grad_(True)
    sig = torch.tensor(sigma, requires_grad=True)
    gl = torch.randn(shape, generator=gen)
    torch.manual_seed(seed + 7)
    m = torch.distributions.cauchy.Cauchy(torch.tensor([0.]), torch.tensor([1.]))
    U = torch.clamp(m.sample((S, N, H, W, K)).squeeze(-1), min=-1e7, max=1e7)
    torch.manual_seed(seed + 7)
    y = sr.randomHeaviside_wovr.apply(x, S, sig, "cauchy")
    (y * gl).sum().backward()
    assert torch.equal(y.detach(), ((x.detach() + sigma * U) >= 0).float().mean(0))
    out = dict(x=x.detach().numpy(), sigma=np.float32(sigma), S=np.int32(S), U=U.numpy(), grad_l=gl.numpy(),
               prob=y.detach().numpy(), grad_x=x.grad.numpy(), grad_sigma=sig.grad.numpy())
    path = os.path.join(OUT, f"ops_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def main():
    torch.set_num_threads(1)  # reduction order independent of the host
    rr, sr, sa = load_reference()
    if "--soft-only" in sys.argv:  # leave the committed Gaussian goldens untouched
        return soft_cases(rr, sr, sa)
    if "--cauchy-wovr-only" in sys.argv:
        return run_cauchy_wovr_case(sr, "cauchy_wovr", shape=(2, 3, 4, 6), S=12, sigma=1e-3, seed=43)
    if "--wovr-only" in sys.argv:
        return run_wovr_ops_case(sr, sa, "wovr", shape=(2, 3, 4, 6), S=12, sigma=1e-3, gamma=1e-2, seed=41)
    if "--cauchy-only" in sys.argv:
        return run_cauchy_ops_case(sr, sa, "cauchy", shape=(2, 3, 4, 6), S=12, sigma=1e-3, gamma=1e-2, seed=31)
    # 1: small, two batch elements with different depth planes, alpha != 1, non-white background
    run_case(rr, sr, sa, "small", N=2, H=6, W=6, K=5, S_r=8, S_a=8, sigma=1e-3, gamma=1e-2, alpha=1.3,
             seed=1, znear=[1.0, 0.5], zfar=[100.0, 50.0], background=(0.2, 0.5, 0.9))
    # 2: config-1-like (K=50, S=16) scaled to 8x8 pixels
    run_case(rr, sr, sa, "k50", N=1, H=8, W=8, K=50, S_r=16, S_a=16, sigma=1e-3, gamma=1e-2, alpha=1.0,
             seed=2, znear=[1.0], zfar=[100.0], background=(1.0, 1.0, 1.0), mean_valid=6.0)
    # 3: unequal sample counts (eval.py:150-151), odd K, non-square image, README smoothing
    run_case(rr, sr, sa, "uneven", N=1, H=5, W=7, K=7, S_r=16, S_a=8, sigma=1e-4, gamma=1e-3, alpha=1.0,
             seed=3, znear=[1.0], zfar=[100.0], background=(1.0, 1.0, 1.0))
    # 4: one batch element entirely empty; dense-in-K other element; S not a multiple of 4
    run_case(rr, sr, sa, "empty", N=2, H=4, W=4, K=6, S_r=6, S_a=10, sigma=1e-3, gamma=4e-2, alpha=0.7,
             seed=4, znear=[1.0, 1.0], zfar=[100.0, 100.0], background=(0.0, 0.0, 0.0),
             p_empty=0.0, mean_valid=6.0, all_empty_batch=1)
    run_ops_case(sr, sa, "small", shape=(2, 3, 4, 6), S=12, sigma=1e-3, gamma=1e-2, seed=11)
    soft_cases(rr, sr, sa)
    run_cauchy_ops_case(sr, sa, "cauchy", shape=(2, 3, 4, 6), S=12, sigma=1e-3, gamma=1e-2, seed=31)
    run_wovr_ops_case(sr, sa, "wovr", shape=(2, 3, 4, 6), S=12, sigma=1e-3, gamma=1e-2, seed=41)
    run_cauchy_wovr_case(sr, "cauchy_wovr", shape=(2, 3, 4, 6), S=12, sigma=1e-3, seed=43)


def soft_cases(rr, sr, sa):
    run_soft_case(rr, sr, sa, "small", N=2, H=6, W=6, K=5, sigma=1e-3, gamma=1e-2, alpha=1.3, seed=21,
                  znear=[1.0, 0.5], zfar=[100.0, 50.0], background=(0.2, 0.5, 0.9))
    run_soft_case(rr, sr, sa, "k50", N=1, H=8, W=8, K=50, sigma=2e-4, gamma=4e-2, alpha=1.0, seed=22,
                  znear=[1.0], zfar=[100.0], background=(1.0, 1.0, 1.0), mean_valid=6.0)
    run_soft_case(rr, sr, sa, "empty", N=2, H=4, W=5, K=7, sigma=1e-3, gamma=4e-3, alpha=0.7, seed=23,
                  znear=[1.0, 1.0], zfar=[100.0, 100.0], background=(0.0, 0.0, 0.0), p_empty=0.0, mean_valid=7.0,
                  all_empty_batch=1)


if __name__ == "__main__":
    main()
