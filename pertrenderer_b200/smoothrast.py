"""Coverage smoothing operators — the reference-facing mirror of ``randomras/smoothrast.py``.

Same class names, constructor arguments, attributes and ``torch.autograd.Function`` signatures as
the reference (``randomHeaviside`` smoothrast.py:12-59, ``SmoothRastBase`` :111-123, ``SoftRast``
:126-134, ``GaussianRast`` :136-147); the Monte-Carlo work runs in the sm_100a kernels behind the
C ABI (``pert_rast_fwd`` / ``pert_rast_bwd``), never in PyTorch ops.  When a ``GaussianRast`` is
paired with a ``GaussianAgg`` inside ``smooth_rgb_blend`` the two stages and the blend run as one
fused kernel instead (random_rasterizer.py).
"""

from __future__ import annotations

import torch
from torch.autograd import Function
from torch.nn import Module

from . import ops

_SUPPORTED = ("gaussian", "cauchy")


def _scalar(v) -> float:
    return float(v.detach()) if torch.is_tensor(v) else float(v)


class randomHeaviside(Function):
    """Perturbed Heaviside: ``mean_s 1[x + sigma*U_s >= 0]`` with the score-function backward and the
    control variate of the reference (smoothrast.py:15-59).

    ``forward(ctx, distances, nb_samples=1, noise_intensity=1e-1, noise_type="gaussian") -> map``
    ``backward(ctx, grad_l) -> (grad_dist, None, grad_sigma, None)``.
    ``noise_intensity`` may be a 0-dim (CPU) tensor and receives a gradient; as in the reference
    that gradient is ``sum(grad_dist)`` (smoothrast.py:57-58).
    """

    @staticmethod
    def forward(ctx, distances, nb_samples=1, noise_intensity=1e-1, noise_type="gaussian"):
        return randomHeaviside._forward(ctx, distances, nb_samples, noise_intensity, noise_type, 0)

    @staticmethod
    def _forward(ctx, distances, nb_samples, noise_intensity, noise_type, extra_flags):
        if noise_type not in _SUPPORTED:
            # the reference prints "noise type not implemented" and then dies on a NameError
            # (smoothrast.py:30-32); the logistic variant has no backward in the reference either
            raise ValueError(f"noise type {noise_type!r} not implemented (supported: {_SUPPORTED})")
        if distances.dim() != 4:
            raise ValueError("distances must be (N,H,W,K)")
        sigma = _scalar(noise_intensity)
        noise, _ = ops.current_explicit_noise()
        ops.refuse_device_seeds("randomHeaviside")
        seed = 0 if noise is not None else ops.draw_seed()
        # the Cauchy branch of randomHeaviside_wovr drops the control variate too (smoothrast.py:99-101), unlike
        # randomArgmax_wovr's (smoothagg.py:125-128)
        flags = ops.current_flags() | extra_flags | (ops.F_CAUCHY if noise_type == "cauchy" else 0)
        prob, rsum = ops.rast_forward(distances, int(nb_samples), sigma, seed=seed, noise=noise, flags=flags)
        ctx.save_for_backward(rsum)
        ctx.nb_samples, ctx.sigma = int(nb_samples), sigma
        ctx.sigma_like = noise_intensity if torch.is_tensor(noise_intensity) else None
        return prob

    @staticmethod
    def backward(ctx, grad_l):
        (rsum,) = ctx.saved_tensors
        grad_x, grad_sigma = ops.rast_backward(grad_l, rsum, ctx.nb_samples, ctx.sigma)
        grad_dist = grad_x if ctx.needs_input_grad[0] else None
        gs = None
        if ctx.sigma_like is not None and ctx.needs_input_grad[2]:
            gs = grad_sigma.to(device=ctx.sigma_like.device, dtype=ctx.sigma_like.dtype).reshape(ctx.sigma_like.shape)
        return grad_dist, None, gs, None


class randomHeaviside_wovr(randomHeaviside):
    """smoothrast.py:61-108: the same perturbed Heaviside WITHOUT the control variate in backward
    (``mean_s h_s U_s / sigma`` instead of ``mean_s (h_s - h0) U_s / sigma``): the paper's variance ablation
    (eval.py:152-154 "gaussian_wovr"), with Gaussian and with Cauchy noise (smoothrast.py:99-101)."""

    @staticmethod
    def forward(ctx, distances, nb_samples=1, noise_intensity=1e-1, noise_type="gaussian"):
        return randomHeaviside._forward(ctx, distances, nb_samples, noise_intensity, noise_type, ops.F_NO_VR)


class SmoothRastBase(Module):
    """smoothrast.py:111-123: holds ``sigma`` as a 0-dim CPU leaf tensor (not a Parameter, not moved
    by ``.to()``; replaced — and its ``.grad`` dropped — by ``update_smoothing``)."""

    def __init__(self, sigma=2e-4):
        super().__init__()
        self.sigma = torch.tensor(sigma, requires_grad=True)
        self.nb_samples = 1

    def update_smoothing(self, sigma):
        self.sigma = torch.tensor(sigma, requires_grad=True)

    def update_nb_samples(self, nb_samples):
        self.nb_samples = nb_samples


class SoftRast(SmoothRastBase):
    """SoftRas sigmoid coverage (smoothrast.py:126-134).  Default argument of the shaders; plain
    elementwise torch, not part of the perturbed hot path."""

    def __init__(self, sigma=2e-4):
        super().__init__(sigma)

    def rasterize(self, dists):
        return torch.sigmoid(dists.neg() / self.sigma)


class GaussianRast(SmoothRastBase):
    """Gaussian-perturbed coverage (smoothrast.py:136-147): ``rasterize(dists)`` is
    ``randomHeaviside.apply(-dists, nb_samples, sigma)``."""

    def __init__(self, nb_samples=16, sigma=2e-4):
        super().__init__(sigma)
        self.nb_samples = nb_samples

    def rasterize(self, dists):
        return randomHeaviside.apply(-dists, self.nb_samples, self.sigma)


class GaussianRast_wovr(SmoothRastBase):
    """smoothrast.py:149-160: ``GaussianRast`` on ``randomHeaviside_wovr``."""

    def __init__(self, nb_samples=16, sigma=2e-4):
        super().__init__(sigma)
        self.nb_samples = nb_samples

    def rasterize(self, dists):
        return randomHeaviside_wovr.apply(-dists, self.nb_samples, self.sigma)


class ArctanRast(SmoothRastBase):
    """Cauchy-perturbed coverage (smoothrast.py:162-173): ``randomHeaviside`` with ``"cauchy"`` noise, whose
    expectation is ``arctan(-dists/sigma)/pi + 1/2``."""

    def __init__(self, nb_samples=16, sigma=2e-4):
        super().__init__(sigma)
        self.nb_samples = nb_samples

    def rasterize(self, dists):
        return randomHeaviside.apply(-dists, self.nb_samples, self.sigma, "cauchy")


class AffineRast(SmoothRastBase):
    """Uniform-noise smoothing in closed form (smoothrast.py:175-185): ``clamp(-dists/sigma + 1/2, 0, 1)``."""

    def __init__(self, nb_samples=16, sigma=2e-4):
        super().__init__(sigma)
        self.nb_samples = nb_samples

    def rasterize(self, dists):
        return (dists.neg() / self.sigma + 0.5).clamp(min=0.0, max=1.0)


class HardRast:
    """No smoothing (smoothrast.py:187-194): ``1[-dists >= 0]``."""

    def rasterize(self, dists):
        return (dists <= 0).to(dists.dtype)
