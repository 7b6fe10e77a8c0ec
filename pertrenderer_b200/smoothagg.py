"""Depth aggregation operators — the reference-facing mirror of ``randomras/smoothagg.py``.

Same names, constructor arguments, attributes and autograd signatures as the reference
(``randomArgmax`` smoothagg.py:10-73, ``SmoothAggBase`` :145-163, ``SoftAgg`` :165-182,
``GaussianAgg`` :185-205, ``log_corrected`` / ``prod_corrected`` :292-337).  The perturbed argmax
runs in the sm_100a kernels (``pert_argmax_fwd`` / ``pert_argmax_bwd``); paired with a
``GaussianRast`` inside ``smooth_rgb_blend`` the whole shader runs as one fused kernel instead.
"""

from __future__ import annotations

import torch
from torch.autograd import Function
from torch.nn import Module

from . import ops

_SUPPORTED = ("gaussian", "cauchy", "uniform", "gumbel")
_FORWARD_ONLY = {"uniform": ops.F_UNIFORM, "gumbel": ops.F_GUMBEL}  # no backward in the reference (smoothagg.py:64-67)


def _scalar(v) -> float:
    return float(v.detach()) if torch.is_tensor(v) else float(v)


class randomArgmax(Function):
    """Perturbed argmax: ``mean_s onehot(argmax_j(z_j + gamma*V_sj))`` (smoothagg.py:13-42) with the
    variance-reduced score-function backward of smoothagg.py:45-73.

    ``forward(ctx, z, nb_samples=1, noise_intensity=1e-1, noise_type="gaussian", fixed_noise=False)``
    ``backward(ctx, grad_l) -> (grad_z, None, grad_gamma, None, None)``.
    Backward regenerates the noise from the Philox counter; only one winner index per pixel and
    sample is saved (the reference saves the one-hot and noise tensors, (S,N,H,W,K+1) each).
    """

    @staticmethod
    def forward(ctx, z, nb_samples=1, noise_intensity=1e-1, noise_type="gaussian", fixed_noise=False):
        return randomArgmax._forward(ctx, z, nb_samples, noise_intensity, noise_type, fixed_noise, 0)

    @staticmethod
    def _forward(ctx, z, nb_samples, noise_intensity, noise_type, fixed_noise, extra_flags):
        if noise_type not in _SUPPORTED:
            # reference: gumbel / cauchy / uniform exist forward-only or as baselines
            # (smoothagg.py:22-32, 64-69); outside the B200 path
            raise ValueError(f"noise type {noise_type!r} not implemented (supported: {_SUPPORTED})")
        if z.dim() != 4:
            raise ValueError("z must be (N,H,W,K+1)")
        if fixed_noise:
            torch.manual_seed(1)  # smoothagg.py:18-19: the reference reseeds the GLOBAL generator
        gamma = _scalar(noise_intensity)
        _, noise = ops.current_explicit_noise()
        ops.refuse_device_seeds("randomArgmax")
        seed = 0 if noise is not None else ops.draw_seed()
        flags = ops.current_flags() | (ops.F_CAUCHY if noise_type == "cauchy" else
                                       _FORWARD_ONLY.get(noise_type, extra_flags))
        weights, winners = ops.argmax_forward(z, int(nb_samples), gamma, seed=seed, noise=noise, flags=flags)
        ctx.save_for_backward(z.detach(), winners)
        ctx.cfg = (int(nb_samples), gamma, seed, noise, flags)
        ctx.gamma_like = noise_intensity if torch.is_tensor(noise_intensity) else None
        return weights

    @staticmethod
    def backward(ctx, grad_l):
        z, winners = ctx.saved_tensors
        S, gamma, seed, noise, flags = ctx.cfg
        if flags & (ops.F_UNIFORM | ops.F_GUMBEL):
            # the reference prints "noise_type not implemented" and dies on a None (smoothagg.py:64-71)
            raise RuntimeError("randomArgmax: uniform / gumbel noise has no backward (forward-only in the reference too)")
        gz, gg = ops.argmax_backward(grad_l, z, winners, S, gamma, seed=seed, noise=noise, flags=flags)
        grad_gamma = None
        if ctx.gamma_like is not None and ctx.needs_input_grad[2]:
            grad_gamma = gg.to(device=ctx.gamma_like.device, dtype=ctx.gamma_like.dtype).reshape(ctx.gamma_like.shape)
        return (gz if ctx.needs_input_grad[0] else None), None, grad_gamma, None, None


class randomArgmax_wovr(randomArgmax):
    """smoothagg.py:75-141: the perturbed argmax WITHOUT variance reduction in backward: ``c_s = <g, onehot(a_s)>``
    instead of ``<g, onehot(a_s) - onehot(a_0)>`` (Gaussian noise; the reference's Cauchy branch of this class keeps
    the control variate, smoothagg.py:123-128, i.e. it is the plain operator)."""

    @staticmethod
    def forward(ctx, z, nb_samples=1, noise_intensity=1e-1, noise_type="gaussian", fixed_noise=False):
        return randomArgmax._forward(ctx, z, nb_samples, noise_intensity, noise_type, fixed_noise, ops.F_NO_VR)


class log_corrected(Function):
    """``log`` whose backward yields 0 (not inf/nan) where the input is 0 (smoothagg.py:292-311)."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.log(x)

    @staticmethod
    def backward(ctx, grad_l):
        if not ctx.needs_input_grad[0]:
            return None
        (x,) = ctx.saved_tensors
        recip = torch.reciprocal(x)
        recip = recip.masked_fill(torch.isinf(recip), 0.0)
        return recip * grad_l


class prod_corrected(Function):
    """``x * y`` for a scalar ``x`` and a tensor ``y`` that may hold -inf: the scalar's gradient
    ignores the infinite entries and the tensor's gradient maps nan to 0 (smoothagg.py:314-337)."""

    @staticmethod
    def forward(ctx, x, y):
        ctx.save_for_backward(x, y)
        return x * y

    @staticmethod
    def backward(ctx, grad_l):
        x, y = ctx.saved_tensors
        gx = gy = None
        if ctx.needs_input_grad[0]:
            gx = (y.masked_fill(torch.isinf(y), 0.0) * grad_l).nansum()
        if ctx.needs_input_grad[1]:
            gy = torch.nan_to_num(x * grad_l, nan=0.0, posinf=float("inf"), neginf=float("-inf"))
        return gx, gy


class SmoothAggBase(Module):
    """smoothagg.py:145-163: ``gamma`` and ``alpha`` are 0-dim CPU leaf tensors (replaced by
    ``update_smoothing``); ``eps`` is the background logit offset."""

    def __init__(self, gamma, alpha, eps, nb_samples=1):
        super().__init__()
        self.gamma = torch.tensor(gamma, requires_grad=True)
        self.alpha = torch.tensor(alpha, requires_grad=True)
        self.nb_samples = nb_samples
        self.eps = eps

    def update_smoothing(self, gamma=4e-2, alpha=1.):
        self.gamma = torch.tensor(gamma, requires_grad=True)
        self.alpha = torch.tensor(alpha, requires_grad=True)

    def update_nb_samples(self, nb_samples):
        self.nb_samples = nb_samples

    def _logits(self, zbuf, zfar, znear, prob_map, mask):
        """The K+1 logits of smoothagg.py:198-202 (shared by every aggregation rule)."""
        z_inv = (zfar - zbuf) / (zfar - znear) * mask
        z_inv_max = z_inv.max(dim=-1, keepdim=True).values.clamp(min=self.eps)
        scaled_log = prod_corrected.apply(self.gamma / self.alpha, log_corrected.apply(prob_map))
        faces = scaled_log + z_inv - z_inv_max
        background = torch.full_like(z_inv_max, 1.0) * self.eps - z_inv_max
        return torch.cat((faces, background), dim=-1)


class SoftAgg(SmoothAggBase):
    """SoftRas softmax aggregation (smoothagg.py:165-182).  Default argument of the shaders; plain
    torch, not part of the perturbed hot path."""

    def __init__(self, gamma=4e-2, alpha=1., eps=1e-10):
        super().__init__(gamma, alpha, eps)

    def aggregate(self, zbuf, zfar, znear, prob_map, mask):
        z_map = self._logits(zbuf, zfar, znear, prob_map, mask)
        return torch.softmax(prod_corrected.apply(1. / self.gamma, z_map), dim=-1)


class GaussianAgg(SmoothAggBase):
    """Gaussian-perturbed aggregation (smoothagg.py:185-205)."""

    def __init__(self, nb_samples=16, gamma=4e-2, alpha=1., eps=1e-10, fixed_noise=False):
        super().__init__(gamma, alpha, eps, nb_samples)
        self.fixed_noise = fixed_noise

    def aggregate(self, zbuf, zfar, znear, prob_map, mask):
        z_map = self._logits(zbuf, zfar, znear, prob_map, mask)
        return randomArgmax.apply(z_map, self.nb_samples, self.gamma, "gaussian", self.fixed_noise)


class GaussianAgg_wovr(SmoothAggBase):
    """smoothagg.py:207-228: ``GaussianAgg`` on ``randomArgmax_wovr`` (no variance reduction in backward)."""

    def __init__(self, nb_samples=16, gamma=4e-2, alpha=1., eps=1e-10, fixed_noise=False):
        super().__init__(gamma, alpha, eps, nb_samples)
        self.fixed_noise = fixed_noise

    def aggregate(self, zbuf, zfar, znear, prob_map, mask):
        z_map = self._logits(zbuf, zfar, znear, prob_map, mask)
        return randomArgmax_wovr.apply(z_map, self.nb_samples, self.gamma, "gaussian", self.fixed_noise)


class CauchyAgg(SmoothAggBase):
    """Cauchy-perturbed aggregation (smoothagg.py:230-250): the same logits, ``randomArgmax`` with
    ``"cauchy"`` noise."""

    def __init__(self, nb_samples=16, gamma=4e-2, alpha=1., eps=1e-10, fixed_noise=False):
        super().__init__(gamma, alpha, eps, nb_samples)
        self.fixed_noise = fixed_noise

    def aggregate(self, zbuf, zfar, znear, prob_map, mask):
        z_map = self._logits(zbuf, zfar, znear, prob_map, mask)
        return randomArgmax.apply(z_map, self.nb_samples, self.gamma, "cauchy", self.fixed_noise)


class UniformAgg(SmoothAggBase):
    """smoothagg.py:252-272: the same logits, ``randomArgmax`` with uniform noise on [-1/2, 1/2).  Forward only: the
    reference has no backward for this noise (smoothagg.py:64-65)."""

    def __init__(self, nb_samples=16, gamma=4e-2, alpha=1., eps=1e-10, fixed_noise=False):
        super().__init__(gamma, alpha, eps, nb_samples)
        self.fixed_noise = fixed_noise

    def aggregate(self, zbuf, zfar, znear, prob_map, mask):
        z_map = self._logits(zbuf, zfar, znear, prob_map, mask)
        return randomArgmax.apply(z_map, self.nb_samples, self.gamma, "uniform", self.fixed_noise)


class HardAgg:
    """No smoothing (smoothagg.py:274-289): one-hot of the arg-max logit, the coverage term weighted by 1e-6."""

    def __init__(self, eps=1e-10):
        self.eps = eps

    def aggregate(self, zbuf, zfar, znear, prob_map, mask):
        z_inv = (zfar - zbuf) / (zfar - znear) * mask
        z_inv_max = z_inv.max(dim=-1, keepdim=True).values.clamp(min=self.eps)
        faces = 1e-6 * torch.log(prob_map) + z_inv - z_inv_max
        z_map = torch.cat((faces, self.eps - z_inv_max), dim=-1)
        return torch.zeros_like(z_map).scatter_(-1, z_map.argmax(dim=-1, keepdim=True), 1.0)
