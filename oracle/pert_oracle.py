"""CPU oracle for the perturbed shading hot path (TEST INFRASTRUCTURE ONLY).

This file is a closed-form restatement, in plain torch-CPU tensor arithmetic, of
what the reference computes on the path

    RandomSimpleShader.forward -> smooth_rgb_blend -> GaussianRast.rasterize
        -> GaussianAgg.aggregate -> blend            (and its autograd backward)

It is NOT part of the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The
product path (``pertrenderer_b200``) never routes through this module and fails
loudly when its CUDA library is missing.

Parity status: PINNED.  ``tests/golden/*.npz`` were produced by executing the
unmodified reference operators (``/root/reference/randomras``) with recorded
noise (``tests/golden/make_golden.py``); ``tests/test_oracle_golden.py`` checks
every function here against them (forward bit-exact, gradients <= 1e-6).

Every function cites the reference lines it restates (paths relative to
``/root/reference``).  Nothing here uses autograd: the backward pass is written
out by hand so it is an independent statement of the estimator.

Notation: P = N*H*W pixels, K faces per pixel, K1 = K+1 logits (the last one
is the background), S_r / S_a noise samples of the two stages.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

F32 = torch.float32


# ----------------------------------------------------------------------------
# noise
# ----------------------------------------------------------------------------
def draw_noise(shape4, nb_samples_rast, nb_samples_agg, generator=None):
    """Draw (U, V) the way the reference does: one standard-normal tensor of shape
    (S_r, N, H, W, K) for the coverage stage FIRST (randomras/smoothrast.py:21), then one of
    shape (S_a, N, H, W, K+1) for the aggregation stage (randomras/smoothagg.py:21)."""
    N, H, W, K = shape4
    U = torch.normal(torch.zeros((nb_samples_rast, N, H, W, K), dtype=F32), 1.0, generator=generator)
    V = torch.normal(torch.zeros((nb_samples_agg, N, H, W, K + 1), dtype=F32), 1.0, generator=generator)
    return U, V


# ----------------------------------------------------------------------------
# stage 1: perturbed Heaviside (coverage probability)
# ----------------------------------------------------------------------------
def random_heaviside_fwd(x, U, sigma):
    """randomras/smoothrast.py:15-37.

    x: (N,H,W,K) = -dists.  U: (S,N,H,W,K).  sigma: python float / 0-dim tensor.
    Returns (prob, h, h0): prob = mean_s h, h = 1[x + sigma*U >= 0] (S,...), h0 = 1[x >= 0].
    heaviside(.,values=1) is an inclusive >= (smoothrast.py:33-34); the perturbed value is a
    rounded multiply followed by a rounded add (smoothrast.py:32)."""
    sig = torch.as_tensor(sigma, dtype=F32)
    pert = x.unsqueeze(0) + sig * U
    h = (pert >= 0).to(F32)
    h0 = (x >= 0).to(F32)
    S = U.shape[0]
    prob = h.sum(dim=0) / S  # == mean(dim=0): the sum of 0/1 floats is exact
    return prob, h, h0


def random_heaviside_bwd(grad_l, h, h0, U, sigma, control_variate=True):
    """randomras/smoothrast.py:40-59 (gaussian branch); ``control_variate=False``: randomHeaviside_wovr.backward
    (smoothrast.py:90-108), the same estimator with ``h`` in place of ``h - h0`` (:95).

    grad_x = grad_l * mean_s[(h - h0) * U / sigma]  (:46,53,56).
    grad_sigma = sum(grad_x): the (U^2-1) expression of :47 is computed and then overwritten
    at :57-58, so the value that reaches sigma.grad is the plain sum of grad_x."""
    sig = torch.as_tensor(sigma, dtype=F32)
    S = U.shape[0]
    hh = h - h0.unsqueeze(0) if control_variate else h
    score = (hh * U / sig).sum(dim=0) / S
    grad_x = score * grad_l
    grad_sigma = grad_x.sum()
    return grad_x, grad_sigma


# ----------------------------------------------------------------------------
# stage 2: perturbed argmax (depth aggregation)
# ----------------------------------------------------------------------------
def build_logits(zbuf, zfar, znear, prob, mask, gamma, alpha, eps):
    """randomras/smoothagg.py:196-202 (GaussianAgg.aggregate up to the randomArgmax call).

    zi   = (zfar - z)/(zfar - znear) * mask                                  (:198)
    zmax = max_k zi, clamped from below at eps                               (:199)
    zeta_k = (gamma/alpha) * log(P_k) + zi_k - zmax,  zeta_K = eps - zmax    (:200-202)
    log(0) = -inf is kept (log_corrected.forward, :299-301)."""
    g = torch.as_tensor(gamma, dtype=F32)
    a = torch.as_tensor(alpha, dtype=F32)
    maskf = mask.to(F32)
    zi = (zfar - zbuf) / (zfar - znear) * maskf
    zi_max_raw, zi_arg = torch.max(zi, dim=-1)
    zmax = zi_max_raw.clamp(min=eps)[..., None]
    logp = prob.log()
    zeta_faces = (g / a) * logp + zi - zmax
    bg = torch.ones(zbuf.shape[:-1] + (1,), dtype=F32) * eps - zmax
    zeta = torch.cat((zeta_faces, bg), dim=-1)
    return zeta, dict(zi=zi, zi_max_raw=zi_max_raw, zi_arg=zi_arg, zmax=zmax, logp=logp)


def random_argmax_fwd(zeta, V, gamma):
    """randomras/smoothagg.py:13-42 (gaussian branch).

    a_s = argmax_j(zeta_j + gamma*V_sj) (:33-34), a_0 = argmax_j zeta_j (:37).
    Returns (weights, a_s, a_0) with weights = mean_s onehot(a_s) (:41)."""
    g = torch.as_tensor(gamma, dtype=F32)
    S = V.shape[0]
    K1 = zeta.shape[-1]
    a_s = torch.max(zeta.unsqueeze(0) + g * V, dim=-1).indices  # (S,N,H,W)
    a_0 = torch.max(zeta, dim=-1).indices  # (N,H,W)
    hist = torch.zeros(zeta.shape, dtype=F32)
    hist.scatter_add_(-1, a_s.permute(1, 2, 3, 0), torch.ones(a_s.permute(1, 2, 3, 0).shape, dtype=F32))
    weights = hist / S
    return weights, a_s, a_0


def random_argmax_bwd(grad_l, a_s, a_0, V, gamma, control_variate=True):
    """randomras/smoothagg.py:45-73 (gaussian branch); ``control_variate=False``: randomArgmax_wovr.backward
    (smoothagg.py:112-141, gaussian branch :118-123), where c_s = <grad_l, onehot(a_s)> without the a_0 term.

    c_s = <grad_l, onehot(a_s) - onehot(a_0)>                      (:51)
    grad_zeta_j = mean_s c_s * V_sj / gamma, for ALL j             (:52,71)
    grad_gamma  = mean_s sum_pixels c_s * (||V_s||^2 - 1) / gamma  (:54-56,72)
    (the norm runs over all K+1 coordinates, and the constant is 1, not K+1)."""
    g = torch.as_tensor(gamma, dtype=F32)
    S = V.shape[0]
    gl_s = grad_l.unsqueeze(0).expand(S, *grad_l.shape)
    g_sel = torch.gather(gl_s, -1, a_s.unsqueeze(-1)).squeeze(-1)  # (S,N,H,W)
    g_ref = torch.gather(grad_l, -1, a_0.unsqueeze(-1)).squeeze(-1)  # (N,H,W)
    c = g_sel - g_ref.unsqueeze(0) if control_variate else g_sel
    grad_zeta = (c.unsqueeze(-1) * V / g).sum(dim=0) / S
    nsq = (V * V).sum(dim=-1)
    grad_gamma = (c * (nsq - 1.0) / g).sum(dim=(1, 2, 3)).sum() / S
    return grad_zeta, grad_gamma


# ----------------------------------------------------------------------------
# the whole shader: forward
# ----------------------------------------------------------------------------
@dataclass
class ShadeState:
    """Everything forward produced; backward consumes it."""
    image: torch.Tensor  # (N,H,W,4)
    prob: torch.Tensor  # (N,H,W,K) masked coverage probability P
    weights: torch.Tensor  # (N,H,W,K1)
    counts: torch.Tensor  # (N,H,W,K) int32: number of samples with h=1 (unmasked)
    a_s: torch.Tensor  # (S_a,N,H,W) int64 winners
    a_0: torch.Tensor  # (N,H,W) int64 unperturbed winner
    h: torch.Tensor
    h0: torch.Tensor
    aux: dict
    zeta: torch.Tensor


def _as_background(background):
    if torch.is_tensor(background):
        return background.to(F32).reshape(3)
    return torch.tensor(tuple(background), dtype=F32)


def _as_depth_plane(v, N):
    """znear / zfar arrive either as python floats (smooth_rgb_blend's defaults,
    randomras/random_rasterizer.py:35) or as (N,1,1,1) tensors (:172-173)."""
    if torch.is_tensor(v):
        return v.to(F32).reshape(-1, 1, 1, 1)
    return torch.full((1, 1, 1, 1), float(v), dtype=F32)


def shade_forward(pix_to_face, zbuf, dists, colors, background, znear, zfar,
                  sigma, gamma, alpha, eps, U, V) -> ShadeState:
    """randomras/random_rasterizer.py:34-56 with GaussianRast (smoothrast.py:144-147) and
    GaussianAgg (smoothagg.py:196-205), fed with explicit noise U, V."""
    N = pix_to_face.shape[0]
    bg = _as_background(background)
    zn, zf = _as_depth_plane(znear, N), _as_depth_plane(zfar, N)
    mask = pix_to_face >= 0  # random_rasterizer.py:46
    p_hat, h, h0 = random_heaviside_fwd(-dists, U, sigma)  # smoothrast.py:146
    prob = p_hat * mask  # :47
    alpha_chan = torch.prod(1.0 - prob, dim=-1)  # :48
    zeta, aux = build_logits(zbuf, zf, zn, prob, mask, gamma, alpha, eps)
    weights, a_s, a_0 = random_argmax_fwd(zeta, V, gamma)
    wz, wb = weights[..., :-1], weights[..., -1:]  # :50
    rgb = (wz[..., None] * colors).sum(dim=-2) + wb * bg  # :51-53
    image = torch.ones(pix_to_face.shape[:3] + (4,), dtype=F32)
    image[..., :3] = rgb
    image[..., 3] = 1.0 - alpha_chan  # :54
    counts = h.sum(dim=0).to(torch.int32)
    aux = dict(aux, mask=mask, zn=zn, zf=zf, bg=bg)
    return ShadeState(image, prob, weights, counts, a_s, a_0, h, h0, aux, zeta)


# ----------------------------------------------------------------------------
# the whole shader: backward (hand-derived)
# ----------------------------------------------------------------------------
def _prod_excluding_self(t):
    """prod_{l != k} t_l along the last dim, exact in the presence of zeros."""
    K = t.shape[-1]
    left = torch.ones_like(t)
    right = torch.ones_like(t)
    for k in range(1, K):
        left[..., k] = left[..., k - 1] * t[..., k - 1]
    for k in range(K - 2, -1, -1):
        right[..., k] = right[..., k + 1] * t[..., k + 1]
    return left * right


def _finish_backward(prob, weights, aux, grad_image, zbuf, colors, gamma, alpha, eps, grad_zeta, grad_gamma_score,
                     coverage_score):
    """Chain rule from (grad_zeta, coverage score) to the inputs — everything in backward that is not
    a sum over noise samples.  ``coverage_score`` = mean_s (h-h0) U / sigma (smoothrast.py:46,53)."""
    g = torch.as_tensor(gamma, dtype=F32)
    a = torch.as_tensor(alpha, dtype=F32)
    maskf = aux["mask"].to(F32)
    K = zbuf.shape[-1]
    G_rgb, G_a = grad_image[..., :3], grad_image[..., 3]
    grad_colors = weights[..., :-1, None] * G_rgb[..., None, :]

    # zeta = cat(gal*logP + zi - zmax, eps - zmax)
    gz_faces = grad_zeta[..., :K]
    g_zmax = -grad_zeta.sum(dim=-1)  # both the K face logits and the background subtract zmax
    # zmax = clamp(max_k zi, min=eps): the gradient passes where the raw max is >= eps and goes
    # to the single index torch.max returned
    passes = (aux["zi_max_raw"] >= eps).to(F32)
    g_zi = gz_faces.clone()
    g_zi.scatter_add_(-1, aux["zi_arg"][..., None], (g_zmax * passes)[..., None])
    grad_zbuf = -(g_zi * maskf) / (aux["zf"] - aux["zn"])

    # prod_corrected(gamma/alpha, logP): scalar side nansum with inf -> 0, tensor side nan -> 0
    logp = aux["logp"]
    logp_fin = torch.where(torch.isinf(logp), torch.zeros_like(logp), logp)
    q = (logp_fin * gz_faces).nansum()
    grad_gamma = grad_gamma_score + q / a
    grad_alpha = -q * g / (a * a)
    g_logp = (g / a) * gz_faces
    g_logp = torch.where(torch.isnan(g_logp), torch.zeros_like(g_logp), g_logp)
    # log_corrected: 1/x with inf -> 0
    inv = 1.0 / prob
    inv = torch.where(torch.isinf(inv), torch.zeros_like(inv), inv)
    g_prob = inv * g_logp
    # alpha channel: image[...,3] = 1 - prod_k (1 - P_k)
    g_prob = g_prob + G_a[..., None] * _prod_excluding_self(1.0 - prob)
    # P = p_hat * mask
    g_phat = g_prob * maskf
    grad_x = coverage_score * g_phat  # smoothrast.py:56
    grad_sigma = grad_x.sum()  # smoothrast.py:57-58
    return dict(dists=-grad_x, zbuf=grad_zbuf, colors=grad_colors, sigma=grad_sigma, gamma=grad_gamma,
                alpha=grad_alpha, zeta=grad_zeta, prob=g_prob)


def _grad_weights(colors, grad_image, bg):
    """d loss / d w_j = <G_rgb, colour_j> (blend backward, random_rasterizer.py:50-53)."""
    G_rgb = grad_image[..., :3]
    g_face = (colors * G_rgb[..., None, :]).sum(dim=-1)
    g_bg = (G_rgb * bg).sum(dim=-1, keepdim=True)
    return torch.cat((g_face, g_bg), dim=-1)  # (N,H,W,K1)


def shade_backward(st: ShadeState, grad_image, zbuf, colors, sigma, gamma, alpha, eps, U, V):
    """Gradients of <grad_image, image> w.r.t. dists, zbuf, colors, sigma, gamma, alpha.

    Follows, in reverse, random_rasterizer.py:50-54 (blend), smoothagg.py:45-73
    (randomArgmax.backward), :303-311 / :325-337 (log_corrected / prod_corrected backward),
    the autograd rules of cat / max / clamp / div used at :198-202, torch.prod (:48 of
    random_rasterizer.py) and smoothrast.py:40-59."""
    sig = torch.as_tensor(sigma, dtype=F32)
    g = torch.as_tensor(gamma, dtype=F32)
    grad_w = _grad_weights(colors, grad_image, st.aux["bg"])
    grad_zeta, grad_gamma = random_argmax_bwd(grad_w, st.a_s, st.a_0, V, g)
    S = U.shape[0]
    score = ((st.h - st.h0.unsqueeze(0)) * U / sig).sum(dim=0) / S
    return _finish_backward(st.prob, st.weights, st.aux, grad_image, zbuf, colors, gamma, alpha, eps, grad_zeta,
                            grad_gamma, score)


# ----------------------------------------------------------------------------
# noise-sample shards: per-shard sums, and the finish from reduced sums
# ----------------------------------------------------------------------------
def rast_shard_sums(dists, U_shard, sigma):
    """Coverage phase on a shard of samples: (hit counts, sum_s (h-h0) U), both (N,H,W,K) float."""
    _, h, h0 = random_heaviside_fwd(-dists, U_shard, sigma)
    return h.sum(dim=0), ((h - h0.unsqueeze(0)) * U_shard).sum(dim=0)


def logits_from_counts(pix_to_face, zbuf, counts, S_rast, znear, zfar, gamma, alpha, eps):
    """Logits from ALL-shard hit counts (the non-linear log needs the full-S mean)."""
    N = pix_to_face.shape[0]
    mask = pix_to_face >= 0
    prob = (counts / S_rast) * mask
    zn, zf = _as_depth_plane(znear, N), _as_depth_plane(zfar, N)
    zeta, aux = build_logits(zbuf, zf, zn, prob, mask, gamma, alpha, eps)
    return zeta, prob, dict(aux, mask=mask, zn=zn, zf=zf)


def argmax_shard(zeta, V_shard, gamma):
    """Argmax phase on a shard: (winner histogram (N,H,W,K1) int32, a_s, a_0)."""
    w, a_s, a_0 = random_argmax_fwd(zeta, V_shard, gamma)
    return (w * V_shard.shape[0]).round().to(torch.int32), a_s, a_0


def blend_from_hist(hist, S_agg, prob, colors, background):
    weights = hist.to(F32) / S_agg
    bg = _as_background(background)
    rgb = (weights[..., :-1, None] * colors).sum(dim=-2) + weights[..., -1:] * bg
    alpha_chan = torch.prod(1.0 - prob, dim=-1)
    return torch.cat((rgb, (1.0 - alpha_chan)[..., None]), dim=-1), weights


def argmax_score_sums(grad_w, a_s, a_0, V_shard):
    """Backward sample phase on a shard, packed (N,H,W,K1+2):
    [..., :K1] = sum_s c_s V_sj, [..., K1] = sum_s c_s ||V_s||^2, [..., K1+1] = sum_s c_s."""
    S = V_shard.shape[0]
    gl_s = grad_w.unsqueeze(0).expand(S, *grad_w.shape)
    c = torch.gather(gl_s, -1, a_s.unsqueeze(-1)).squeeze(-1) - torch.gather(grad_w, -1, a_0.unsqueeze(-1)).squeeze(-1)
    acc = (c.unsqueeze(-1) * V_shard).sum(dim=0)
    t2 = (c * (V_shard * V_shard).sum(dim=-1)).sum(dim=0)
    return torch.cat((acc, t2[..., None], c.sum(dim=0)[..., None]), dim=-1)


def shade_backward_from_sums(prob, weights, aux, grad_image, zbuf, colors, sigma, gamma, alpha, eps, S_rast, S_agg,
                             rsum, packed):
    """Finish phase from all-shard sums: grad_zeta = acc/(S gamma), score term of gamma =
    sum_pixels (t2 - csum)/(S gamma) (smoothagg.py:52-56,71-72), coverage score = rsum/(S sigma)."""
    K1 = zbuf.shape[-1] + 1
    g = float(gamma)
    grad_zeta = packed[..., :K1] / (S_agg * g)
    grad_gamma = ((packed[..., K1] - packed[..., K1 + 1]) / (S_agg * g)).sum()
    score = rsum / (S_rast * float(sigma))
    aux = dict(aux, bg=_as_background(aux["bg"]) if "bg" in aux else None)
    return _finish_backward(prob, weights, aux, grad_image, zbuf, colors, gamma, alpha, eps, grad_zeta, grad_gamma,
                            score)


def shade_fwd_bwd(pix_to_face, zbuf, dists, colors, background, znear, zfar,
                  sigma, gamma, alpha, eps, U, V, grad_image):
    st = shade_forward(pix_to_face, zbuf, dists, colors, background, znear, zfar,
                       sigma, gamma, alpha, eps, U, V)
    grads = shade_backward(st, grad_image, zbuf, colors, sigma, gamma, alpha, eps, U, V)
    return st, grads


# ----------------------------------------------------------------------------
# closed-form known answers used by the Philox-mode tests
# ----------------------------------------------------------------------------
# ----------------------------------------------------------------------------
# the SoftRas pair: SoftRast + SoftAgg (the shaders' default operators), deterministic
# ----------------------------------------------------------------------------
def soft_shade_fwd_bwd(pix_to_face, zbuf, dists, colors, background, znear, zfar, sigma, gamma, alpha, eps,
                       grad_image=None):
    """smooth_rgb_blend (randomras/random_rasterizer.py:34-56) with SoftRast.rasterize
    (randomras/smoothrast.py:132-134: sigmoid(-dists/sigma)) and SoftAgg.aggregate
    (randomras/smoothagg.py:172-182: softmax((1/gamma) * logits)), forward and hand-written backward
    (log_corrected / prod_corrected rules of smoothagg.py:292-337).

    Returns (image, prob, weights) or, with ``grad_image``, also the gradient dict."""
    N, H, W, K = pix_to_face.shape
    sig = torch.as_tensor(sigma, dtype=F32)
    g = torch.as_tensor(gamma, dtype=F32)
    a = torch.as_tensor(alpha, dtype=F32)
    bg = _as_background(background)
    zn, zf = _as_depth_plane(znear, N), _as_depth_plane(zfar, N)
    mask = pix_to_face >= 0
    maskf = mask.to(F32)
    p = torch.sigmoid(-dists / sig)  # smoothrast.py:133
    prob = p * maskf  # random_rasterizer.py:47
    alpha_px = 1.0 - torch.prod(1.0 - prob, dim=-1)  # :48, :54
    zeta, aux = build_logits(zbuf, zf, zn, prob, mask, gamma, alpha, eps)  # smoothagg.py:174-180
    inv_g = 1.0 / g
    y = inv_g * zeta  # prod_corrected(1/gamma, z_map), smoothagg.py:181
    weights = torch.softmax(y, dim=-1)
    rgb = (weights[..., :K, None] * colors).sum(dim=-2) + weights[..., K:] * bg  # random_rasterizer.py:50-53
    image = torch.cat((rgb, alpha_px[..., None]), dim=-1)
    if grad_image is None:
        return image, prob, weights

    G_rgb, G_a = grad_image[..., :3], grad_image[..., 3]
    grad_colors = weights[..., :K, None] * G_rgb[..., None, :]
    gw = _grad_weights(colors, grad_image, bg)  # d/dw_j
    gy = weights * (gw - (weights * gw).sum(dim=-1, keepdim=True))  # softmax backward
    # prod_corrected(1/gamma, zeta): scalar side ignores infinite logits, tensor side nan -> 0
    zeta_fin = torch.where(torch.isinf(zeta), torch.zeros_like(zeta), zeta)
    g_inv_g = (zeta_fin * gy).nansum()
    gzeta = inv_g * gy
    gzeta = torch.where(torch.isnan(gzeta), torch.zeros_like(gzeta), gzeta)
    grad_gamma = -g_inv_g / (g * g)
    gz_faces = gzeta[..., :K]
    g_zmax = -gzeta.sum(dim=-1)
    passes = (aux["zi_max_raw"] >= eps).to(F32)
    g_zi = gz_faces.clone()
    g_zi.scatter_add_(-1, aux["zi_arg"][..., None], (g_zmax * passes)[..., None])
    grad_zbuf = -(g_zi * maskf) / (zf - zn)
    logp = aux["logp"]
    logp_fin = torch.where(torch.isinf(logp), torch.zeros_like(logp), logp)
    q = (logp_fin * gz_faces).nansum()
    grad_gamma = grad_gamma + q / a
    grad_alpha = -q * g / (a * a)
    g_logp = (g / a) * gz_faces
    g_logp = torch.where(torch.isnan(g_logp), torch.zeros_like(g_logp), g_logp)
    inv = 1.0 / prob
    inv = torch.where(torch.isinf(inv), torch.zeros_like(inv), inv)
    g_prob = inv * g_logp + G_a[..., None] * _prod_excluding_self(1.0 - prob)
    g_p = g_prob * maskf
    g_u = g_p * p * (1.0 - p)  # sigmoid backward, u = -dists / sigma
    grad_dists = -g_u / sig
    grad_sigma = (g_u * dists / (sig * sig)).sum()
    return image, prob, weights, dict(dists=grad_dists, zbuf=grad_zbuf, colors=grad_colors, sigma=grad_sigma,
                                      gamma=grad_gamma, alpha=grad_alpha)


def normal_cdf(t):
    return 0.5 * (1.0 + torch.erf(torch.as_tensor(t, dtype=torch.float64) / 2.0 ** 0.5))


def normal_pdf(t):
    t = torch.as_tensor(t, dtype=torch.float64)
    return torch.exp(-0.5 * t * t) / (2.0 * torch.pi) ** 0.5


def expected_coverage(dists, sigma):
    """E[p_hat_k] = Phi(x_k / sigma), x = -dists (estimator definition, smoothrast.py:32-36)."""
    return normal_cdf(-dists.double() / float(sigma))


def expected_coverage_score(dists, sigma):
    """E[(h-h0) U]/sigma = phi(x/sigma)/sigma."""
    return normal_pdf(-dists.double() / float(sigma)) / float(sigma)


def expected_two_way_weight(zeta0, zeta1, gamma):
    """Two live logits: P[zeta0 + g V0 > zeta1 + g V1] = Phi((zeta0 - zeta1)/(g sqrt 2))."""
    return normal_cdf((zeta0 - zeta1) / (float(gamma) * 2.0 ** 0.5))
