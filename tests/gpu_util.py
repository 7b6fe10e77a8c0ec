"""Helpers shared by the -m gpu tests: run the CUDA path through the C ABI on golden / synthetic
inputs, and the CPU oracle on the same inputs."""

import torch

from oracle import pert_oracle as O
from pertrenderer_b200 import ops

DEV = "cuda"


def problem_from_case(g, explicit=True, seed_rast=0, seed_agg=0, flags=0, **over):
    """ShadeProblem on cuda:0 from a golden dict (tests/golden) or any dict with the same keys."""
    kw = dict(
        pix_to_face=g["pix_to_face"].to(DEV), zbuf=g["zbuf"].to(DEV), dists=g["dists"].to(DEV),
        colors=g["colors"].to(DEV), znear=g["znear"].to(DEV), zfar=g["zfar"].to(DEV),
        background=tuple(float(v) for v in g["background"]), sigma=float(g["sigma"]), gamma=float(g["gamma"]),
        alpha=float(g["alpha"]), eps=float(g["eps"]), S_rast=int(g["S_r"]), S_agg=int(g["S_a"]),
        seed_rast=seed_rast, seed_agg=seed_agg, flags=flags,
        noise_rast=g["U"].to(DEV) if explicit else None, noise_agg=g["V"].to(DEV) if explicit else None)
    kw.update(over)
    return ops.ShadeProblem(**kw)


def counts_u16(saved):
    return saved.counts.to(torch.int32) & 0xFFFF


def winners_long(saved):
    """Winner of every sample; rows of inactive pixels (never written by the kernel) are a0."""
    return saved.winners_full()


def run_cuda(pr, grad_image, need_colors=True):
    """One fused forward + backward exactly as the product launches them (no phase flags, no global
    histogram: the production kernel instantiation); the winner histogram is rebuilt from the winners."""
    image, saved = ops.shade_forward(pr)
    gd, gz, gc, scal = ops.shade_backward(pr, saved, grad_image.to(DEV), need_colors=need_colors)
    torch.cuda.synchronize()
    K1 = pr.shape[3] + 1
    w = saved.winners_full().long()
    saved.hist = torch.zeros(w.shape[:-1] + (K1,), dtype=torch.int32, device=w.device).scatter_add_(
        -1, w, torch.ones_like(w, dtype=torch.int32))
    mask = pr.pix_to_face >= 0
    saved.counts.masked_fill_(~mask, 0)  # only valid entries are defined
    saved.rsum.masked_fill_(~mask, 0)
    return dict(image=image.cpu(), counts=counts_u16(saved).cpu(), winners=winners_long(saved).cpu(),
                hist=saved.hist.cpu(), rsum=saved.rsum.cpu(), grad_dists=gd.cpu(), grad_zbuf=gz.cpu(),
                grad_colors=None if gc is None else gc.cpu(), scalars=scal.cpu(), saved=saved)


def run_oracle(g, U, V):
    zn = g["znear"].reshape(-1, 1, 1, 1)
    zf = g["zfar"].reshape(-1, 1, 1, 1)
    return O.shade_fwd_bwd(g["pix_to_face"], g["zbuf"], g["dists"], g["colors"], g["background"], zn, zf,
                           g["sigma"], g["gamma"], g["alpha"], g["eps"], U, V, g["grad_image"])


def synthetic_case(N, H, W, K, S_r, S_a, kind="realistic", sigma=1e-3, gamma=1e-2, alpha=1.0, seed=0,
                   background=(1.0, 1.0, 1.0), znear=None, zfar=None, **frag_kw):
    """A dict with the golden keys, built on the CPU from the package's synthetic generator."""
    from pertrenderer_b200 import synthetic_fragments
    fr, col = synthetic_fragments(N, H, W, K, kind=kind, sigma=sigma, seed=seed, device="cpu", **frag_kw)
    gen = torch.Generator().manual_seed(seed + 99)
    return dict(pix_to_face=fr.pix_to_face, zbuf=fr.zbuf, dists=fr.dists, colors=col,
                znear=torch.tensor(znear if znear is not None else [1.0] * N),
                zfar=torch.tensor(zfar if zfar is not None else [100.0] * N),
                background=torch.tensor(background), sigma=sigma, gamma=gamma, alpha=alpha, eps=1e-10,
                S_r=S_r, S_a=S_a, grad_image=torch.randn((N, H, W, 4), generator=gen))
