"""Multi-GPU partitioning of the perturbed shader: one process per GPU.

Batch sharding (BASELINE config 3)
    Pixels are independent, so a batch of views is split over ranks with NO collective on the data
    path; ``pixel_offset`` keeps the Philox counters global, so the union of the shards is bit-
    identical to the single-GPU job.  Only d/d(sigma, gamma, alpha) are summed (3 floats).

Noise-sample sharding (BASELINE config 4: few pixels, thousands of samples)
    Inputs are replicated; rank r draws samples [s0_r, s1_r).  Forward is two-stage and the second
    stage consumes log(mean over ALL samples of stage one) (smoothagg.py:200-201), so there are
    three exchanges per forward+backward, each ONE all-reduce:
        AR1  hit counts + score sums      (2,P,K)    float   after the coverage phase
        AR2  winner histogram             (P,K1)     int32   after the argmax phase
        AR3  score sums of the argmax     (P,K1+2)   float   in backward
    after which every rank finishes identically.  The collective plumbing is torch.distributed
    (NCCL over NVLink on the GPU box, gloo in the CPU tests); the phases are the PERT_PH_* phases of
    the fused kernels.  The orchestration is written against a small ``stages`` interface so that
    the CPU tests can drive it with the oracle while production drives it with the CUDA kernels.
"""

from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist
from torch.autograd import Function

# ------------------------------------------------------------------------------------------------
# partitioning
# ------------------------------------------------------------------------------------------------


def batch_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split of ``n_items`` batch elements: the first ``n % world`` ranks get
    one more."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def sample_range(n_samples: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of the noise samples in whole quads (the kernels draw four samples per
    Philox call, so shard boundaries are multiples of 4).  Ranks past the last quad get an empty
    range."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    quads = (n_samples + 3) // 4
    q0, q1 = batch_range(quads, world, rank)
    return min(4 * q0, n_samples), min(4 * q1, n_samples)


def check_sample_sharding(n_rast: int, n_agg: int, world: int) -> None:
    """Every rank needs at least one quad of each stage's samples.  Raises the same ValueError on every rank."""
    for name, n in (("GaussianRast", n_rast), ("GaussianAgg", n_agg)):
        if (n + 3) // 4 < world:
            raise ValueError(f"{name}.nb_samples = {n} is too small to shard over {world} ranks "
                             "(the kernels draw four samples per Philox call: one quad per rank at least)")


def all_reduce_scalar_grads(tensors, group=None, device=None):
    """Batch sharding: sum the gradients of the 0-dim CPU leaves (sigma, gamma, alpha) over ranks
    with one 3-float all-reduce."""
    grads = [t.grad if t.grad is not None else torch.zeros_like(t) for t in tensors]
    buf = torch.stack([g.detach().float().reshape(()) for g in grads])
    if device is not None:
        buf = buf.to(device)
    dist.all_reduce(buf, group=group)
    buf = buf.cpu()
    for t, v in zip(tensors, buf):
        t.grad = v.to(t.dtype).reshape(t.shape)


# ------------------------------------------------------------------------------------------------
# sample-sharded orchestration
# ------------------------------------------------------------------------------------------------
class CudaStages:
    """The phases of the fused kernels (include/pertshade.h PERT_PH_*) on one rank's sample shard."""

    def __init__(self, pr):
        from . import _cabi, ops
        self.ops, self.cabi, self.pr = ops, _cabi, pr
        self.saved = None

    def rast(self):
        _, self.saved = self.ops.shade_forward(self.pr, want_hist=True, phases=self.cabi.PH_RAST)
        counts = (self.saved.counts.to(torch.int32) & 0xFFFF).float()
        return counts, self.saved.rsum

    def agg(self, counts, rsum):
        self.saved.counts.copy_(counts.to(torch.int32).to(torch.int16))  # uint16 bit pattern
        self.saved.rsum.copy_(rsum)
        self.ops.shade_forward(self.pr, phases=self.cabi.PH_AGG, saved=self.saved)
        return self.saved.hist

    def blend(self, hist):
        if hist is not self.saved.hist:
            self.saved.hist.copy_(hist)
        image, _ = self.ops.shade_forward(self.pr, phases=self.cabi.PH_BLEND, saved=self.saved)
        return image

    def bwd_sample(self, grad_image):
        N, H, W, K = self.pr.shape
        dev = self.pr.device
        acc = torch.empty((N, H, W, K + 3), dtype=torch.float32, device=dev)
        # the kernel writes acc (P,K1) and pixstat (P,2) separately; pack them for ONE all-reduce
        a = torch.empty((N, H, W, K + 1), dtype=torch.float32, device=dev)
        t = torch.empty((N, H, W, 2), dtype=torch.float32, device=dev)
        self.ops.shade_backward(self.pr, self.saved, grad_image, phases=self.cabi.PH_BWD_SAMPLE, acc=a, pixstat=t,
                                use_hist=True)
        acc[..., :K + 1] = a
        acc[..., K + 1:] = t
        return acc

    def bwd_finish(self, grad_image, packed, need_colors=True):
        K = self.pr.shape[3]
        a = packed[..., :K + 1].contiguous()
        t = packed[..., K + 1:].contiguous()
        return self.ops.shade_backward(self.pr, self.saved, grad_image, need_colors=need_colors,
                                       phases=self.cabi.PH_BWD_FINISH, acc=a, pixstat=t, use_hist=True)


def sharded_forward(stages, group=None):
    """AR1 and AR2 around the three forward phases.  Returns the image (identical on every rank)."""
    counts, rsum = stages.rast()
    packed = torch.stack((counts, rsum))  # counts <= 65535 are exact in fp32
    dist.all_reduce(packed, group=group)
    hist = stages.agg(packed[0], packed[1])
    dist.all_reduce(hist, group=group)
    return stages.blend(hist)


def sharded_backward(stages, grad_image, group=None, need_colors=True):
    """AR3 between the two backward phases.  Every rank returns the full gradients."""
    packed = stages.bwd_sample(grad_image)
    dist.all_reduce(packed, group=group)
    return stages.bwd_finish(grad_image, packed, need_colors=need_colors)


class _SampleShardedShade(Function):
    """Autograd wrapper of the sample-sharded shader; same differentiable inputs as the fused
    single-GPU Function (random_rasterizer._PerturbedShade)."""

    @staticmethod
    def forward(ctx, colors, dists, zbuf, sigma, gamma, alpha, pix_to_face, znear, zfar, cfg):
        from . import ops
        group = cfg.get("group")
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        dev = pix_to_face.device
        if cfg.get("sync_seeds", False):
            # reproduce the single-GPU sample path: rank 0 draws the two seeds, everyone receives them
            # (one broadcast and one host read per forward)
            seeds = torch.zeros(2, dtype=torch.int64)
            if rank == 0:
                seeds[0] = ops.draw_seed()
                if cfg["fixed_noise"]:
                    torch.manual_seed(1)
                seeds[1] = ops.draw_seed()
            seeds = seeds.to(dev)
            dist.broadcast(seeds, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            seed_r, seed_a = (int(v) for v in seeds.cpu())
        else:
            # every rank draws its own seeds from its own torch generator: the shards own disjoint sample
            # indices, so whatever the seeds are the union is S independent samples (no communication)
            seed_r = ops.draw_seed()
            if cfg["fixed_noise"]:
                torch.manual_seed(1)
            seed_a = ops.draw_seed()
        S_r, S_a = int(cfg["S_rast"]), int(cfg["S_agg"])
        # decided on data every rank has, so that either all ranks raise or none does (a rank that raised alone would
        # leave the others blocked in the first all-reduce)
        check_sample_sharding(S_r, S_a, world)
        sr, sa = sample_range(S_r, world, rank), sample_range(S_a, world, rank)
        pr = ops.ShadeProblem(
            pix_to_face=pix_to_face, zbuf=zbuf, dists=dists, colors=colors, znear=znear, zfar=zfar,
            background=cfg["background"], sigma=float(sigma), gamma=float(gamma), alpha=float(alpha),
            eps=float(cfg["eps"]), S_rast=S_r, S_agg=S_a, seed_rast=seed_r, seed_agg=seed_a,
            flags=ops.current_flags() | int(cfg.get("flags", 0)), s_rast=sr, s_agg=sa)
        stages = CudaStages(pr)
        image = sharded_forward(stages, group)
        ctx.stages, ctx.group, ctx.scalars = stages, group, (sigma, gamma, alpha)
        return image

    @staticmethod
    def backward(ctx, grad_image):
        need = ctx.needs_input_grad
        gd, gz, gc, scal = sharded_backward(ctx.stages, grad_image.contiguous(), ctx.group, need_colors=need[0])
        out = [None, None, None]
        if any(need[3:6]):
            host = scal.cpu()
            for i, t in enumerate(ctx.scalars):
                if need[3 + i] and torch.is_tensor(t):
                    out[i] = host[i].to(dtype=t.dtype).reshape(t.shape).to(t.device)
        return (gc if need[0] else None, gd if need[1] else None, gz if need[2] else None,
                out[0], out[1], out[2], None, None, None, None)


def smooth_rgb_blend_sample_sharded(colors, fragments, smoothrast, smoothagg, blend_params, znear=1.0, zfar=100,
                                    group=None, sync_seeds=False) -> torch.Tensor:
    """``smooth_rgb_blend`` with the noise samples split over the ranks of ``group`` (inputs
    replicated on every rank; every rank returns the same image and, after backward, the same
    gradients).  ``sync_seeds=True`` makes rank 0's torch generator drive the noise of every rank, which
    reproduces the single-GPU sample path exactly (one broadcast + host read per call); by default each
    rank seeds its own disjoint sample shard from its own generator."""
    from .random_rasterizer import _background_tuple
    from .smoothagg import GaussianAgg
    from .smoothrast import GaussianRast
    if not (isinstance(smoothrast, GaussianRast) and isinstance(smoothagg, GaussianAgg)):
        raise ValueError("sample sharding is implemented for the (GaussianRast, GaussianAgg) pair")
    cfg = dict(background=_background_tuple(blend_params), eps=smoothagg.eps, S_rast=smoothrast.nb_samples,
               S_agg=smoothagg.nb_samples, fixed_noise=bool(smoothagg.fixed_noise), group=group, sync_seeds=sync_seeds)
    return _SampleShardedShade.apply(colors, fragments.dists, fragments.zbuf, smoothrast.sigma, smoothagg.gamma,
                                     smoothagg.alpha, fragments.pix_to_face, znear, zfar, cfg)
