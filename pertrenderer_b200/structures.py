"""Data contracts at the drop-in boundary, and synthetic inputs of the benchmark shapes.

pytorch3d (0.4.0 in the reference's requirements.txt:7) is not installable in this image; the hot
path only touches its objects by attribute, so duck-typed shims with the
same field names are enough.  When pytorch3d is installed its own ``Fragments`` / ``BlendParams`` /
cameras / ``Meshes`` objects work unchanged.
"""

from __future__ import annotations

import math
from typing import NamedTuple, Optional, Sequence, Union

import torch


class Fragments(NamedTuple):
    """pytorch3d.renderer.mesh.rasterizer.Fragments: (N,H,W,K) each, bary_coords (N,H,W,K,3).
    pix_to_face int64 with -1 = empty; zbuf / dists float32 with -1 = empty; K sorted near -> far,
    padding last."""
    pix_to_face: torch.Tensor
    zbuf: torch.Tensor
    bary_coords: Optional[torch.Tensor]
    dists: torch.Tensor


class BlendParams(NamedTuple):
    """pytorch3d.renderer.blending.BlendParams (only background_color is read on this path,
    random_rasterizer.py:39)."""
    sigma: float = 1e-4
    gamma: float = 1e-4
    background_color: Union[torch.Tensor, Sequence[float]] = (1.0, 1.0, 1.0)


class DepthCameras:
    """Minimal stand-in for a pytorch3d camera batch: the shader only reads ``znear`` / ``zfar`` as
    1-D tensors of length N (random_rasterizer.py:172-173)."""

    def __init__(self, znear=1.0, zfar=100.0, n=1, device="cpu"):
        self.znear = torch.as_tensor(znear, dtype=torch.float32, device=device).reshape(-1)
        self.zfar = torch.as_tensor(zfar, dtype=torch.float32, device=device).reshape(-1)
        if self.znear.numel() == 1 and n > 1:
            self.znear = self.znear.expand(n).clone()
        if self.zfar.numel() == 1 and n > 1:
            self.zfar = self.zfar.expand(n).clone()

    def to(self, device):
        self.znear, self.zfar = self.znear.to(device), self.zfar.to(device)
        return self


class TexelMeshes:
    """Stand-in for ``Meshes`` on this path: ``sample_textures(fragments)`` returns a preset
    (N,H,W,K,3) texel tensor (random_rasterizer.py:170)."""

    def __init__(self, texels: torch.Tensor):
        self.texels = texels

    def sample_textures(self, fragments):
        return self.texels


class FaceTexels:
    """Lazy texels: the colour of fragment (n,h,w,k) is ``face_colors[pix_to_face[n,h,w,k]]``.
    ``smooth_rgb_blend`` gathers it inside the fused kernels (and scatters the gradient back to
    ``face_colors``), so the (N,H,W,K,3) tensor of ``Meshes.sample_textures``
    (random_rasterizer.py:170) is never materialised."""

    def __init__(self, face_colors: torch.Tensor):
        if face_colors.dim() != 2 or face_colors.shape[1] != 3:
            raise ValueError("face_colors must be (F,3)")
        self.face_colors = face_colors

    def materialize(self, pix_to_face: torch.Tensor) -> torch.Tensor:
        mask = pix_to_face >= 0
        return self.face_colors[pix_to_face.clamp(min=0)] * mask[..., None]


class FaceColorMeshes:
    """Stand-in for ``Meshes`` with one colour per (packed) face: ``sample_textures`` returns lazy
    :class:`FaceTexels` instead of a texel tensor."""

    def __init__(self, face_colors: torch.Tensor):
        self.face_colors = face_colors

    def sample_textures(self, fragments):
        return FaceTexels(self.face_colors)


def blur_radius(sigma: float) -> float:
    """experiments/eval.py:137: log(1/1e-4 - 1) * sigma."""
    return math.log(1.0 / 1e-4 - 1.0) * sigma


def synthetic_fragments(N, H, W, K, kind="dense", sigma=1e-3, n_faces=1280, seed=0, device="cuda",
                        coverage=0.6, mean_valid=4.0, frac_edge=0.15):
    """Synthetic Fragments + texels of the benchmark shapes.

    dense:     every pixel has all K faces valid, dists ~ U(-b, b), b = blur radius.
    realistic: 60 % of pixels covered, valid count per covered pixel geometric with mean 4 (capped
               at K, padding last), dists 85 % interior -U(0, 0.05) and 15 % edge band U(-b, b).
    Valid entries: pix_to_face ~ U{0..F-1}, zbuf ascending in K ~ U(5.5, 8.0), colours ~ U(0,1).
    Padded entries: pix_to_face = -1, zbuf = dists = -1, colours = 0.
    Returns (Fragments, colors (N,H,W,K,3)).
    """
    g = torch.Generator(device=device).manual_seed(seed)
    b = blur_radius(sigma)
    shape = (N, H, W, K)
    u = lambda *s: torch.rand(*s, generator=g, device=device)  # noqa: E731
    if kind == "dense":
        valid = torch.ones(shape, dtype=torch.bool, device=device)
        dists = (u(shape) * 2 - 1) * b
    elif kind == "realistic":
        covered = u((N, H, W, 1)) < coverage
        # geometric with mean m on {1,2,...}: 1 + floor(log(U)/log(1-1/m))
        q = 1.0 - 1.0 / mean_valid
        nvalid = (1 + torch.floor(torch.log(u((N, H, W, 1)).clamp_min(1e-12)) / math.log(q))).clamp(max=K)
        nvalid = torch.where(covered, nvalid, torch.zeros_like(nvalid))
        valid = torch.arange(K, device=device).view(1, 1, 1, K) < nvalid
        edge = u(shape) < frac_edge
        dists = torch.where(edge, (u(shape) * 2 - 1) * b, -0.05 * u(shape))
    else:
        raise ValueError(f"unknown fragment kind {kind!r}")
    z = 5.5 + 2.5 * u(shape)
    z = torch.where(valid, z, torch.full_like(z, float("inf"))).sort(dim=-1).values  # ascending, padding last
    zbuf = torch.where(valid, z, torch.full_like(z, -1.0))
    dists = torch.where(valid, dists, torch.full_like(dists, -1.0))
    faces = torch.randint(0, n_faces, shape, generator=g, device=device, dtype=torch.int64)
    pix_to_face = torch.where(valid, faces, torch.full_like(faces, -1))
    colors = u((N, H, W, K, 3)) * valid[..., None]
    return Fragments(pix_to_face, zbuf.float(), None, dists.float()), colors.float()
