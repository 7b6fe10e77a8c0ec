"""Import the unmodified reference operators from ``oracle/_ref`` (see oracle/build_ref.py).  TEST INFRASTRUCTURE ONLY.

``load()`` returns ``(random_rasterizer, smoothrast, smoothagg)`` modules of the reference's ``randomras`` package, or
``None`` when ``oracle/_ref`` has not been installed.  The pytorch3d names the package imports at module load time
(randomras/random_rasterizer.py:8-26) are stubbed: the shading path only reads attributes of the objects it is handed
(``fragments.pix_to_face / .zbuf / .dists``, ``blend_params.background_color``, ``cameras.znear / .zfar``,
``meshes.sample_textures``).
"""

from __future__ import annotations

import importlib
import os
import sys
import types
from collections import namedtuple

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

Fragments = namedtuple("Fragments", ["pix_to_face", "zbuf", "bary_coords", "dists"])
Blend = namedtuple("Blend", ["sigma", "gamma", "background_color"])


class Cameras:
    """What RandomSimpleShader.forward reads of a camera batch (random_rasterizer.py:172-173)."""

    def __init__(self, znear, zfar):
        self.znear, self.zfar = znear, zfar


class Texels:
    """Meshes stand-in: sample_textures returns a preset (N,H,W,K,3) tensor (random_rasterizer.py:170)."""

    def __init__(self, texels):
        self.texels = texels

    def sample_textures(self, fragments):
        return self.texels


def _stub_pytorch3d():
    if "pytorch3d" in sys.modules and not getattr(sys.modules["pytorch3d"], "_pert_stub", False):
        return  # a real pytorch3d is installed: use it
    names = ["look_at_view_transform", "OpenGLPerspectiveCameras", "PointLights", "DirectionalLights",
             "Materials", "RasterizationSettings", "MeshRenderer", "MeshRasterizer", "SoftPhongShader",
             "HardPhongShader", "SoftSilhouetteShader", "hard_rgb_blend", "softmax_rgb_blend",
             "TexturesVertex", "BlendParams"]
    p3d = types.ModuleType("pytorch3d")
    p3d._pert_stub = True
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


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "randomras", "random_rasterizer.py"))


def load():
    if not available():
        return None
    try:
        import pytorch3d  # noqa: F401
    except Exception:
        _stub_pytorch3d()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    rr = importlib.import_module("randomras.random_rasterizer")
    sr = importlib.import_module("randomras.smoothrast")
    sa = importlib.import_module("randomras.smoothagg")
    return rr, sr, sa
