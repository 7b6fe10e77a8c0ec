"""pertrenderer_b200 — B200-native perturbed shading (GaussianRast -> GaussianAgg -> blend).

Drop-in for the hot path of quentinll/pertrenderer (``randomras``): the same public names
(``randomras/__init__.py:1-3``), backed by hand-written sm_100a kernels behind a C ABI
(``include/pertshade.h``).  CUDA only: there is no CPU or PyTorch fallback on this path.
"""

from .random_rasterizer import RandomPhongShader, RandomSimpleShader, SimpleShader, SoftSimpleShader, smooth_rgb_blend
from .rasterizer import (FoVPerspectiveCameras, MeshRasterizer, MeshRenderer, OpenGLPerspectiveCameras,
                         RasterizationSettings, look_at_view_transform, rasterize_meshes)
from .shading import phong_shading, sample_lazy_textures
from .smoothagg import CauchyAgg, GaussianAgg, GaussianAgg_wovr, HardAgg, SoftAgg, UniformAgg, randomArgmax, randomArgmax_wovr
from .smoothrast import (AffineRast, ArctanRast, GaussianRast, GaussianRast_wovr, HardRast, SoftRast, randomHeaviside,
                         randomHeaviside_wovr)
from .structures import (AtlasTexels, BlendParams, DepthCameras, DirectionalLights, FaceColorMeshes, FaceTexels, Fragments, Materials,
                         PointLights, TexelMeshes, TriMeshes, UVTexels, VertexTexels, ViewCameras, synthetic_bary, synthetic_fragments,
                         synthetic_mesh)
from .ops import explicit_noise, kernel_flags

__all__ = [
    "RandomPhongShader", "RandomSimpleShader", "SimpleShader", "SoftSimpleShader", "UniformAgg", "phong_shading", "PointLights", "DirectionalLights",
    "MeshRasterizer", "MeshRenderer", "RasterizationSettings", "FoVPerspectiveCameras", "OpenGLPerspectiveCameras",
    "look_at_view_transform", "rasterize_meshes", "Materials", "ViewCameras", "TriMeshes", "VertexTexels", "UVTexels", "sample_lazy_textures", "synthetic_mesh", "synthetic_bary", "smooth_rgb_blend", "GaussianAgg", "SoftAgg", "CauchyAgg", "HardAgg",
    "randomArgmax", "GaussianRast", "GaussianRast_wovr", "GaussianAgg_wovr", "randomHeaviside_wovr", "randomArgmax_wovr", "ArctanRast", "AffineRast", "HardRast",
    "SoftRast", "randomHeaviside", "BlendParams", "DepthCameras", "Fragments", "TexelMeshes",
    "synthetic_fragments", "explicit_noise", "kernel_flags", "FaceColorMeshes", "FaceTexels", "AtlasTexels",
]
__version__ = "0.1.0"
