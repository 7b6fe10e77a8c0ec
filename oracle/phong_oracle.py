"""CPU oracle for the Phong colours of RandomPhongShader (TEST INFRASTRUCTURE ONLY).

Restates, in plain torch-CPU tensor arithmetic, what the reference computes at
``randomras/random_rasterizer.py:101-113`` before the blend:

    texels = meshes.sample_textures(fragments)
    colors = pytorch3d.renderer.mesh.shading.phong_shading(meshes, fragments, texels, lights, cameras, materials)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module; the
product path (``pertrenderer_b200.shading``) never routes through it.

Parity status: UNPINNED at this boundary.  The arithmetic lives in pytorch3d 0.4.0
(``requirements.txt:7``), a third-party dependency that is neither vendored under ``/root/reference`` nor
installable here, and the reference holds no test or golden vector for it.  The functions below restate
pytorch3d's published algorithm (file / function names of the 0.4.0 release):

    renderer/mesh/shading.py   phong_shading, _apply_lighting
    renderer/lighting.py       PointLights.diffuse/specular, DirectionalLights.diffuse/specular, diffuse(), specular()
    ops/interp_face_attrs.py   interpolate_face_attributes  (masked entries -> 0)
    structures/meshes.py       Meshes._compute_vertex_normals (in pertrenderer_b200.structures.TriMeshes)

Anchors that ARE checked: closed-form known answers (``tests/test_phong_oracle.py``: head-on light, grazing
light, back-facing, mirror direction) and torch.autograd.gradcheck-style finite differences in float64.
Gradients here come from autograd over this restatement; the CUDA backward is hand-derived, so the two are
independent statements.
"""

from __future__ import annotations

import torch
import torch.nn.functional as F


def interpolate_face_attributes(pix_to_face, bary, face_attrs):
    """ops/interp_face_attrs.py: out[n,h,w,k] = sum_v bary[n,h,w,k,v] * face_attrs[pix_to_face[n,h,w,k], v],
    0 where pix_to_face < 0."""
    mask = pix_to_face >= 0
    idx = pix_to_face.clamp(min=0)
    vals = (bary[..., None] * face_attrs[idx]).sum(dim=-2)
    return torch.where(mask[..., None], vals, torch.zeros_like(vals))


def _bcast(t, like):
    """(rows,3) light / material attribute against (N,H,W,K,3) points: rows = 1 or N."""
    t = t.reshape(-1, t.shape[-1])
    return t.reshape(t.shape[0], 1, 1, 1, t.shape[-1]).to(like.dtype)


def diffuse(normals, color, direction):
    """renderer/lighting.py diffuse(): renormalise, relu(n.d), colour * angle."""
    normals = F.normalize(normals, p=2, dim=-1, eps=1e-6)
    direction = F.normalize(direction, p=2, dim=-1, eps=1e-6)
    angle = F.relu(torch.sum(normals * direction, dim=-1))
    return color * angle[..., None]


def specular(points, normals, direction, color, camera_position, shininess):
    """renderer/lighting.py specular(): reflect = -d + 2 (n.d) n, alpha = relu(v.r) * [n.d > 0], colour * alpha^s."""
    normals = F.normalize(normals, p=2, dim=-1, eps=1e-6)
    direction = F.normalize(direction, p=2, dim=-1, eps=1e-6)
    cos_angle = torch.sum(normals * direction, dim=-1)
    mask = (cos_angle > 0).to(points.dtype)
    view_direction = F.normalize(camera_position - points, p=2, dim=-1, eps=1e-6)
    reflect_direction = -direction + 2 * (cos_angle[..., None] * normals)
    alpha = F.relu(torch.sum(view_direction * reflect_direction, dim=-1)) * mask
    return color * torch.pow(alpha, shininess)[..., None]


def phong_colors(pix_to_face, bary, face_verts, face_normals, texels, *, light_location=None, light_direction=None,
                 light_ambient, light_diffuse, light_specular, mat_ambient, mat_diffuse, mat_specular, shininess,
                 camera_center):
    """renderer/mesh/shading.py phong_shading + _apply_lighting.

    pix_to_face (N,H,W,K) int64; bary (N,H,W,K,3); face_verts, face_normals (F,3,3); texels (N,H,W,K,3);
    light / material colours (rows,3), shininess (rows,), camera_center (rows,3), rows = 1 or N.
    Exactly one of light_location (PointLights) / light_direction (DirectionalLights).
    Returns colors (N,H,W,K,3) = (ambient + diffuse) * texels + specular."""
    points = interpolate_face_attributes(pix_to_face, bary, face_verts)
    normals = interpolate_face_attributes(pix_to_face, bary, face_normals)
    if light_location is not None:
        direction = _bcast(light_location, points) - points  # PointLights.diffuse / .specular
    else:
        direction = _bcast(light_direction, points).expand_as(points)
    light_d = diffuse(normals, _bcast(light_diffuse, points), direction)
    sh = shininess.reshape(-1).to(points.dtype).reshape(-1, 1, 1, 1)
    light_s = specular(points, normals, direction, _bcast(light_specular, points), _bcast(camera_center, points), sh)
    ambient = _bcast(mat_ambient, points) * _bcast(light_ambient, points)
    diff = _bcast(mat_diffuse, points) * light_d
    spec = _bcast(mat_specular, points) * light_s
    return (ambient + diff) * texels + spec


def phong_colors_from(meshes, fragments, lights, cameras, materials, texels):
    """The same call on the shim objects of pertrenderer_b200.structures (attribute access only)."""
    verts, faces = meshes.verts_packed(), meshes.faces_packed()
    kw = dict(light_ambient=lights.ambient_color, light_diffuse=lights.diffuse_color,
              light_specular=lights.specular_color, mat_ambient=materials.ambient_color,
              mat_diffuse=materials.diffuse_color, mat_specular=materials.specular_color,
              shininess=materials.shininess, camera_center=cameras.get_camera_center())
    if hasattr(lights, "location"):
        kw["light_location"] = lights.location
    else:
        kw["light_direction"] = lights.direction
    return phong_colors(fragments.pix_to_face, fragments.bary_coords, verts[faces], meshes.verts_normals_packed()[faces],
                        texels, **kw)
