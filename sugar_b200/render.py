"""Host-side mirror of SuGaR.render_image_gaussian_rasterizer (sugar_scene/sugar_model.py:2085-2294).

The reference builds the view / projection matrices on the CPU with numpy (`.cpu().numpy()` +
`np.linalg.inv`, a host sync every call, sugar_model.py:2136-2161), evaluates SH colours with ~30
PyTorch temporaries when `compute_color_in_rasterizer=False` (the trainers' setting, coarse_sdf.py:51), and
then calls GaussianRasterizer.  Here the camera algebra stays on the device (no sync); colours come from the
in-kernel SH path by default (arithmetically the reference rasterizer's own, forward.cu:20-71), or -- with
`compute_color_in_rasterizer=False` -- from `points_rgb`, the reference's python evaluation, op for op.
"""
import math

import torch

from .diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer


def camera_matrices(c2w: torch.Tensor, fov_x: float, fov_y: float, znear: float = 0.01, zfar: float = 100.0,
                    principal_point=(0.0, 0.0)):
    """c2w: [3,4] or [4,4] camera-to-world in the OpenGL/Blender convention nerfstudio stores (Y up, Z back).
    Returns (world_view_transform, full_proj_transform, camera_center) exactly as sugar_model.py:2136-2163
    builds them: both matrices transposed (row-vector convention), projection patched with the principal point."""
    dev, dt = c2w.device, torch.float32
    m = torch.eye(4, device=dev, dtype=dt)
    m[:3, :4] = c2w[:3, :4].to(dt)
    m[:3, 1:3] *= -1                       # OpenGL -> COLMAP axes (sugar_model.py:2142)
    w2c = torch.linalg.inv(m)
    world_view = w2c.transpose(0, 1).contiguous()   # getWorld2View(R=w2c[:3,:3].T, t=w2c[:3,3]).T == w2c.T
    tx, ty = math.tan(fov_x / 2), math.tan(fov_y / 2)
    P = torch.zeros(4, 4, device=dev, dtype=dt)     # getProjectionMatrix (graphics_utils.py:65-85)
    P[0, 0] = 1.0 / tx
    P[1, 1] = 1.0 / ty
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    proj = P.transpose(0, 1).contiguous()
    proj[2, 0] = -principal_point[0]                # sugar_model.py:2158-2159
    proj[2, 1] = -principal_point[1]
    full_proj = world_view @ proj
    return world_view, full_proj, m[:3, 3].contiguous()


_SH_C0 = 0.28209479177387814
_SH_C1 = 0.4886025119029199
_SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
          1.445305721320277, -0.5900435899266435)


def points_rgb(positions, sh_coordinates, camera_center, sh_levels):
    """SuGaR.get_points_rgb (sugar_model.py:839-883) with eval_sh (sugar_utils/spherical_harmonics.py:117-172):
    colours = clamp_min(eval_sh(normalize(positions - camera_center)) + 0.5, 0), same operation order."""
    d = torch.nn.functional.normalize(positions - camera_center.reshape(1, 3), dim=-1)
    sh = sh_coordinates[:, :sh_levels ** 2].transpose(-1, -2).reshape(-1, 3, sh_levels ** 2)
    deg = sh_levels - 1
    res = _SH_C0 * sh[..., 0]
    if deg > 0:
        x, y, z = d[..., 0:1], d[..., 1:2], d[..., 2:3]
        res = (res - _SH_C1 * y * sh[..., 1] + _SH_C1 * z * sh[..., 2] - _SH_C1 * x * sh[..., 3])
        if deg > 1:
            xx, yy, zz = x * x, y * y, z * z
            xy, yz, xz = x * y, y * z, x * z
            res = (res + _SH_C2[0] * xy * sh[..., 4] + _SH_C2[1] * yz * sh[..., 5] +
                   _SH_C2[2] * (2.0 * zz - xx - yy) * sh[..., 6] + _SH_C2[3] * xz * sh[..., 7] +
                   _SH_C2[4] * (xx - yy) * sh[..., 8])
            if deg > 2:
                res = (res + _SH_C3[0] * y * (3 * xx - yy) * sh[..., 9] + _SH_C3[1] * xy * z * sh[..., 10] +
                       _SH_C3[2] * y * (4 * zz - xx - yy) * sh[..., 11] +
                       _SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12] +
                       _SH_C3[4] * x * (4 * zz - xx - yy) * sh[..., 13] + _SH_C3[5] * z * (xx - yy) * sh[..., 14] +
                       _SH_C3[6] * x * (xx - 3 * yy) * sh[..., 15])
    return torch.clamp_min(res + 0.5, 0.0).view(-1, 3)


def render_image_gaussian_rasterizer(points, scaling, quaternions, opacities, sh_coordinates, c2w, fov_x, fov_y,
                                     image_height, image_width, bg_color=None, sh_deg=3, point_colors=None,
                                     principal_point=(0.0, 0.0), return_2d_radii=False,
                                     compute_color_in_rasterizer=True):
    """Render one view.  `point_colors` (P,3) selects the reference's colors_precomp path (used for its depth
    renders, coarse_sdf.py:578-590); otherwise colours come from `sh_coordinates` [P,M,3]: in-kernel, or with
    `compute_color_in_rasterizer=False` by the reference's python evaluation (`points_rgb`) as colors_precomp,
    exactly the trainers' call.  Returns image (H,W,3) like the reference, or a dict with radii / viewspace_points."""
    dev = points.device
    if bg_color is None:
        bg_color = torch.zeros(3, device=dev)
    world_view, full_proj, cam_center = camera_matrices(c2w, fov_x, fov_y, principal_point=principal_point)
    settings = GaussianRasterizationSettings(
        image_height=int(image_height), image_width=int(image_width), tanfovx=math.tan(fov_x * 0.5),
        tanfovy=math.tan(fov_y * 0.5), bg=bg_color, scale_modifier=1., viewmatrix=world_view, projmatrix=full_proj,
        sh_degree=sh_deg, campos=cam_center, prefiltered=False, debug=False)
    means2D = torch.zeros_like(points, requires_grad=True)      # gradient holder (sugar_model.py:2248)
    if point_colors is None and not compute_color_in_rasterizer:
        point_colors = points_rgb(points, sh_coordinates, cam_center, sh_deg + 1)   # sugar_model.py:2187-2193
    kw = dict(colors_precomp=point_colors) if point_colors is not None else dict(shs=sh_coordinates)
    image, radii = GaussianRasterizer(settings)(means3D=points, means2D=means2D, opacities=opacities,
                                                scales=scaling, rotations=quaternions, **kw)
    image = image.permute(1, 2, 0)
    if return_2d_radii:
        return {"image": image, "radii": radii, "viewspace_points": means2D}
    return image
