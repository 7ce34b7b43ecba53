"""Golden for SuGaR.render_image_gaussian_rasterizer (sugar_scene/sugar_model.py:2085-2294) as the trainers call
it (compute_color_in_rasterizer=False, coarse_sdf.py:51): the reference's OWN wrapper is run on the CPU with a
recording stand-in for GaussianRasterizer, so everything the wrapper computes -- view / projection matrices with
the principal-point patch, camera centre, python SH colours, activations -- is captured exactly as the rasterizer
would receive it.  Needs /root/reference:  python tests/golden/make_render_golden.py -> tests/golden/render_wrapper.npz
"""
import math
import os
import sys
from typing import NamedTuple

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_field_golden import import_reference_sugar  # noqa: E402


class Settings(NamedTuple):  # the 12 fields of diff_gaussian_rasterization.GaussianRasterizationSettings (:157-169)
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


RECORD = {}


class Recorder:
    def __init__(self, raster_settings):
        RECORD["settings"] = raster_settings

    def __call__(self, **kw):
        RECORD["call"] = kw
        rs = RECORD["settings"]
        return torch.zeros(3, rs.image_height, rs.image_width), torch.zeros(kw["means3D"].shape[0], dtype=torch.int32)


def main():
    sm = import_reference_sugar()
    sm.GaussianRasterizer = Recorder
    sm.GaussianRasterizationSettings = Settings
    torch.Tensor.cuda = lambda self, *a, **k: self     # the wrapper moves its matrices with .cuda()
    SuGaR = sm.SuGaR
    g = torch.Generator().manual_seed(0)
    P, H, W = 400, 90, 150
    fov_x, fov_y = 1.05, 0.68
    A = torch.randn(3, 3, generator=g).double()
    Q, _ = torch.linalg.qr(A)
    if torch.det(Q) < 0:
        Q[:, 0] = -Q[:, 0]
    c2w = torch.cat([Q, torch.tensor([[0.3], [-0.2], [0.5]]).double()], 1).float()            # nerfstudio c2w [3,4]
    # points in front of this camera: OpenGL axes (camera looks down -z)
    cam_pts = torch.stack([(torch.rand(P, generator=g) - 0.5) * 4, (torch.rand(P, generator=g) - 0.5) * 2.4,
                           -(2 + 6 * torch.rand(P, generator=g))], 1)
    points = cam_pts @ c2w[:, :3].T + c2w[:, 3]
    pp = (0.02, -0.015)

    class P3D:
        znear = torch.tensor([0.01]); zfar = torch.tensor([100.0])
        K = torch.zeros(1, 4, 4)

        def get_camera_center(self):
            return c2w[:, 3].view(1, 3).clone()
    P3D.K[0, 0, 2], P3D.K[0, 1, 2] = pp

    class Cams:
        p3d_cameras = [P3D()]
        camera_to_worlds = c2w[None]

    class Fake:
        device = "cpu"
        image_height, image_width = H, W
        tanfovx, tanfovy = math.tan(fov_x / 2), math.tan(fov_y / 2)
        n_points = P
        _points = points
        sh_coordinates = torch.cat([0.5 * torch.randn(P, 1, 3, generator=g), 0.1 * torch.randn(P, 15, 3, generator=g)], 1)
        strengths = torch.sigmoid(torch.randn(P, 1, generator=g))
        scaling = torch.exp(torch.randn(P, 3, generator=g) * 0.4 - 3.0)
        quaternions = torch.nn.functional.normalize(torch.randn(P, 4, generator=g), dim=-1)

        def get_points_rgb(self, **kw):
            return SuGaR.get_points_rgb(self, **kw)
    Fake.fov_x, Fake.fov_y, Fake.points = fov_x, fov_y, points
    fake = Fake()
    out = SuGaR.render_image_gaussian_rasterizer(fake, nerf_cameras=Cams(), camera_indices=0, sh_deg=3,
                                                 compute_color_in_rasterizer=False, bg_color=torch.tensor([0.1, 0.2, 0.3]))
    assert out.shape == (H, W, 3)
    rs, call = RECORD["settings"], RECORD["call"]
    assert call["shs"] is None and call["cov3D_precomp"] is None
    np.savez_compressed(
        os.path.join(HERE, "render_wrapper.npz"), c2w=c2w.numpy(), fov=np.array([fov_x, fov_y]), pp=np.array(pp),
        hw=np.array([H, W]), points=points.numpy(), sh=fake.sh_coordinates.numpy(), strengths=fake.strengths.numpy(),
        scaling=fake.scaling.numpy(), quaternions=fake.quaternions.numpy(), bg=rs.bg.numpy(),
        viewmatrix=rs.viewmatrix.numpy(), projmatrix=rs.projmatrix.numpy(), campos=rs.campos.numpy(),
        tanfov=np.array([rs.tanfovx, rs.tanfovy]), sh_degree=np.array(rs.sh_degree),
        colors_precomp=call["colors_precomp"].numpy(), opacities=call["opacities"].numpy(),
        scales=call["scales"].numpy(), rotations=call["rotations"].numpy(), means3D=call["means3D"].numpy())
    print("ok", rs.viewmatrix.shape, call["colors_precomp"].shape, float(call["colors_precomp"].mean()))


if __name__ == "__main__":
    main()
