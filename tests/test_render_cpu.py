"""CPU: camera algebra of sugar_b200.render.camera_matrices against a numpy restatement of
sugar_scene/sugar_model.py:2136-2163 + sugar_utils/graphics_utils.py:38-85 (the rasterizer call itself is GPU-only)."""
import math

import numpy as np
import torch


def reference_matrices(c2w34, fov_x, fov_y, znear, zfar, pp):
    c2w = np.concatenate([c2w34, np.array([[0, 0, 0, 1.0]])], 0)
    c2w[:3, 1:3] *= -1
    w2c = np.linalg.inv(c2w)
    R = np.transpose(w2c[:3, :3]); T = w2c[:3, 3]
    Rt = np.zeros((4, 4)); Rt[:3, :3] = R.transpose(); Rt[:3, 3] = T; Rt[3, 3] = 1.0      # getWorld2View
    world_view = np.float32(Rt).T
    tY, tX = math.tan(fov_y / 2), math.tan(fov_x / 2)
    top, right = tY * znear, tX * znear
    P = np.zeros((4, 4), np.float32)
    P[0, 0] = 2 * znear / (2 * right); P[1, 1] = 2 * znear / (2 * top)
    P[3, 2] = 1.0; P[2, 2] = zfar / (zfar - znear); P[2, 3] = -(zfar * znear) / (zfar - znear)
    proj = P.T.copy(); proj[2, 0] = -pp[0]; proj[2, 1] = -pp[1]
    return world_view, world_view @ proj, c2w[:3, 3]


def test_camera_matrices_match_reference_construction():
    from sugar_b200.render import camera_matrices
    rng = np.random.default_rng(0)
    for _ in range(5):
        A = rng.normal(size=(3, 3)); Q, _ = np.linalg.qr(A)
        c2w = np.concatenate([Q, rng.normal(size=(3, 1))], 1)
        fx, fy, pp = 1.1, 0.7, (0.03, -0.02)
        wv, fp, cc = camera_matrices(torch.from_numpy(c2w), fx, fy, 0.01, 100.0, pp)
        rwv, rfp, rcc = reference_matrices(c2w.copy(), fx, fy, 0.01, 100.0, pp)
        assert np.allclose(wv.numpy(), rwv, atol=1e-5)
        assert np.allclose(fp.numpy(), rfp, atol=1e-4)
        assert np.allclose(cc.numpy(), rcc, atol=1e-6)


def test_wrapper_mirror_matches_the_reference_wrapper_golden():
    """tests/golden/render_wrapper.npz holds what the reference's own render_image_gaussian_rasterizer handed to
    the rasterizer (make_render_golden.py); the mirror must build the same matrices, camera centre and colours."""
    import os
    from sugar_b200 import render
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "render_wrapper.npz"))
    wv, fp, cc = render.camera_matrices(torch.from_numpy(g["c2w"]), float(g["fov"][0]), float(g["fov"][1]),
                                        principal_point=(float(g["pp"][0]), float(g["pp"][1])))
    assert np.allclose(wv.numpy(), g["viewmatrix"], atol=2e-6)
    assert np.allclose(fp.numpy(), g["projmatrix"], atol=2e-5)
    assert np.allclose(cc.numpy(), g["campos"].reshape(3), atol=1e-6)
    assert abs(math.tan(g["fov"][0] / 2) - g["tanfov"][0]) < 1e-7
    cols = render.points_rgb(torch.from_numpy(g["points"]), torch.from_numpy(g["sh"]), cc, int(g["sh_degree"]) + 1)
    assert np.abs(cols.numpy() - g["colors_precomp"]).max() <= 2e-6
    assert np.array_equal(g["means3D"], g["points"]) and np.array_equal(g["opacities"], g["strengths"])
