"""GPU: level-set ray sampling (sugar_b200/levelset.py, fused field kernel with shared neighbour rows) vs the
reference's own lines (golden) and vs the CPU oracle at a larger size.  A density within fp32 rounding of a
level can flip a ray between valid/empty or move its first crossing by one sample, so rays are matched through
the validity masks and at least 99.5 % must agree; matched points within 1e-4 of the scene scale, normals 1e-3."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


def _ours(case, cam, levels, density_factor):
    from sugar_b200 import levelset
    t = lambda k: torch.from_numpy(case[k]).cuda()
    out = levelset.level_surface_points(t("x"), torch.from_numpy(cam).cuda(), t("nbr_idx"), t("points"), t("scaling"),
                                        t("quaternions"), t("strengths"), surface_levels=levels,
                                        density_factor=density_factor, return_normals=True)
    return {lv: {k: v.detach().cpu().numpy() for k, v in o.items()} for lv, o in out.items()}


def _compare(ours, ref_valid, ref_points, ref_normals, n):
    both = ours["valid"] & ref_valid
    assert (ours["valid"] == ref_valid).mean() >= 0.995
    po = np.zeros((n, 3), np.float32); po[ours["valid"]] = ours["intersection_points"]
    pr = np.zeros((n, 3), np.float32); pr[ref_valid] = ref_points
    no = np.zeros((n, 3), np.float32); no[ours["valid"]] = ours["normals"]
    nr = np.zeros((n, 3), np.float32); nr[ref_valid] = ref_normals
    dp = np.abs(po[both] - pr[both]).max(axis=1)
    dn = np.abs(no[both] - nr[both]).max(axis=1)
    scale = np.abs(pr[both]).max()
    assert (dp <= 1e-4 * scale).mean() >= 0.995 and (dn <= 1e-3).mean() >= 0.995
    assert both.sum() > 0


@pytest.mark.parametrize("name", ["k16", "k8"])
def test_levelset_matches_reference_golden(name):
    from make_levelset_golden import CASES, LEVELS, make_inputs
    gold = np.load(os.path.join(HERE, "golden", f"levelset_{name}.npz"))
    case, cam = make_inputs(CASES[name])
    ours = _ours(case, cam, LEVELS, CASES[name]["density_factor"])
    n = case["x"].shape[0]
    for lv in LEVELS:
        gi = gold[f"gaussian_idx_{lv}"]
        # the golden keeps the rays' Gaussian ids: rebuild its validity mask by matching them in order
        ref_valid = np.zeros(n, bool)
        j = 0
        for i in range(n):
            if j < len(gi) and case["gaussian_idx"][i] == gi[j]:
                ref_valid[i] = True
                j += 1
        assert j == len(gi)
        _compare(ours[lv], ref_valid, gold[f"points_{lv}"], gold[f"normals_{lv}"], n)


def test_levelset_matches_oracle_large():
    from oracle import field_oracle as fo
    case = fo.make_case(P=20000, N=30000, K=16, seed=21, density_factor=1.0)
    cam = np.array([0.5, 0.1, -7.0], np.float32)
    levels = [0.1, 0.3, 0.5]
    ours = _ours(case, cam, levels, 1.0)
    t = lambda k: torch.from_numpy(case[k])
    ref = fo.level_surface_points_torch(t("x"), torch.from_numpy(cam), t("nbr_idx"), t("points"), t("scaling"),
                                        t("quaternions"), t("strengths"), surface_levels=levels, density_factor=1.0)
    for lv in levels:
        r = ref[lv]
        _compare(ours[lv], r["valid"].numpy(), r["intersection_points"].numpy(), r["normals"].numpy(), 30000)


def test_level_surface_points_from_camera_finds_a_rendered_sheet():
    """The per-view call with the Gaussian-depth front half (sugar_model.py:1898-1962): a dense sheet of flat,
    opaque Gaussians on the plane z = 5 in front of an identity camera must give level points on that plane,
    with normals along the view axis, for (nearly) every covered pixel."""
    from sugar_b200 import levelset, scenes, steps
    W, H, P = 160, 96, 60_000
    sc = scenes.make_scene(1000, W, H, seed=0)           # camera only (identity pose)
    g = torch.Generator().manual_seed(0)
    xy = (torch.rand(P, 2, generator=g) - 0.5) * torch.tensor([2 * 5 * sc.tanfovx, 2 * 5 * sc.tanfovy]) * 1.1
    pts = torch.cat([xy, torch.full((P, 1), 5.0)], 1).cuda()
    scaling = torch.tensor([0.06, 0.06, 0.004]).repeat(P, 1).cuda()     # flat along z
    quats = torch.tensor([1.0, 0.0, 0.0, 0.0]).repeat(P, 1).cuda()
    strengths = torch.full((P,), 0.95).cuda()
    cam = steps.camera_from_scene(sc, "cuda")
    out = levelset.level_surface_points_from_camera(pts, scaling, quats, strengths, cam, surface_levels=(0.3,),
                                                    return_normals=True, return_pixel_idx=True)[0.3]
    n_valid = int(out["valid"].sum())
    assert n_valid > 0.8 * W * H and out["pixel_idx"].numel() == n_valid
    z = out["intersection_points"][:, 2]
    assert float((z - 5.0).abs().quantile(0.99)) < 0.02 and float(z.mean()) < 5.0     # in front of the sheet
    assert float(out["normals"][:, 2].abs().quantile(0.05)) > 0.95
    # subsampling keeps the request
    sub = levelset.level_surface_points_from_camera(pts, scaling, quats, strengths, cam, surface_levels=(0.3,),
                                                    n_surface_points=1000)[0.3]
    assert sub["valid"].numel() == 1000
