"""Golden vectors for the level-set ray sampling from the reference's OWN code.

Runs in the build container (needs /root/reference):  python tests/golden/make_levelset_golden.py
The per-ray part of SuGaR.compute_level_surface_points_from_camera_fast is a block in the middle of a long
method whose first half needs pytorch3d cameras and a mesh rasterizer.  This script reads exactly
sugar_scene/sugar_model.py:1970-2081 from /root/reference at run time, dedents the lines and executes them
unchanged with the locals that precede them in the method (`all_world_points`, `closest_gaussians_idx`,
`gaussian_idx`, `fov_cameras`, the keyword flags at their defaults); `self.get_covariance` is the
reference's own method called unbound, pytorch3d's quaternion helpers are the restatements of
oracle/field_oracle.py.  Only the resulting numbers are stored.
Outputs: tests/golden/levelset_<case>.npz
"""
import os
import sys
import textwrap

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import field_oracle as fo  # noqa: E402
from make_field_golden import import_reference_sugar  # noqa: E402

CASES = {"k16": dict(P=600, N=900, K=16, seed=11, density_factor=1.0),
         "k8": dict(P=400, N=500, K=8, seed=12, density_factor=1.0)}
BLOCK = ("/root/reference/sugar_scene/sugar_model.py", 1970, 2081)
LEVELS = [0.1, 0.3, 0.5]


def make_inputs(cfg):
    """World points near the cloud's surface as seen from a camera: the samples of fo.make_case pushed a bit
    towards the camera, with the neighbours of the Gaussian they were drawn from."""
    case = fo.make_case(density_threshold=1.0, **cfg)
    cam = np.array([0.3, -0.2, -6.0], np.float32)
    return case, cam


def main():
    sm = import_reference_sugar()
    SuGaR = sm.SuGaR
    lines = open(BLOCK[0]).read().splitlines()[BLOCK[1] - 1:BLOCK[2]]
    code = compile(textwrap.dedent("\n".join(lines)), "sugar_model.py:1970-2081", "exec")
    for name, cfg in CASES.items():
        case, cam = make_inputs(cfg)
        t = lambda k: torch.from_numpy(case[k])

        class Fake:
            device = "cpu"
            points = t("points")
            scaling = t("scaling")
            quaternions = t("quaternions")
            strengths = t("strengths").view(-1, 1)
            knn_to_track = cfg["K"]

            def get_covariance(self, **kw):
                return SuGaR.get_covariance(self, **kw)

        class Cam:
            def get_camera_center(self):
                return torch.from_numpy(cam)[None]
        ns = dict(torch=torch, self=Fake(), fov_cameras=Cam(), all_world_points=t("x"),
                  closest_gaussians_idx=t("nbr_idx"), gaussian_idx=t("gaussian_idx"),
                  quaternion_apply=fo.quaternion_apply,
                  quaternion_invert=lambda q: q * q.new_tensor([1, -1, -1, -1]),
                  range_size=3.0, n_points_in_range=21, n_points_per_pass=2_000_000, density_factor=cfg["density_factor"],
                  compute_intersection_for_flat_gaussian=False, compute_flat_normals=False,
                  just_use_depth_as_level=False, surface_levels=LEVELS, return_pixel_idx=False,
                  return_gaussian_idx=True, return_normals=True)
        with torch.no_grad():
            exec(code, ns)
        out = {"camera_center": cam}
        for lv in LEVELS:
            o = ns["all_outputs"][lv]
            out[f"points_{lv}"] = o["intersection_points"].numpy()
            out[f"normals_{lv}"] = o["normals"].numpy()
            out[f"gaussian_idx_{lv}"] = o["gaussian_idx"].numpy()
            print(name, lv, o["intersection_points"].shape)
        np.savez_compressed(os.path.join(HERE, f"levelset_{name}.npz"), **out)


if __name__ == "__main__":
    main()
