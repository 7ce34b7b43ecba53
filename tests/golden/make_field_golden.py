"""Golden vectors for the density / SDF field from the reference's OWN code.

Runs in the build container (needs /root/reference):  python tests/golden/make_field_golden.py
sugar_scene/sugar_model.py cannot be imported as is (open3d, pytorch3d, simple_knn, the
rasterizer extension and plotly are absent), so those modules are stubbed in sys.modules; the only
third-party function the field path really calls, pytorch3d's quaternion_to_matrix, is the
restatement in oracle/field_oracle.py.  SuGaR.get_field_values / get_covariance / get_beta are then
called UNBOUND on a duck-typed object, i.e. the arithmetic executed is the reference's.
Outputs: tests/golden/field_<case>.npz (values + autograd gradients of a fixed scalar loss).
"""
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import field_oracle as fo  # noqa: E402

CASES = {"c1_1k_2k": dict(P=1000, N=2000, K=16, seed=0, density_factor=1.0 / 16.0),
         "dense_factor1": dict(P=300, N=500, K=16, seed=3, density_factor=1.0),
         "k8": dict(P=400, N=600, K=8, seed=5, density_factor=1.0 / 8.0)}


def import_reference_sugar():
    for name in ["open3d", "pytorch3d", "pytorch3d.renderer", "pytorch3d.structures", "pytorch3d.ops",
                 "pytorch3d.renderer.blending", "pytorch3d.renderer.cameras", "pytorch3d.loss", "pytorch3d.io",
                 "simple_knn", "simple_knn._C", "diff_gaussian_rasterization", "plotly", "plotly.graph_objects",
                 "plyfile", "sugar_scene.gs_model", "sugar_scene.cameras", "rich", "rich.console"]:
        sys.modules.setdefault(name, mock.MagicMock())
    tr = types.ModuleType("pytorch3d.transforms")
    tr.quaternion_to_matrix = fo.quaternion_to_matrix
    tr.quaternion_apply = fo.quaternion_apply
    tr.quaternion_invert = lambda q: q * q.new_tensor([1, -1, -1, -1])
    tr.matrix_to_quaternion = mock.MagicMock()
    sys.modules["pytorch3d.transforms"] = tr
    sys.path.insert(0, "/root/reference")
    import sugar_scene.sugar_model as sm
    return sm


def main():
    sm = import_reference_sugar()
    SuGaR = sm.SuGaR
    for name, cfg in CASES.items():
        case = fo.make_case(density_threshold=1.0, **cfg)
        leaf = {k: torch.from_numpy(case[k]).clone().requires_grad_(True)
                for k in ("x", "points", "scaling", "quaternions", "strengths")}
        nbr = torch.from_numpy(case["nbr_idx"])

        class Fake:
            beta_mode = "average"
            device = "cpu"
            points = leaf["points"]
            scaling = leaf["scaling"]
            quaternions = leaf["quaternions"]
            strengths = leaf["strengths"].view(-1, 1)

            def get_covariance(self, **kw):
                return SuGaR.get_covariance(self, **kw)

            def get_beta(self, *a, **kw):
                return SuGaR.get_beta(self, *a, **kw)
        fields = SuGaR.get_field_values(Fake(), leaf["x"], closest_gaussians_idx=nbr, return_sdf=True,
                                        density_threshold=case["density_threshold"],
                                        density_factor=case["density_factor"], return_closest_gaussian_opacities=True,
                                        return_beta=True, return_sdf_grad=True, sdf_grad_max_value=10.)
        sdf_grad = fields.pop("sdf_grad").detach()   # a value-only extra: not part of the scalar loss below
        g = torch.Generator().manual_seed(99)
        w = {k: torch.randn(fields[k].shape, generator=g) for k in ("density", "sdf", "beta", "closest_gaussian_opacities")}
        loss = sum((fields[k] * w[k]).sum() for k in w)
        loss.backward()
        out = {k: v.detach().numpy() for k, v in fields.items()}
        out["sdf_grad"] = sdf_grad.numpy()
        out.update({"w_" + k: v.numpy() for k, v in w.items()})
        out.update({"grad_" + k: v.grad.numpy() for k, v in leaf.items()})
        np.savez_compressed(os.path.join(HERE, f"field_{name}.npz"), **out)
        print(name, {k: float(np.abs(v).max()) for k, v in out.items() if k.startswith("grad_")})


if __name__ == "__main__":
    main()
