"""Golden vectors for mesh-bound Gaussians from the reference's OWN property code.

Runs in the build container (needs /root/reference):  python tests/golden/make_meshbind_golden.py
SuGaR.points / .scaling / .quaternions (sugar_scene/sugar_model.py:384-398, 415-441, 443-479) are called
UNBOUND on a duck-typed object that carries the fields those properties read; the module's missing third-party
imports are stubbed (tests/golden/make_field_golden.py), and the two pytorch3d functions the bound path really
calls -- Meshes.faces_normals_list and matrix_to_quaternion -- are the restatements in
oracle/meshbind_oracle.py.  Outputs: tests/golden/meshbind_<case>.npz (values + autograd gradients).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import meshbind_oracle as mo  # noqa: E402
from make_field_golden import import_reference_sugar  # noqa: E402

CASES = {"n6": dict(F=150, V=90, n_per=6, seed=0), "n1": dict(F=80, V=60, n_per=1, seed=1),
         "n3": dict(F=60, V=45, n_per=3, seed=2)}


def main():
    sm = import_reference_sugar()
    sm.matrix_to_quaternion = mo.matrix_to_quaternion     # `from pytorch3d.transforms import ...` bound it at import
    SuGaR = sm.SuGaR
    for name, cfg in CASES.items():
        case = mo.make_case(**cfg)
        leaf = {k: case[k].clone().requires_grad_(True) for k in ("verts", "scales_raw", "complex_raw")}
        F, n = case["faces"].shape[0], case["bary"].shape[0]

        class Mesh:
            def faces_normals_list(self_inner):
                return [mo.faces_normals(leaf["verts"], case["faces"])]

        class Fake:
            binded_to_surface_mesh = True
            editable = False
            learnable_positions = True
            device = "cpu"
            _points = leaf["verts"]
            _surface_mesh_faces = case["faces"]
            surface_triangle_bary_coords = case["bary"][..., None]
            n_gaussians_per_surface_triangle = n
            _n_points = F * n
            _scales = leaf["scales_raw"]
            _quaternions = leaf["complex_raw"]
            surface_mesh_thickness = torch.tensor(case["thickness"])
            scale_activation = staticmethod(torch.exp)
            surface_mesh = Mesh()
        fake = Fake()
        p = SuGaR.points.fget(fake)
        s = SuGaR.scaling.fget(fake)
        q = SuGaR.quaternions.fget(fake)
        wp, ws, wq = mo.loss_weights(F * n, cfg["seed"])
        ((p * wp).sum() + (s * ws).sum() + (q * wq).sum()).backward()
        out = dict(points=p, scaling=s, quaternions=q, g_verts=leaf["verts"].grad, g_scales_raw=leaf["scales_raw"].grad,
                   g_complex_raw=leaf["complex_raw"].grad)
        np.savez_compressed(os.path.join(HERE, f"meshbind_{name}.npz"), cfg=np.array([cfg[k] for k in ("F", "V", "n_per", "seed")]),
                            **{k: v.detach().numpy() for k, v in out.items()})
        print(name, {k: tuple(v.shape) for k, v in out.items()})


if __name__ == "__main__":
    main()
