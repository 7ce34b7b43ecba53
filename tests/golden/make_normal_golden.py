"""Golden vectors for the "better normal" loss from the reference's OWN code.

Runs in the build container (needs /root/reference):  python tests/golden/make_normal_golden.py
The loss is not a function in the reference: it is a block inside the training loop
(sugar_trainers/coarse_sdf.py:688-716).  This script reads exactly those source lines from
/root/reference at run time, dedents them and executes them unchanged in a namespace that provides
what the loop has in scope at that point (`sugar`, `fields`, `sdf_samples`, `sdf_gaussian_idx`, the
flags of coarse_sdf.py:140-146); `sugar.get_normals` / `get_smallest_axis` are the reference's
SuGaR methods called unbound on a duck-typed object (as in make_field_golden.py).  Nothing of the
reference is copied into the repository; only the resulting numbers are stored.
Outputs: tests/golden/normal_<case>.npz (normals, per-sample loss, d(mean loss)/d quaternions).
"""
import os
import sys
import textwrap
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import field_oracle as fo  # noqa: E402
from make_field_golden import import_reference_sugar  # noqa: E402

CASES = {"c1_1k_2k": dict(P=1000, N=2000, K=16, seed=0, density_factor=1.0 / 16.0),
         "k8": dict(P=400, N=600, K=8, seed=5, density_factor=1.0 / 8.0)}
BLOCK = ("/root/reference/sugar_trainers/coarse_sdf.py", 688, 716)


def main():
    sm = import_reference_sugar()
    SuGaR = sm.SuGaR
    lines = open(BLOCK[0]).read().splitlines()[BLOCK[1] - 1:BLOCK[2]]
    code = compile(textwrap.dedent("\n".join(lines)), "coarse_sdf.py:688-716", "exec")
    for name, cfg in CASES.items():
        case = fo.make_case(density_threshold=1.0, **cfg)
        t = lambda k: torch.from_numpy(case[k])
        quats = t("quaternions").clone().requires_grad_(True)
        P, K = cfg["P"], cfg["K"]

        class Fake:
            binded_to_surface_mesh = False
            points = t("points")
            scaling = t("scaling")
            quaternions = quats
            knn_idx = fo.knn_idx(t("points"), K)

            def get_smallest_axis(self, **kw):
                return SuGaR.get_smallest_axis(self, **kw)

            def get_normals(self, **kw):
                return SuGaR.get_normals(self, **kw)
        sugar = Fake()
        gi = t("gaussian_idx")
        assert torch.equal(sugar.knn_idx[gi], t("nbr_idx"))
        with torch.no_grad():
            opac = fo.field_values_torch(t("x"), t("nbr_idx"), t("points"), t("scaling"), t("quaternions"),
                                         t("strengths"), cfg["density_factor"])["closest_gaussian_opacities"]
        ns = dict(torch=torch, sugar=sugar, fields={"closest_gaussian_opacities": opac}, sdf_samples=t("x"),
                  sdf_gaussian_idx=gi, use_sdf_better_normal_loss=True, iteration=10, start_sdf_better_normal_from=0,
                  sdf_better_normal_gradient_through_normal_only=True, sdf_better_normal_factor=1.0, loss=0.0,
                  CONSOLE=mock.MagicMock())
        exec(code, ns)
        ns["loss"].backward()
        out = dict(normals=sugar.get_normals().detach().numpy(), nbr_opacity=opac.numpy(),
                   loss=ns["sdf_better_normal_loss"].detach().numpy(), grad_quaternions=quats.grad.numpy())
        np.savez_compressed(os.path.join(HERE, f"normal_{name}.npz"), **out)
        print(name, float(out["loss"].mean()), float(np.abs(out["grad_quaternions"]).max()))


if __name__ == "__main__":
    main()
