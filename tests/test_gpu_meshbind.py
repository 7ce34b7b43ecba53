"""Mesh-bound Gaussians on the GPU (sugar_b200/meshbind.py, csrc/sgr_meshbind.cu) against goldens made by the
reference's own property code (sugar_model.py:384-398, 415-441, 443-479) and against the oracle in fp64 on a
larger soup; plus the refine step composed from it (sugar_b200/steps.py)."""
import os

import numpy as np
import pytest

import helpers as h

pytestmark = pytest.mark.gpu
GOLD = os.path.join(h.ROOT, "tests", "golden")


def _run(case, seed):
    import torch
    from oracle import meshbind_oracle as mo
    from sugar_b200 import meshbind
    leaf = {k: case[k].cuda().requires_grad_(True) for k in ("verts", "scales_raw", "complex_raw")}
    b = meshbind.bind_to_mesh(leaf["verts"], case["faces"].cuda(), case["bary"].cuda(), leaf["scales_raw"],
                              leaf["complex_raw"], case["thickness"])
    wp, ws, wq = (w.cuda() for w in mo.loss_weights(b.points.shape[0], seed))
    ((b.points * wp).sum() + (b.scaling * ws).sum() + (b.quaternions * wq).sum()).backward()
    out = dict(points=b.points, scaling=b.scaling, quaternions=b.quaternions)
    out.update({"g_" + k: v.grad for k, v in leaf.items()})
    return {k: v.detach().cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize("name", ["n6", "n1", "n3"])
def test_matches_reference_goldens(name):
    from oracle import meshbind_oracle as mo
    gold = np.load(os.path.join(GOLD, f"meshbind_{name}.npz"))
    F, V, n_per, seed = (int(v) for v in gold["cfg"])
    got = _run(mo.make_case(F=F, V=V, n_per=n_per, seed=seed), seed)
    for k in ("points", "scaling", "quaternions"):
        assert h.rel_err(got[k], gold[k]) <= 2e-6, k
    for k in ("g_verts", "g_scales_raw", "g_complex_raw"):
        assert h.rel_err(got[k], gold[k]) <= 1e-4, k


def test_large_soup_matches_fp64_oracle_and_unit_quaternions():
    import torch
    from oracle import meshbind_oracle as mo
    case = mo.make_case(F=20_000, V=9_000, n_per=6, seed=7)
    got = _run(case, 7)
    want = mo.values_and_grads({k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v)
                                for k, v in case.items()}, 7)
    for k in ("points", "scaling", "quaternions"):
        assert h.rel_err(got[k], want[k]) <= 2e-6, k
    for k in ("g_verts", "g_scales_raw", "g_complex_raw"):
        assert h.rel_err(got[k], want[k]) <= 1e-4, k
    assert np.abs(np.linalg.norm(got["quaternions"], axis=1) - 1).max() < 1e-5
    assert np.all(got["scaling"][:, 0] == np.float32(case["thickness"]))


def test_refine_step_renders_and_backpropagates_to_vertices():
    """steps.refine_step: bound Gaussians -> render -> L1 -> backward reaches the mesh vertices."""
    import torch
    from types import SimpleNamespace
    from sugar_b200 import meshbind, scenes, steps
    W, H, F, n = 320, 192, 8000, 6
    sc = scenes.make_scene(1000, W, H, seed=4)
    g = torch.Generator().manual_seed(0)
    centers = torch.stack([(torch.rand(F, generator=g) - 0.5) * 6, (torch.rand(F, generator=g) - 0.5) * 3.5,
                           4 + 4 * torch.rand(F, generator=g)], 1)
    tri = centers[:, None] + 0.05 * torch.randn(F, 3, 3, generator=g)
    verts = tri.reshape(-1, 3).cuda().requires_grad_(True)
    faces = torch.arange(3 * F).view(F, 3).cuda()
    P = F * n
    raw = dict(verts=verts, sh_dc=(0.5 * torch.randn(P, 1, 3, generator=g)).cuda().requires_grad_(True),
               sh_rest=(0.1 * torch.randn(P, 15, 3, generator=g)).cuda().requires_grad_(True),
               densities=torch.randn(P, 1, generator=g).cuda().requires_grad_(True),
               scales=(torch.randn(P, 2, generator=g) * 0.3 - 4.0).cuda().requires_grad_(True),
               quaternions=torch.randn(P, 2, generator=g).cuda().requires_grad_(True))
    mesh = SimpleNamespace(faces=faces, bary=meshbind.bary_coords(n, "cuda"), thickness=1e-5)
    cam = steps.camera_from_scene(sc, "cuda")
    gt = torch.rand(3, H, W, device="cuda")
    loss, stats = steps.refine_step(raw, mesh, cam, gt, steps.ours_ops())
    assert torch.isfinite(loss) and stats["visible"] > P // 4
    for k, v in raw.items():
        assert v.grad is not None and bool(torch.isfinite(v.grad).all()), k
    assert float(raw["verts"].grad.abs().max()) > 0 and float(raw["quaternions"].grad.abs().max()) > 0
