"""CPU: the field oracle (oracle/field_oracle.py) against golden vectors produced by the
reference's own SuGaR.get_field_values code (tests/golden/make_field_golden.py)."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_field_golden import CASES  # noqa: E402
from oracle import field_oracle as fo  # noqa: E402


@pytest.mark.parametrize("name", sorted(CASES))
def test_field_oracle_matches_reference_code(name):
    cfg = CASES[name]
    gold = np.load(os.path.join(HERE, "golden", f"field_{name}.npz"))
    case = fo.make_case(density_threshold=1.0, **cfg)
    leaf = {k: torch.from_numpy(case[k]).clone().requires_grad_(True)
            for k in ("x", "points", "scaling", "quaternions", "strengths")}
    out = fo.field_values_torch(leaf["x"], torch.from_numpy(case["nbr_idx"]), leaf["points"], leaf["scaling"],
                                leaf["quaternions"], leaf["strengths"], case["density_factor"],
                                case["density_threshold"])
    for k in ("density", "sdf", "beta", "closest_gaussian_opacities"):
        assert np.allclose(out[k].detach().numpy(), gold[k], rtol=1e-6, atol=1e-7), k
    loss = sum((out[k] * torch.from_numpy(gold["w_" + k])).sum() for k in ("density", "sdf", "beta",
                                                                            "closest_gaussian_opacities"))
    loss.backward()
    for k, v in leaf.items():
        g = gold["grad_" + k]
        if not np.isfinite(g).all():
            # density >= 1 makes the reference itself produce NaN (sqrt'(0) through the straight-through
            # clamp, sugar_model.py:1280-1306); nothing to compare against
            assert cfg["density_factor"] == 1.0
            continue
        assert np.allclose(v.grad.numpy(), g, rtol=1e-5, atol=1e-6), k


def test_c1_config_sizes():
    """BASELINE config 1: 1k Gaussians, 2k sample points, pure PyTorch on CPU."""
    case = fo.make_case(P=1000, N=2000, K=16, seed=0, density_factor=1 / 16)
    out = fo.field_values(**case)
    assert out["density"].shape == (2000,) and out["closest_gaussian_opacities"].shape == (2000, 16)
    assert np.isfinite(out["sdf"]).all() and (out["density"] >= 0).all() and (out["density"] <= 1.0).all()
    assert (case["nbr_idx"][:, 0] >= 0).all()


def test_sampling_matches_reference_semantics():
    """field.sample_points_in_gaussians vs the oracle's restatement of sugar_model.py:885-928 (same RNG stream)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("sgr_field_nolib", os.path.join(HERE, "..", "sugar_b200", "field.py"))
    src = open(spec.origin).read()
    # field.py imports the CUDA library at module import; only its pure-torch helpers are needed here
    ns = {}
    head = src.index("def quaternion_apply")
    exec("import torch\n" + src[head:], ns)
    g = torch.Generator().manual_seed(5)
    P, N = 200, 1000
    points = torch.randn(P, 3, generator=g); scaling = torch.exp(torch.randn(P, 3, generator=g) - 2)
    q = torch.nn.functional.normalize(torch.randn(P, 4, generator=g), dim=-1); st = torch.rand(P, generator=g)
    g1, g2 = torch.Generator().manual_seed(9), torch.Generator().manual_seed(9)
    x, idx = ns["sample_points_in_gaussians"](points, scaling, q, st, N, 1.5, generator=g1)
    areas = (scaling[:, 0] * scaling[:, 1] * scaling[:, 2]).abs()
    gi = torch.multinomial(areas / areas.sum(), N, replacement=True, generator=g2)
    ref = points[gi] + fo.quaternion_apply(q[gi], 1.5 * scaling[gi] * torch.randn(N, 3, generator=g2))
    assert torch.equal(idx, gi) and torch.allclose(x, ref, atol=1e-6)


@pytest.mark.parametrize("name", ["c1_1k_2k", "k8"])
def test_normal_loss_oracle_matches_reference_code(name):
    """oracle better_normal_loss_torch vs the trainer's own lines (tests/golden/make_normal_golden.py)."""
    cfg = CASES[name]
    gold = np.load(os.path.join(HERE, "golden", f"normal_{name}.npz"))
    case = fo.make_case(density_threshold=1.0, **cfg)
    t = lambda k: torch.from_numpy(case[k])
    q = t("quaternions").clone().requires_grad_(True)
    assert np.allclose(fo.smallest_axis(t("scaling"), q).detach().numpy(), gold["normals"], rtol=1e-6, atol=1e-7)
    loss = fo.better_normal_loss_torch(t("x"), t("gaussian_idx"), t("nbr_idx"), t("points"), t("scaling"), q,
                                       torch.from_numpy(gold["nbr_opacity"]))
    assert np.allclose(loss.detach().numpy(), gold["loss"], rtol=1e-5, atol=1e-7)
    loss.mean().backward()
    assert np.allclose(q.grad.numpy(), gold["grad_quaternions"], rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("name", ["k16", "k8"])
def test_levelset_oracle_matches_reference_code(name):
    """oracle level_surface_points_torch vs sugar_model.py:1970-2081 executed as it stands
    (tests/golden/make_levelset_golden.py)."""
    from make_levelset_golden import CASES as LCASES, LEVELS, make_inputs
    gold = np.load(os.path.join(HERE, "golden", f"levelset_{name}.npz"))
    case, cam = make_inputs(LCASES[name])
    t = lambda k: torch.from_numpy(case[k])
    out = fo.level_surface_points_torch(t("x"), torch.from_numpy(cam), t("nbr_idx"), t("points"), t("scaling"),
                                        t("quaternions"), t("strengths"), surface_levels=LEVELS,
                                        density_factor=LCASES[name]["density_factor"])
    for lv in LEVELS:
        o = out[lv]
        assert np.array_equal(t("gaussian_idx")[o["valid"]].numpy(), gold[f"gaussian_idx_{lv}"])
        assert np.allclose(o["intersection_points"].numpy(), gold[f"points_{lv}"], rtol=1e-5, atol=1e-6)
        assert np.allclose(o["normals"].numpy(), gold[f"normals_{lv}"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", ["c1_1k_2k", "k8"])
def test_sdf_grad_matches_reference_code(name):
    """return_sdf_grad (sugar_model.py:1307-1314), a value the reference never differentiates through."""
    cfg = CASES[name]
    gold = np.load(os.path.join(HERE, "golden", f"field_{name}.npz"))
    case = fo.make_case(density_threshold=1.0, **cfg)
    t = lambda k: torch.from_numpy(case[k])
    out = fo.field_values_torch(t("x"), t("nbr_idx"), t("points"), t("scaling"), t("quaternions"), t("strengths"),
                                case["density_factor"], case["density_threshold"], return_sdf_grad=True)
    assert np.allclose(out["sdf_grad"].numpy(), gold["sdf_grad"], rtol=1e-5, atol=1e-6)
