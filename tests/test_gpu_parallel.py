"""SH factor mode of the view-parallel step (sugar_b200/parallel.py) on one GPU: the dL_dsh rebuilt from
the per-view factors (sgr_sh_grad_from_factors) must equal the sum of the per-view dL_dsh of the
ordinary backward, and every other gradient must be unchanged by the mode."""
import numpy as np
import pytest

import helpers as h

pytestmark = pytest.mark.gpu


def _err(a, b):
    return h.rel_err(a.detach().float().cpu().numpy(), b.detach().float().cpu().numpy())


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_sh_factors_match_summed_dsh(deg):
    import torch
    from sugar_b200 import diff_gaussian_rasterization as mod
    from sugar_b200 import parallel, scenes
    P, W, H = 6000, 160, 96
    base = scenes.make_scene(P, W, H, seed=31 + deg, camera="posed")
    views = [base, scenes.with_camera_offset(base, 0.15, (0.3, -0.1, 0.2)),
             scenes.with_camera_offset(base, -0.2, (-0.4, 0.2, 0.5))]
    bg = (0.1, 0.2, 0.3)
    want, factors, campos = None, [], []
    for v, sc in enumerate(views):
        dL = scenes.upstream_grad(W, H, seed=5 + v)
        ref = h.run_module(mod, sc, bg, dL=dL, sh_degree=deg)["grads"]
        with parallel.sh_factor_mode():
            t = h.to_torch(sc)
            leaf = lambda x: x.clone().requires_grad_(True)
            ps = {k: leaf(t[k]) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
            means2D = torch.zeros_like(ps["means3D"], requires_grad=True)
            st = mod.GaussianRasterizationSettings(
                image_height=H, image_width=W, tanfovx=sc.tanfovx, tanfovy=sc.tanfovy,
                bg=torch.tensor(bg, device="cuda"), scale_modifier=1.0, viewmatrix=t["viewmatrix"],
                projmatrix=t["projmatrix"], sh_degree=deg, campos=t["campos"], prefiltered=False, debug=False)
            color, _ = mod.GaussianRasterizer(st)(means3D=ps["means3D"], means2D=means2D, opacities=ps["opacities"],
                                                  shs=ps["shs"], scales=ps["scales"], rotations=ps["rotations"])
            (color * torch.from_numpy(dL).cuda()).sum().backward()
            arena = parallel.GradArena(P, 16, "cuda")
            buf = arena._shared_base(ps)
            assert buf is not None
            o_col = arena.flat.numel() + 3 * P
            factors.append(arena._base[o_col:o_col + 3 * P].clone().view(P, 3))
            campos.append(t["campos"].clone())
            # single-process all_reduce_from == this view's own gradients, dL_dsh rebuilt from its factor
            arena.all_reduce_from(ps, campos=t["campos"], sh_degree=deg)
        for k in ("means3D", "opacities", "scales", "rotations"):
            assert _err(ps[k].grad, ref[k]) <= 1e-4, k  # fp32 atomics: run-to-run order differs
        assert _err(ps["shs"].grad, ref["shs"]) <= 1e-4
        want = ref["shs"].double() if want is None else want + ref["shs"].double()
    got = parallel.sh_grad_from_factors(torch.from_numpy(base.means3D).cuda(), torch.stack(campos).contiguous(),
                                        torch.stack(factors).contiguous(), 16, deg)
    assert _err(got, want) <= 1e-4
    used = (deg + 1) ** 2
    if used < 16:
        assert float(got[:, used:].abs().max()) == 0.0


def test_staged_backward_hook_sees_final_factors():
    """sgr_rasterize_backward_staged: at hook time the factors enqueued so far are already the final
    masked dL/dRGB, and the remaining gradients equal the unstaged factor-mode backward."""
    import torch
    from sugar_b200 import _C, parallel, scenes
    from sugar_b200 import diff_gaussian_rasterization as mod
    P, W, H, deg = 5000, 160, 96, 3
    sc = scenes.make_scene(P, W, H, seed=77, camera="posed")
    dL = torch.from_numpy(scenes.upstream_grad(W, H, seed=9)).cuda()
    t = h.to_torch(sc)

    def run(hook):
        ps = {k: t[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
        means2D = torch.zeros_like(ps["means3D"], requires_grad=True)
        st = mod.GaussianRasterizationSettings(
            image_height=H, image_width=W, tanfovx=sc.tanfovx, tanfovy=sc.tanfovy, bg=torch.zeros(3, device="cuda"),
            scale_modifier=1.0, viewmatrix=t["viewmatrix"], projmatrix=t["projmatrix"], sh_degree=deg,
            campos=t["campos"], prefiltered=False, debug=False)
        with parallel.sh_factor_mode():
            _C.FACTOR_HOOK = hook
            color, _ = mod.GaussianRasterizer(st)(means3D=ps["means3D"], means2D=means2D, opacities=ps["opacities"],
                                                  shs=ps["shs"], scales=ps["scales"], rotations=ps["rotations"])
            (color * dL).sum().backward()
            arena = parallel.GradArena(P, 16, "cuda")
            assert arena._shared_base(ps) is not None
            o_col = arena.flat.numel() + 3 * P
            final = arena._base[o_col:o_col + 3 * P].clone().view(P, 3)
        return ps, final

    seen = []
    ps_a, fin_a = run(lambda d: seen.append(d.clone()))   # clone is stream-ordered: snapshot at hook time
    ps_b, fin_b = run(None)
    assert len(seen) == 1 and seen[0].shape == (P, 3)
    assert torch.equal(seen[0], fin_a)
    assert _err(fin_a, fin_b) <= 1e-4 and float(fin_a.abs().max()) > 0
    for k in ("means3D", "opacities", "scales", "rotations"):
        assert _err(ps_a[k].grad, ps_b[k].grad) <= 1e-4, k
