"""View-parallel exchange (sugar_b200/parallel.py) on ONE GPU: `ViewParallel(force=True)` drives the same
record / chunk / factor / finalize path as a multi-rank run, with the collectives left out, so its gradients
must equal the plain backward's.  The multi-rank numbers are checked by bench.py itself at N > 1 (`exchange_check`
in its JSON line) and by scripts/check_view_parallel.py under torchrun."""
import numpy as np
import pytest

import helpers as h

pytestmark = pytest.mark.gpu


def _err(a, b):
    return h.rel_err(a.detach().float().cpu().numpy(), b.detach().float().cpu().numpy())


def _render_loss(mod, t, sc, dL, deg, ps, bg=(0.1, 0.2, 0.3), colors=None):
    import torch
    means2D = torch.zeros_like(ps["means3D"], requires_grad=True)
    st = mod.GaussianRasterizationSettings(
        image_height=sc.height, image_width=sc.width, tanfovx=sc.tanfovx, tanfovy=sc.tanfovy,
        bg=torch.tensor(bg, device="cuda"), scale_modifier=1.0, viewmatrix=t["viewmatrix"], projmatrix=t["projmatrix"],
        sh_degree=deg, campos=t["campos"], prefiltered=False, debug=False)
    kw = dict(colors_precomp=colors) if colors is not None else dict(shs=ps["shs"])
    color, radii = mod.GaussianRasterizer(st)(means3D=ps["means3D"], means2D=means2D, opacities=ps["opacities"],
                                              scales=ps["scales"], rotations=ps["rotations"], **kw)
    return (color * dL).sum(), means2D


@pytest.mark.parametrize("deg,chunks,factors,side", [(3, 4, True, False), (1, 3, True, False), (0, 1, True, False),
                                                     (3, 4, False, False), (2, 7, True, False), (3, 4, True, True),
                                                     (2, 3, False, True)])
@pytest.mark.parametrize("peer", [True, False])
def test_forced_exchange_matches_plain_backward(deg, chunks, factors, side, peer):
    import torch
    from sugar_b200 import diff_gaussian_rasterization as mod
    from sugar_b200 import parallel, scenes
    P, W, H = 6001, 160, 96          # not a multiple of the CTA size: partial last block, ragged chunks
    sc = scenes.make_scene(P, W, H, seed=31 + deg, camera="posed")
    dL = torch.from_numpy(scenes.upstream_grad(W, H, seed=5)).cuda()
    t = h.to_torch(sc)
    leaf = lambda x: x.clone().requires_grad_(True)
    names = ("means3D", "opacities", "shs", "scales", "rotations")
    ps_a = {k: leaf(t[k]) for k in names}
    loss, m2a = _render_loss(mod, t, sc, dL, deg, ps_a)
    loss.backward()
    # peer=True: the peer-memory exchange (csrc/sgr_peer.cu) with this GPU as its only rank -- flags, two-shot
    # record reduce, pointer-table finalize, both side streams
    vp = parallel.ViewParallel(sh_factors=factors, chunks=chunks, scale=0.5, force=True, side_stream=side, peer=peer)
    for _ in range(3 if peer else 1):      # step parity alternates the factor blocks
        ps_b = {k: leaf(t[k]) for k in names}
        with vp.context():
            loss, m2b = _render_loss(mod, t, sc, dL, deg, ps_b)
            loss.backward()
    assert vp.stats["backwards"] == (3 if peer else 1)
    assert any(v is not None for v in vp._peer_states.values()) == (peer and factors)
    vp.close()
    for k in names:
        assert _err(ps_b[k].grad, 0.5 * ps_a[k].grad) <= 1e-4, k   # fp32 atomics: run-to-run order differs
    assert _err(m2b.grad, m2a.grad) <= 1e-4                        # per-view statistic: neither summed nor scaled
    used = (deg + 1) ** 2
    if used < 16:
        assert float(ps_b["shs"].grad[:, used:].abs().max()) == 0.0


def test_exchange_with_precomputed_colours_and_activations_in_front():
    """The trainers' path: raw parameters -> exp / sigmoid / normalize (+ python colours) -> rasterizer.  The
    exchange acts on the op's output gradients, so non-leaf inputs need nothing special."""
    import torch
    from sugar_b200 import diff_gaussian_rasterization as mod
    from sugar_b200 import parallel, scenes
    P, W, H = 5000, 160, 96
    sc = scenes.make_scene(P, W, H, seed=3, camera="posed")
    dL = torch.from_numpy(scenes.upstream_grad(W, H, seed=9)).cuda()
    t = h.to_torch(sc)

    def raw():
        r = dict(means3D=t["means3D"].clone(), log_scales=t["scales"].log(), quats=t["rotations"] * 1.7,
                 logit=torch.logit(t["opacities"].clamp(1e-4, 1 - 1e-4)), colors=t["colors_precomp"].clone())
        return {k: v.requires_grad_(True) for k, v in r.items()}

    def run(r):
        ps = dict(means3D=r["means3D"], opacities=torch.sigmoid(r["logit"]), scales=torch.exp(r["log_scales"]),
                  rotations=torch.nn.functional.normalize(r["quats"], dim=-1))
        loss, _ = _render_loss(mod, t, sc, dL, 0, ps, colors=r["colors"] * 1.0)
        loss.backward()

    a = raw()
    run(a)
    b = raw()
    vp = parallel.ViewParallel(chunks=4, force=True)
    with vp.context():
        run(b)
    for k in a:
        assert _err(b[k].grad, a[k].grad) <= 1e-4, k


def test_two_backwards_per_step_accumulate():
    """Two views rendered by one rank in one step (gradient accumulation): each backward runs its own exchange."""
    import torch
    from sugar_b200 import diff_gaussian_rasterization as mod
    from sugar_b200 import parallel, scenes
    P, W, H, deg = 5000, 160, 96, 3
    sc1 = scenes.make_scene(P, W, H, seed=77, camera="posed")
    sc2 = scenes.with_camera_offset(sc1, 0.15, (0.3, -0.1, 0.2))
    dL = torch.from_numpy(scenes.upstream_grad(W, H, seed=9)).cuda()
    names = ("means3D", "opacities", "shs", "scales", "rotations")
    t1, t2 = h.to_torch(sc1), h.to_torch(sc2)

    def both(ctxmgr):
        ps = {k: t1[k].clone().requires_grad_(True) for k in names}
        with ctxmgr:
            for t, sc in ((t1, sc1), (t2, sc2)):
                loss, _ = _render_loss(mod, t, sc, dL, deg, ps)
                loss.backward()
        return ps
    import contextlib
    a = both(contextlib.nullcontext())
    vp = parallel.ViewParallel(chunks=2, force=True)
    b = both(vp.context())
    assert vp.stats["backwards"] == 2
    for k in names:
        assert _err(b[k].grad, a[k].grad) <= 1e-4, k


def test_sh_factors_of_three_views_rebuild_the_summed_dsh():
    """sgr_sh_grad_from_factors on the gathered factors of several views == the sum of the views' dL_dsh; the
    factor of a view is the dL_dcolors the backward returns with SH colours (clamp-masked dL/dRGB)."""
    import torch
    from sugar_b200 import _C, parallel, scenes
    P, W, H, deg = 6000, 160, 96, 3
    base = scenes.make_scene(P, W, H, seed=34, camera="posed")
    views = [base, scenes.with_camera_offset(base, 0.15, (0.3, -0.1, 0.2)),
             scenes.with_camera_offset(base, -0.2, (-0.4, 0.2, 0.5))]
    want, factors, campos = None, [], []
    E = torch.Tensor([])
    for v, sc in enumerate(views):
        t = h.to_torch(sc)
        dL = torch.from_numpy(scenes.upstream_grad(W, H, seed=5 + v)).cuda()
        bg = torch.zeros(3, device="cuda")
        R, color, radii, geom, binning, img = _C.rasterize_gaussians(
            bg, t["means3D"], E, t["opacities"], t["scales"], t["rotations"], 1.0, E, t["viewmatrix"], t["projmatrix"],
            sc.tanfovx, sc.tanfovy, H, W, t["shs"], deg, t["campos"], False, False)
        g = _C.rasterize_gaussians_backward(bg, t["means3D"], radii, E, t["scales"], t["rotations"], 1.0, E,
                                            t["viewmatrix"], t["projmatrix"], sc.tanfovx, sc.tanfovy, dL, t["shs"], deg,
                                            t["campos"], geom, R, binning, img, False)
        factors.append(g[1].clone())
        campos.append(t["campos"].clone())
        want = g[5].double() if want is None else want + g[5].double()
    got = parallel.sh_grad_from_factors(torch.from_numpy(base.means3D).cuda(), torch.stack(campos).contiguous(),
                                        torch.stack(factors).contiguous(), 16, deg)
    assert _err(got, want) <= 1e-4


def test_peer_exchange_with_three_emulated_ranks():
    """The peer-memory exchange (csrc/sgr_peer.cu) as rank 0 of THREE, on one GPU: ranks 1 and 2 are static stand-ins --
    their factor blocks and record arrays are filled from plain backwards of their views, their flag words are preset,
    and this process reduces their record slices as well.  Everything the real multi-rank run executes runs here: the
    flag waits, the per-Gaussian pass summing three views' SH gradients through the pointer table, the two-shot reduce
    over three record arrays, the split.  The result must be scale x (sum of the three views' plain gradients)."""
    import ctypes as C
    import torch
    from sugar_b200 import _C, _lib, diff_gaussian_rasterization as mod
    from sugar_b200 import parallel, scenes
    P, W, H, deg, scale = 6001, 160, 96, 3, 0.25
    base = scenes.make_scene(P, W, H, seed=41, camera="posed")
    views = [base, scenes.with_camera_offset(base, 0.15, (0.3, -0.1, 0.2)),
             scenes.with_camera_offset(base, -0.2, (-0.4, 0.2, 0.5))]
    names = ("means3D", "opacities", "shs", "scales", "rotations")
    E = torch.Tensor([])
    plain, factors, records = [], [], []
    for v, sc in enumerate(views):
        t = h.to_torch(sc)
        dL = torch.from_numpy(scenes.upstream_grad(W, H, seed=5 + v)).cuda()
        bg = torch.zeros(3, device="cuda")
        R, color, radii, geom, binning, img = _C.rasterize_gaussians(
            bg, t["means3D"], E, t["opacities"], t["scales"], t["rotations"], 1.0, E, t["viewmatrix"], t["projmatrix"],
            sc.tanfovx, sc.tanfovy, H, W, t["shs"], deg, t["campos"], False, False)
        g = _C.rasterize_gaussians_backward(bg, t["means3D"], radii, E, t["scales"], t["rotations"], 1.0, E,
                                            t["viewmatrix"], t["projmatrix"], sc.tanfovx, sc.tanfovy, dL, t["shs"], deg,
                                            t["campos"], geom, R, binning, img, False)
        plain.append({"means3D": g[3].double(), "opacities": g[2].double(), "shs": g[5].double(), "scales": g[6].double(),
                      "rotations": g[7].double()})
        factors.append(torch.cat([g[1].reshape(-1), t["campos"].reshape(3), torch.zeros(1, device="cuda")]).contiguous())
        records.append(torch.cat([g[3], g[2], g[6], g[7]], dim=1).contiguous())     # the 44-byte record layout
    want = {k: scale * sum(p[k] for p in plain) for k in names}

    vp = parallel.ViewParallel(chunks=3, scale=scale, force=True, peer=True)
    vp._nstage = 3      # staging arrays for three source ranks in this (single) rank's buffer
    st = vp._peer_state(_lib.lib, _lib.check, P, torch.device("cuda", torch.cuda.current_device()))
    # stand-ins for ranks 1 and 2: their records sit in this rank's staging arrays 1 and 2 (as if their per-Gaussian
    # passes had stored them), their sum arrays are plain local tensors, their factor blocks are static
    stage = lambda j: torch.as_tensor(parallel._DevMem(st.base + st.off["STAGE"] + j * st.stage_stride, 11 * P),
                                      device="cuda")
    stage(1).copy_(records[1].reshape(-1))
    stage(2).copy_(records[2].reshape(-1))
    S_fake = [torch.zeros(11 * P + 16, device="cuda"), torch.zeros(11 * P + 16, device="cuda")]
    tab = lambda own, others: torch.tensor([int(own)] + [o.data_ptr() for o in others], dtype=torch.int64, device="cuda")
    st.world = 3
    st.F_tab = [tab(st.F_tab[0][0], factors[1:]), tab(st.F_tab[1][0], factors[1:])]
    st.S_tab = tab(st.S_tab[0], S_fake)
    st.stage_tab = torch.tensor([st.R_ptr] * 3, dtype=torch.int64, device="cuda")   # every owner's array for OUR records = ours
    st.flag_tab = torch.tensor([st.base] * 3, dtype=torch.int64, device="cuda")   # "every rank's flags" = ours
    flags = torch.as_tensor(type("M", (), {"__cuda_array_interface__": {
        "shape": (64, 64), "typestr": "<i4", "data": (st.base, False), "version": 2}})(), device="cuda")
    flags[:63, 1:3] = 1 << 30   # ranks 1 and 2 have "already signalled" every slot of every step (row 63: CTA counters)
    vp._emulated = (1, 2)
    t0, sc0 = h.to_torch(views[0]), views[0]
    dL0 = torch.from_numpy(scenes.upstream_grad(W, H, seed=5)).cuda()
    for step in range(2):       # both factor-block parities
        ps = {k: t0[k].clone().requires_grad_(True) for k in names}
        with vp.context():
            loss, _ = _render_loss(mod, t0, sc0, dL0, deg, ps, bg=(0.0, 0.0, 0.0))
            loss.backward()
        for k in names:
            assert _err(ps[k].grad, want[k].float()) <= 1e-4, (step, k)
        for S in S_fake:        # every rank's sum array received every slice
            assert _err(S[:11 * P], (records[0].double() + records[1].double() + records[2].double()).float().reshape(-1)) <= 1e-4
    torch.cuda.synchronize()
    st.world = 1
    vp.close()
