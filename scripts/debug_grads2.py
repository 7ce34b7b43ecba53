import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as h
import test_gpu_parity as tp
from sugar_b200 import diff_gaussian_rasterization as ours, scenes
ref = h.load_ref_module()
names = dict(means3D="dL_dmeans3D", means2D="dL_dmeans2D", opacities="dL_dopacity", shs="dL_dsh",
             colors_precomp="dL_dcolors", scales="dL_dscales", rotations="dL_drotations", cov3D_precomp="dL_dcov3D")
for case in tp.CASES:
    name, P, W, H, camera, use_sh, deg, covpre, bg = case
    sc = tp._scene(name, P, W, H, camera)
    dL = scenes.upstream_grad(W, H)
    cov3D = tp._cov_from_oracle(sc) if covpre else None
    opts = dict(use_sh=use_sh, sh_degree=deg, use_cov_precomp=covpre, cov3D=cov3D)
    a = h.run_module(ours, sc, bg, dL, **opts)
    b = h.run_module(ref, sc, bg, dL, **opts)
    b2 = h.run_module(ref, sc, bg, dL, **opts)
    fw, bw = h.run_oracle(sc, np.asarray(bg, np.float32), dL, **opts)
    for k in sorted(b["grads"]):
        ga, gb, gb2 = (x["grads"][k].cpu().numpy() for x in (a, b, b2))
        go = bw[names[k]].reshape(gb.shape)
        print(f"{name:14s} {k:14s} ours-ref {h.rel_err(ga, gb):.2e}  ref-ref {h.rel_err(gb2, gb):.2e}  orc-ref {h.rel_err(go, gb):.2e} ours-orc {h.rel_err(ga, go):.2e}")
