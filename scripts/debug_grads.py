import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as h
from sugar_b200 import diff_gaussian_rasterization as ours, scenes
ref = h.load_ref_module()
P, W, H = 4000, 200, 120
sc = scenes.make_scene(P, W, H, seed=5, camera="posed")
dL = scenes.upstream_grad(W, H)
for use_sh in (True, False):
    a = h.run_module(ours, sc, (0.1, 0.2, 0.3), dL, use_sh=use_sh, sh_degree=3)
    b = h.run_module(ref, sc, (0.1, 0.2, 0.3), dL, use_sh=use_sh, sh_degree=3)
    fw, bw = h.run_oracle(sc, np.array((0.1, 0.2, 0.3), np.float32), dL, use_sh=use_sh, sh_degree=3)
    names = dict(means3D="dL_dmeans3D", means2D="dL_dmeans2D", opacities="dL_dopacity", shs="dL_dsh",
                 colors_precomp="dL_dcolors", scales="dL_dscales", rotations="dL_drotations")
    for k in sorted(b["grads"]):
        ga, gb = a["grads"][k].cpu().numpy().astype(np.float64), b["grads"][k].cpu().numpy().astype(np.float64)
        go = bw[names[k]].reshape(gb.shape).astype(np.float64)
        cos = (ga * gb).sum() / (np.linalg.norm(ga) * np.linalg.norm(gb) + 1e-300)
        print(f"sh={use_sh} {k:16s} |ours|={np.abs(ga).max():.3e} |ref|={np.abs(gb).max():.3e} |orc|={np.abs(go).max():.3e} "
              f"rel(ours,ref)={h.rel_err(ga, gb):.2e} rel(orc,ref)={h.rel_err(go, gb):.2e} cos={cos:.6f}")
        if h.rel_err(ga, gb) > 1e-3:
            i = np.unravel_index(np.abs(ga - gb).argmax(), ga.shape)
            print("    worst", i, ga[i], gb[i], go[i], " rows:", ga[i[0]].ravel()[:6], gb[i[0]].ravel()[:6])
