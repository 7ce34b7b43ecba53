import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import helpers as h
from sugar_b200 import diff_gaussian_rasterization as ours, scenes
P, W, H = 3000, 160, 96
use_sh = len(sys.argv) > 1 and sys.argv[1] == "sh"
sc = scenes.make_scene(P, W, H, seed=5, camera="posed")
dL = scenes.upstream_grad(W, H)
a = h.run_module(ours, sc, (0.1, 0.2, 0.3), dL, use_sh=use_sh, sh_degree=3)
torch.cuda.synchronize()
print("img finite", bool(torch.isfinite(a["color"]).all()), {k: bool(torch.isfinite(v).all()) for k, v in a["grads"].items()})
