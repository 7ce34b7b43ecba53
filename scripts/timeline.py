"""Kernel timeline of one fwd+bwd step from the library's own CUDA-event log (sgr_profile_timeline):
python scripts/timeline.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sugar_b200 import diff_gaussian_rasterization as mod, scenes, _C, _lib
P, W, H = 3_000_000, 1920, 1080
sc = scenes.make_scene(P, W, H, seed=0)
dev = torch.device("cuda")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
params = {k: t(getattr(sc, k)).requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
means2D = torch.zeros_like(params["means3D"], requires_grad=True)
dL = t(scenes.upstream_grad(W, H))
st = mod.GaussianRasterizationSettings(H, W, sc.tanfovx, sc.tanfovy, torch.zeros(3, device=dev), 1.0, t(sc.viewmatrix),
                                       t(sc.projmatrix), 3, t(sc.campos), False, False)
def step():
    color, radii = mod.GaussianRasterizer(st)(means3D=params["means3D"], means2D=means2D, opacities=params["opacities"],
                                              shs=params["shs"], scales=params["scales"], rotations=params["rotations"])
    torch.autograd.backward(color, dL)
    for p in params.values():
        p.grad = None
    means2D.grad = None
_lib.profile(True)
for _ in range(6): step()
torch.cuda.synchronize(); _lib.profile_read()
for _ in range(4): step()
torch.cuda.synchronize()
tl = _lib.profile_timeline()
prev_end = None
for name, b, e in tl:
    gap = (b - prev_end) if prev_end is not None else 0.0
    print(f"{name:22s} start {b:8.3f}  dur {e-b:7.3f}  gap_before {gap:7.3f}")
    prev_end = e
