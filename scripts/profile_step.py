"""Two fwd+bwd steps of the bench workload (for ncu): python scripts/profile_step.py [P W H [steps]]
SGR_PROFILE_EXCHANGE=<views>: run the view-parallel exchange path on this one rank (parallel.ViewParallel(force=True)): the
same kernels as a multi-GPU step (factor-mode per-Gaussian backward + finalize) without NCCL, so ncu can capture them."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sugar_b200 import diff_gaussian_rasterization as mod, scenes
P, W, H = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (3_000_000, 1920, 1080)
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
sc = scenes.make_scene(P, W, H, seed=0)
dev = torch.device("cuda")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
params = {k: t(getattr(sc, k)).requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
means2D = torch.zeros_like(params["means3D"], requires_grad=True)
dL = t(scenes.upstream_grad(W, H))
st = mod.GaussianRasterizationSettings(H, W, sc.tanfovx, sc.tanfovy, torch.zeros(3, device=dev), 1.0, t(sc.viewmatrix),
                                       t(sc.projmatrix), 3, t(sc.campos), False, False)
import contextlib
ctx = contextlib.nullcontext()
if os.environ.get("SGR_PROFILE_EXCHANGE"):
    from sugar_b200 import parallel
    ctx = parallel.ViewParallel(force=True, chunks=int(os.environ.get("SGR_PROFILE_CHUNKS", "4"))).context()
with ctx:
    for _ in range(steps):
        color, radii = mod.GaussianRasterizer(st)(means3D=params["means3D"], means2D=means2D,
                                                  opacities=params["opacities"], shs=params["shs"],
                                                  scales=params["scales"], rotations=params["rotations"])
        torch.autograd.backward(color, dL)
        for p in params.values():
            p.grad = None
torch.cuda.synchronize()
print("done", int((radii > 0).sum()))
