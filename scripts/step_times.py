"""Per-step wall times of the bench workload and allocator statistics (how the allocator-cycle fix in
sugar_b200/_C.py was found): python scripts/step_times.py"""
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
N = 60
ev = [torch.cuda.Event(enable_timing=True) for _ in range(N + 1)]
cpu = []
allocs = []
ev[0].record()
for i in range(N):
    t0 = time.perf_counter(); step(); cpu.append(time.perf_counter() - t0)
    ev[i + 1].record()
    allocs.append(torch.cuda.memory_stats()['num_device_alloc'])
torch.cuda.synchronize()
gpu = [ev[i].elapsed_time(ev[i + 1]) for i in range(N)]
print("gpu ms/step:", " ".join(f"{x:.2f}" for x in gpu))
print("cpu ms/step:", " ".join(f"{x*1e3:.2f}" for x in cpu))
print("allocs:", " ".join(str(a) for a in allocs))
print(torch.cuda.memory_stats()["num_device_alloc"], torch.cuda.memory_stats()["reserved_bytes.all.current"] / 1e9)
