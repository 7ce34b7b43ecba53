import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sugar_b200 import diff_gaussian_rasterization as mod, scenes, _C
P, W, H = 3_000_000, 1920, 1080
sc = scenes.make_scene(P, W, H, seed=0)
dev = torch.device("cuda")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
params = {k: t(getattr(sc, k)).requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
means2D = torch.zeros_like(params["means3D"], requires_grad=True)
dL = t(scenes.upstream_grad(W, H))
st = mod.GaussianRasterizationSettings(H, W, sc.tanfovx, sc.tanfovy, torch.zeros(3, device=dev), 1.0, t(sc.viewmatrix),
                                       t(sc.projmatrix), 3, t(sc.campos), False, False)
def step(timing=None):
    t0 = time.perf_counter()
    color, radii = mod.GaussianRasterizer(st)(means3D=params["means3D"], means2D=means2D, opacities=params["opacities"],
                                              shs=params["shs"], scales=params["scales"], rotations=params["rotations"])
    t1 = time.perf_counter()
    torch.autograd.backward(color, dL)
    t2 = time.perf_counter()
    for p in params.values():
        p.grad = None
    means2D.grad = None
    t3 = time.perf_counter()
    if timing is not None:
        timing.append((t1 - t0, t2 - t1, t3 - t2))
for _ in range(5): step()
torch.cuda.synchronize()
tm = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
w0 = time.perf_counter(); e0.record()
for _ in range(20): step(tm)
e1.record(); torch.cuda.synchronize(); w1 = time.perf_counter()
a = np.array(tm) * 1e3
print("gpu ms/step", e0.elapsed_time(e1) / 20, "wall ms/step", (w1 - w0) * 1e3 / 20)
print("cpu ms: fwd call %.3f  bwd call %.3f  zero %.3f" % tuple(a.mean(0)))
# raw C-ABI calls without autograd
with torch.no_grad():
    args = (st.bg, params["means3D"], torch.Tensor([]), params["opacities"], params["scales"], params["rotations"], 1.0,
            torch.Tensor([]), st.viewmatrix, st.projmatrix, st.tanfovx, st.tanfovy, H, W, params["shs"], 3, st.campos, False, False)
    for _ in range(3):
        R, color, radii, g, b, im = _C.rasterize_gaussians(*args)
        _C.rasterize_gaussians_backward(st.bg, params["means3D"], radii, torch.Tensor([]), params["scales"], params["rotations"], 1.0,
                                        torch.Tensor([]), st.viewmatrix, st.projmatrix, st.tanfovx, st.tanfovy, dL, params["shs"], 3,
                                        st.campos, g, R, b, im, False)
    torch.cuda.synchronize()
    e0.record(); w0 = time.perf_counter()
    for _ in range(20):
        R, color, radii, g, b, im = _C.rasterize_gaussians(*args)
        _C.rasterize_gaussians_backward(st.bg, params["means3D"], radii, torch.Tensor([]), params["scales"], params["rotations"], 1.0,
                                        torch.Tensor([]), st.viewmatrix, st.projmatrix, st.tanfovx, st.tanfovy, dL, params["shs"], 3,
                                        st.campos, g, R, b, im, False)
    e1.record(); torch.cuda.synchronize(); w1 = time.perf_counter()
    print("raw shim: gpu ms/step", e0.elapsed_time(e1) / 20, "wall", (w1 - w0) * 1e3 / 20)
