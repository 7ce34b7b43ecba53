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
params = {k: t(getattr(sc, k)) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
dL = t(scenes.upstream_grad(W, H))
bg = torch.zeros(3, device=dev); vm = t(sc.viewmatrix); pm = t(sc.projmatrix); cp = t(sc.campos)
E = torch.Tensor([])
def fwd():
    return _C.rasterize_gaussians(bg, params["means3D"], E, params["opacities"], params["scales"], params["rotations"], 1.0,
            E, vm, pm, sc.tanfovx, sc.tanfovy, H, W, params["shs"], 3, cp, False, False)
def bwd(R, radii, g, b, im):
    return _C.rasterize_gaussians_backward(bg, params["means3D"], radii, E, params["scales"], params["rotations"], 1.0,
            E, vm, pm, sc.tanfovx, sc.tanfovy, dL, params["shs"], 3, cp, g, R, b, im, False)
for _ in range(3):
    R, color, radii, g, b, im = fwd(); bwd(R, radii, g, b, im)
torch.cuda.synchronize()
for label, n_f, n_b in (("fwd only", 1, 0), ("fwd+bwd", 1, 1)):
    tf = tb = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); w0 = time.perf_counter()
    for _ in range(20):
        a = time.perf_counter(); R, color, radii, g, b, im = fwd(); c = time.perf_counter(); tf += c - a
        if n_b:
            bwd(R, radii, g, b, im); tb += time.perf_counter() - c
    e1.record(); torch.cuda.synchronize(); w1 = time.perf_counter()
    print(label, "gpu ms/step %.3f wall %.3f  cpu fwd %.3f bwd %.3f" % (e0.elapsed_time(e1) / 20, (w1 - w0) * 50, tf * 50, tb * 50))
# backward only repeated on the same forward state
R, color, radii, g, b, im = fwd(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); w0 = time.perf_counter()
for _ in range(20): bwd(R, radii, g, b, im)
w1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
print("bwd only: gpu ms %.3f cpu ms %.3f" % (e0.elapsed_time(e1) / 20, (w1 - w0) * 50))
print(torch.cuda.memory_stats()["num_alloc_retries"], torch.cuda.memory_stats()["num_device_alloc"], torch.cuda.memory_stats()["num_device_free"])
# ---- where does the backward call block?  time the C call separately
from sugar_b200 import _lib
orig_b = _lib.lib.sgr_rasterize_backward
orig_f = _lib.lib.sgr_rasterize_forward
acc = {"b": 0.0, "f": 0.0}
def tb(*a):
    t0 = time.perf_counter(); r = orig_b(*a); acc["b"] += time.perf_counter() - t0; return r
def tf(*a):
    t0 = time.perf_counter(); r = orig_f(*a); acc["f"] += time.perf_counter() - t0; return r
_C.lib = type("L", (), {})()
for name in dir(_lib.lib):
    pass
class Wrap:
    def __getattr__(self, k):
        if k == "sgr_rasterize_backward": return tb
        if k == "sgr_rasterize_forward": return tf
        return getattr(_lib.lib, k)
_C.lib = Wrap()
torch.cuda.synchronize()
tF = tB = 0.0
for _ in range(20):
    a = time.perf_counter(); R, color, radii, g, b, im = fwd(); c = time.perf_counter(); tF += c - a
    bwd(R, radii, g, b, im); tB += time.perf_counter() - c
torch.cuda.synchronize()
print("python fwd %.3f (C %.3f)  python bwd %.3f (C %.3f)" % (tF * 50, acc["f"] * 50, tB * 50, acc["b"] * 50))
