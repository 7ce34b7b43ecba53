"""Parity + timing at the larger BASELINE configs: python scripts/big_parity.py P W H [mesh]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import helpers as h
from sugar_b200 import diff_gaussian_rasterization as ours, scenes
P, W, H = (int(a) for a in sys.argv[1:4])
mesh = len(sys.argv) > 4
sc = scenes.make_scene(P, W, H, seed=0, mesh_bound=mesh)
dL = scenes.upstream_grad(W, H)
ref = h.load_ref_module()
res = {}
for name, mod in (("ours", ours), ("ref", ref)):
    out = h.run_module(mod, sc, (0, 0, 0), dL, use_sh=True, sh_degree=3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        out = h.run_module(mod, sc, (0, 0, 0), dL, use_sh=True, sh_degree=3)
    torch.cuda.synchronize()
    res[name] = out
    print(name, "R", out["num_rendered"], "visible", int((out["radii"] > 0).sum()))
a, b = res["ours"], res["ref"]
print("num_rendered equal", a["num_rendered"] == b["num_rendered"], "radii equal", bool(torch.equal(a["radii"], b["radii"])))
print("image bit-exact", bool(torch.equal(a["color"].view(torch.int32), b["color"].view(torch.int32))))
for k in sorted(b["grads"]):
    print(f"  grad {k:10s} rel err {h.rel_err(a['grads'][k].cpu().numpy(), b['grads'][k].cpu().numpy()):.2e}")
