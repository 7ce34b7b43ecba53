"""Timeline of the view-parallel step (host calls + kernels + NCCL) from torch.profiler (CUPTI), to see where the GPU idles:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
        scripts/trace_step.py [--chunks 4] [--out gpurun_out/trace]
Rank 0 writes <out>_n<world>.json (chrome trace of three steps) and prints, for the middle step, every GPU interval with the
idle gap in front of it and the host time of the call that launched it.  Numbers taken under the profiler are not bench values."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--chunks", type=int, default=4)
ap.add_argument("--gaussians", type=int, default=3_000_000)
ap.add_argument("--out", default="gpurun_out/trace")
ap.add_argument("--side-stream", action="store_true")
ap.add_argument("--force-exchange", action="store_true", help="one rank: run the exchange path anyway (no NCCL)")
args = ap.parse_args()

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
import bench
from sugar_b200 import diff_gaussian_rasterization as mod, parallel

scenes = bench.load_scenes()
P, W, H, D = args.gaussians, 1920, 1080, 3
sc = scenes.make_scene(P, W, H, seed=0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
params = {k: t(getattr(sc, k)).requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
means2D = torch.zeros_like(params["means3D"], requires_grad=True)
Pm = scenes.projection_matrix(0.01, 100.0, sc.tanfovx, sc.tanfovy)
vm = t(np.eye(4, dtype=np.float32)); pm = t(Pm.T.astype(np.float32))
cp = torch.zeros(3, device=dev); bg = torch.zeros(3, device=dev)
g = t(scenes.upstream_grad(W, H, seed=1 + rank) / max(world, 1))
settings = mod.GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=sc.tanfovx, tanfovy=sc.tanfovy, bg=bg,
                                             scale_modifier=1.0, viewmatrix=vm, projmatrix=pm, sh_degree=D, campos=cp,
                                             prefiltered=False, debug=False)
vp = parallel.ViewParallel(chunks=args.chunks, force=args.force_exchange, side_stream=args.side_stream) if (world > 1 or args.force_exchange) else None


def step():
    rast = mod.GaussianRasterizer(settings)
    color, radii = rast(means3D=params["means3D"], means2D=means2D, opacities=params["opacities"], shs=params["shs"],
                        scales=params["scales"], rotations=params["rotations"])
    torch.autograd.backward(color, g)
    for p in params.values():
        p.grad = None
    means2D.grad = None


import contextlib
with (vp.context() if vp is not None else contextlib.nullcontext()):
    for _ in range(10):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            with torch.profiler.record_function("sgr_step"):
                step()
        torch.cuda.synchronize()
if rank == 0:
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    path = f"{args.out}_n{world}.json"
    prof.export_chrome_trace(path)
    ev = json.load(open(path))["traceEvents"]
    gpu = sorted((e for e in ev if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")),
                 key=lambda e: e["ts"])
    launch = {e["args"]["correlation"]: e for e in ev
              if e.get("ph") == "X" and e.get("cat") in ("cuda_runtime", "cuda_driver") and "correlation" in e.get("args", {})}
    # the middle step on the GPU = from the second preprocess kernel to the third
    pre = [k for k in gpu if "preprocess_kernel" in k["name"]]
    lo, hi = pre[1]["ts"], pre[2]["ts"]
    print(f"GPU step (preprocess to preprocess): {(hi - lo) / 1e3:.3f} ms")
    prev_end = None
    for k in gpu:
        if not lo <= k["ts"] < hi:
            continue
        lc = launch.get(k["args"].get("correlation"))
        gap = (k["ts"] - prev_end) if prev_end is not None else 0.0
        print(f"{(k['ts'] - lo) / 1e3:8.3f}  dur {k['dur'] / 1e3:7.3f}  gap {gap / 1e3:7.3f}  s{k['args'].get('stream')}"
              f"  launched {((lc['ts'] - lo) / 1e3) if lc else float('nan'):8.3f}  {k['name'][:64]}")
        prev_end = max(prev_end or 0, k["ts"] + k["dur"])
    print("host calls > 25 us around the step (same origin):")
    for e in sorted((e for e in ev if e.get("ph") == "X" and e.get("cat") in ("cuda_runtime", "cuda_driver", "cpu_op",
                                                                              "user_annotation")
                     and lo - 1500 <= e["ts"] < hi and e.get("dur", 0) > 25), key=lambda e: e["ts"]):
        print(f"{(e['ts'] - lo) / 1e3:8.3f}  dur {e['dur'] / 1e3:7.3f}  tid {e.get('tid')}  {e['cat']:14s} {e['name'][:70]}")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
