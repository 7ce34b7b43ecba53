"""Per-kernel timeline of one view-parallel step on rank 0 (the library's own CUDA-event log), under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        scripts/timeline_peer.py [--peer]
Prints the un-profiled step time first, then every launch of the last profiled step: stream-agnostic begin / duration."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
from sugar_b200 import diff_gaussian_rasterization as mod, scenes, _lib, parallel
P, W, H = 3_000_000, 1920, 1080
sc = scenes.make_scene(P, W, H, seed=0)
sc = scenes.with_camera_offset(sc, 0.05 * rank, (0.0, 0.0, 0.0))
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
params = {k: t(getattr(sc, k)).requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
means2D = torch.zeros_like(params["means3D"], requires_grad=True)
dL = t(scenes.upstream_grad(W, H, seed=1 + rank))
st = mod.GaussianRasterizationSettings(H, W, sc.tanfovx, sc.tanfovy, torch.zeros(3, device=dev), 1.0, t(sc.viewmatrix),
                                       t(sc.projmatrix), 3, t(sc.campos), False, False)
vp = parallel.ViewParallel(chunks=int(os.environ.get("CHUNKS", "4")), peer=True if "--peer" in sys.argv else False,
                           taper="--no-taper" not in sys.argv)
def step():
    with vp.context():
        color, radii = mod.GaussianRasterizer(st)(means3D=params["means3D"], means2D=means2D, opacities=params["opacities"],
                                                  shs=params["shs"], scales=params["scales"], rotations=params["rotations"])
        torch.autograd.backward(color, dL)
    for p in params.values():
        p.grad = None
    means2D.grad = None
for _ in range(10): step()
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30): step()
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / 30], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
_lib.profile(True)
for _ in range(4): step()
torch.cuda.synchronize(); _lib.profile_read()
for _ in range(2): step()
torch.cuda.synchronize()
tl = _lib.profile_timeline()
if rank == 0:
    print("config", {k: v for k, v in os.environ.items() if k.startswith("SGR_") or k == "CHUNKS"}, sys.argv[1:])
    print("step_ms_unprofiled_max_over_ranks %.3f" % float(ms.item()))
    half = len(tl) // 2
    t0 = tl[half][1]
    for name, b, e in tl[half:]:
        print(f"{name:20s} begin {b - t0:8.3f}  dur {e - b:7.3f}  end {e - t0:8.3f}")
dist.barrier()
dist.destroy_process_group()
