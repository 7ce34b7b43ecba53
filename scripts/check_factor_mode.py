"""torchrun --nproc-per-node N scripts/check_factor_mode.py : the view-parallel step's reduced gradients
with SH factor mode (all-gather factors + rebuild) against the plain all-reduce of the full arena."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sugar_b200 import diff_gaussian_rasterization as mod  # noqa: E402
from sugar_b200 import parallel, scenes  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    P, W, H, D = 200_000, 640, 360, 3
    base = scenes.make_scene(P, W, H, seed=3, camera="posed")
    sc = scenes.with_camera_offset(base, 0.1 * rank, (0.2 * rank, -0.1 * rank, 0.3 * rank))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    dL = t(scenes.upstream_grad(W, H, seed=1 + rank) / world)
    results = {}
    for mode in ("plain", "factors"):
        params = {k: t(getattr(base, k)).requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
        means2D = torch.zeros_like(params["means3D"], requires_grad=True)
        st = mod.GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=sc.tanfovx, tanfovy=sc.tanfovy,
                                               bg=torch.zeros(3, device=dev), scale_modifier=1.0, viewmatrix=t(sc.viewmatrix),
                                               projmatrix=t(sc.projmatrix), sh_degree=D, campos=t(sc.campos),
                                               prefiltered=False, debug=False)
        arena = parallel.GradArena(P, 16, dev)
        with parallel.sh_factor_mode(mode == "factors"):
            color, _ = mod.GaussianRasterizer(st)(means3D=params["means3D"], means2D=means2D,
                                                  opacities=params["opacities"], shs=params["shs"],
                                                  scales=params["scales"], rotations=params["rotations"])
            torch.autograd.backward(color, dL)
            arena.all_reduce_from(params, campos=t(sc.campos), sh_degree=D)
        results[mode] = {k: v.grad.detach().clone() for k, v in params.items()}
    worst = 0.0
    for k in results["plain"]:
        a, b = results["factors"][k].double(), results["plain"][k].double()
        e = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
        worst = max(worst, e)
        if rank == 0:
            print(f"{k:10s} rel err factors vs plain all-reduce: {e:.3e}  (|g|max {float(b.abs().max()):.3e})")
    ok = worst <= 1e-4
    if rank == 0:
        print("FACTOR MODE", "OK" if ok else "MISMATCH", f"world={world}")
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
