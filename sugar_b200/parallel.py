"""View-sharded multi-GPU step: one process per GPU, Gaussian parameters replicated, each rank
renders its own view(s), per-Gaussian gradients summed with ONE NCCL all-reduce over
NVLink 5 / NVSwitch (the reference is strictly 1 view / 1 GPU,
sugar_trainers/coarse_sdf.py:98,507, so this is new behaviour: loss = mean over the batch's views).

The path has no other exchange step, so there is no custom collective kernel: the flat arena
    [ points 3 | sh 3M | opacity 1 | scales 3 | quaternions 4 ]  x P   fp32
is reduced in place by `torch.distributed.all_reduce` (backend "nccl", or "gloo" in CPU tests).
Per-view densification statistics (|means2D.grad|, radii) must NOT be summed this way
(sugar_scene/sugar_densifier.py:156-164); they stay rank-local.
"""
from typing import Dict

import torch
import torch.distributed as dist

ARENA_FIELDS = ("means3D", "shs", "opacities", "scales", "rotations")


def shard_views(num_views: int, rank: int, world: int):
    """Views rendered by `rank`: {rank, rank+world, ...} (round-robin keeps ranks balanced)."""
    return list(range(rank, num_views, world))


class GradArena:
    """Flat per-Gaussian gradient buffer laid out for a single all-reduce."""

    def __init__(self, P: int, M: int, device, fields=ARENA_FIELDS):
        widths = {"means3D": 3, "shs": 3 * M, "opacities": 1, "scales": 3, "rotations": 4}
        self.fields = [f for f in fields]
        self.P = P
        self.offsets = {}
        off = 0
        for f in self.fields:
            self.offsets[f] = (off, widths[f] * P)
            off += widths[f] * P
        self.flat = torch.zeros(off, dtype=torch.float32, device=device)

    def view(self, name: str) -> torch.Tensor:
        o, n = self.offsets[name]
        return self.flat[o:o + n]

    def pack(self, params: Dict[str, torch.Tensor]) -> None:
        for f in self.fields:
            g = params[f].grad
            v = self.view(f)
            if g is None:
                v.zero_()
            else:
                v.copy_(g.reshape(-1))

    def all_reduce_from(self, params: Dict[str, torch.Tensor], scale: float = 1.0) -> torch.Tensor:
        """Pack the ranks' local gradients, sum them over the process group, scale (1/num_views)."""
        self.pack(params)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        if scale != 1.0:
            self.flat.mul_(scale)
        return self.flat

    def unpack_to(self, params: Dict[str, torch.Tensor]) -> None:
        """Write the reduced gradients back as .grad of the (replicated) parameters."""
        for f in self.fields:
            params[f].grad = self.view(f).view_as(params[f]).clone()
