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

ARENA_FIELDS = ("means3D", "opacities", "shs", "scales", "rotations")  # = layout of the backward's flat buffer


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

    def _shared_base(self, params: Dict[str, torch.Tensor]):
        """If the gradients already sit back to back in one buffer in arena order (they do when they
        come from sugar_b200's backward), return that slice so the reduction needs no packing."""
        g0 = params[self.fields[0]].grad
        base = getattr(g0, "_base", None) if g0 is not None else None
        if base is None or base.dim() != 1 or base.dtype != torch.float32:
            return None
        esz, start = base.element_size(), base.data_ptr()
        for f in self.fields:
            g = params[f].grad
            o, n = self.offsets[f]
            if g is None or getattr(g, "_base", None) is not base or not g.is_contiguous() or g.numel() != n \
                    or g.data_ptr() != start + o * esz:
                return None
        return base[:self.flat.numel()]

    def all_reduce_from(self, params: Dict[str, torch.Tensor], scale: float = 1.0) -> torch.Tensor:
        """Sum the ranks' local gradients over the process group (in place when possible) and scale
        by 1/num_views.  Returns the reduced flat arena."""
        buf = self._shared_base(params)
        if buf is None:
            self.pack(params)
            buf = self.flat
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        if scale != 1.0:
            buf.mul_(scale)
        self.reduced = buf
        return buf

    def unpack_to(self, params: Dict[str, torch.Tensor]) -> None:
        """Write the reduced gradients back as .grad of the (replicated) parameters."""
        src = getattr(self, "reduced", self.flat)
        for f in self.fields:
            o, n = self.offsets[f]
            if params[f].grad is None or params[f].grad.data_ptr() != src.data_ptr() + o * 4:
                params[f].grad = src[o:o + n].view_as(params[f]).clone()
