"""View-sharded multi-GPU step: one process per GPU, Gaussian parameters replicated, each rank
renders its own view(s), per-Gaussian gradients summed with ONE NCCL all-reduce over
NVLink 5 / NVSwitch (the reference is strictly 1 view / 1 GPU,
sugar_trainers/coarse_sdf.py:98,507, so this is new behaviour: loss = mean over the batch's views).

The path has no other exchange step.  The flat arena
    [ points 3 | opacity 1 | scales 3 | quaternions 4 | sh 3M ]  x P   fp32
is reduced in place by `torch.distributed.all_reduce` (backend "nccl", or "gloo" in CPU tests).

SH factor mode (`sh_factor_mode()`): 3M of the 11+3M floats per Gaussian are dL_dsh, and each view's
dL_dsh is an outer product  basis(dir_view)[M] x dL/dRGB[3]  (backward.cu:20-139).  Instead of
all-reducing 12M bytes per Gaussian the ranks all-gather the 12-byte dL/dRGB factors (plus their
camera positions) and every rank rebuilds the summed dL_dsh with one kernel
(sgr_sh_grad_from_factors); only the other 11 floats go through the all-reduce.  At M=16 that is
5.4x less data on the wire at 2 GPUs and 2.6x less at 8, and each rank's backward skips writing
its 192 B/Gaussian of dL_dsh.
Per-view densification statistics (|means2D.grad|, radii) must NOT be summed this way
(sugar_scene/sugar_densifier.py:156-164); they stay rank-local.
"""
import contextlib
from typing import Dict

import torch
import torch.distributed as dist

ARENA_FIELDS = ("means3D", "opacities", "scales", "rotations", "shs")  # = layout of the backward's flat buffer


def shard_views(num_views: int, rank: int, world: int):
    """Views rendered by `rank`: {rank, rank+world, ...} (round-robin keeps ranks balanced)."""
    return list(range(rank, num_views, world))


@contextlib.contextmanager
def sh_factor_mode(enabled: bool = True):
    """Within this context sugar_b200's rasterizer backward emits SH factors instead of dL_dsh; the
    gradients must then go through `GradArena.all_reduce_from(..., campos=, sh_degree=)`."""
    from . import _C
    old = (_C.SH_FACTOR_MODE, _C.FACTOR_HOOK)
    set_sh_factor_mode(enabled)
    try:
        yield
    finally:
        _C.SH_FACTOR_MODE, _C.FACTOR_HOOK = old
        _early_gather.clear()


_early_gather = {}  # data_ptr of the factor tensor -> (gathered [world,P,3], work handle)


def _start_gather(dRGB: torch.Tensor) -> None:
    """Stage hook of the backward: the factors are final on the current stream, the per-Gaussian
    backward is not enqueued yet -> the all-gather runs on NCCL's stream underneath it."""
    world = dist.get_world_size()
    d_all = torch.empty((world,) + tuple(dRGB.shape), dtype=dRGB.dtype, device=dRGB.device)
    _early_gather.clear()
    _early_gather[dRGB.data_ptr()] = (d_all, dist.all_gather_into_tensor(d_all, dRGB, async_op=True))


def set_sh_factor_mode(enabled: bool = True) -> None:
    """Process-wide switch behind `sh_factor_mode()` (for loops that do not want a context manager)."""
    from . import _C
    _C.SH_FACTOR_MODE = bool(enabled)
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    _C.FACTOR_HOOK = _start_gather if (enabled and multi) else None


def gather_factors(dRGB: torch.Tensor, campos: torch.Tensor):
    """All-gather the per-view SH factors: returns (dRGB_all [V,P,3], campos_all [V,3]), V = world size."""
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    if world == 1:
        return dRGB.reshape(1, -1, 3), campos.reshape(1, 3)
    d_all = torch.empty((world,) + tuple(dRGB.shape), dtype=dRGB.dtype, device=dRGB.device)
    c_all = torch.empty((world, 3), dtype=campos.dtype, device=campos.device)
    dist.all_gather_into_tensor(d_all, dRGB.contiguous())
    dist.all_gather_into_tensor(c_all, campos.reshape(3).contiguous())
    return d_all.view(world, -1, 3), c_all


def sh_grad_from_factors(means3D: torch.Tensor, campos_all: torch.Tensor, dRGB_all: torch.Tensor, M: int,
                         sh_degree: int, out: torch.Tensor = None) -> torch.Tensor:
    """dL_dsh [P,M,3] = sum_v basis(normalize(means3D - campos_all[v])) (x) dRGB_all[v]  (CUDA kernel)."""
    from ._lib import check, lib
    P, V = means3D.shape[0], campos_all.shape[0]
    if out is None:
        out = torch.empty((P, M, 3), dtype=torch.float32, device=means3D.device)
    assert out.is_contiguous() and out.numel() == P * M * 3 and dRGB_all.is_contiguous() and campos_all.is_contiguous()
    with torch.cuda.device(means3D.device):
        check(lib.sgr_sh_grad_from_factors(P, M, sh_degree, V, means3D.contiguous().data_ptr(), campos_all.data_ptr(),
                                           dRGB_all.data_ptr(), out.data_ptr(),
                                           torch.cuda.current_stream(means3D.device).cuda_stream))
    return out


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
        self.numel = off
        self.device = device
        self._flat = None

    @property
    def flat(self) -> torch.Tensor:
        """Staging buffer of the packing fallback; allocated on first use (the in-place path never needs it)."""
        if self._flat is None:
            self._flat = torch.zeros(self.numel, dtype=torch.float32, device=self.device)
        return self._flat

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
        """If the gradients already sit back to back in one storage in arena order (they do when they
        come from sugar_b200's backward: autograd hands the leaves detached aliases of the backward's
        flat buffer), return that run as one flat tensor so the reduction needs no packing."""
        g0 = params[self.fields[0]].grad
        if g0 is None or g0.dtype != torch.float32:
            return None
        store, s0 = g0.untyped_storage(), g0.storage_offset()
        for f in self.fields:
            g = params[f].grad
            o, n = self.offsets[f]
            if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.numel() != n \
                    or g.untyped_storage().data_ptr() != store.data_ptr() or g.storage_offset() != s0 + o:
                return None
        whole = torch.empty(0, dtype=torch.float32, device=g0.device).set_(store)
        self._base = whole[s0:]
        return self._base[:self.numel]

    def _all_reduce_factored(self, params, buf, campos, sh_degree):
        """SH factor mode: all-reduce everything but the sh slot, all-gather the factors, rebuild dL_dsh.
        The collectives are issued asynchronously (factors first) so that the rebuild kernel overlaps
        the all-reduce of the other fields."""
        P = self.P
        o_sh, n_sh = self.offsets["shs"]
        M = n_sh // (3 * P)
        # dL_dcolors sits behind [arena | dL_dmeans2D 3P] in the backward's buffer (sugar_b200/_C.py)
        o_col = self.numel + 3 * P
        dRGB = self._base[o_col:o_col + 3 * P]
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        pending = []
        if world > 1:
            early = _early_gather.pop(dRGB.data_ptr(), None)  # started by the backward's stage hook?
            if early is not None:
                d_all, hg = early[0].view(world, P, 3), [early[1]]
            else:
                d_all = torch.empty((world, P, 3), dtype=torch.float32, device=buf.device)
                hg = [dist.all_gather_into_tensor(d_all, dRGB, async_op=True)]
            if campos.dim() == 1:  # this rank's camera only: gather the others'
                c_all = torch.empty((world, 3), dtype=torch.float32, device=buf.device)
                hg.append(dist.all_gather_into_tensor(c_all, campos.reshape(3).contiguous(), async_op=True))
            else:                  # [world,3]: the caller already knows every rank's camera
                c_all = campos.contiguous()
            # fields before / after the sh slot are contiguous runs of the arena
            if o_sh > 0:
                pending.append(dist.all_reduce(buf[:o_sh], op=dist.ReduceOp.SUM, async_op=True))
            if o_sh + n_sh < buf.numel():
                pending.append(dist.all_reduce(buf[o_sh + n_sh:], op=dist.ReduceOp.SUM, async_op=True))
            for h in hg:
                h.wait()
        else:
            d_all, c_all = dRGB.view(1, P, 3), campos.reshape(-1, 3)[:1].contiguous()
        sh_grad_from_factors(params["means3D"].detach(), c_all, d_all, M, sh_degree, out=buf[o_sh:o_sh + n_sh])
        for h in pending:
            h.wait()

    def all_reduce_from(self, params: Dict[str, torch.Tensor], scale: float = 1.0, campos: torch.Tensor = None,
                        sh_degree: int = None) -> torch.Tensor:
        """Sum the ranks' local gradients over the process group (in place when possible) and scale
        by 1/num_views.  Returns the reduced flat arena.  Under `sh_factor_mode()` pass this rank's
        camera position ([3]; or all ranks' positions [world,3], saving a tiny all-gather) and the active
        SH degree."""
        from . import _C
        buf = self._shared_base(params)
        if _C.SH_FACTOR_MODE and "shs" in self.offsets and self.offsets["shs"][1] > 0:
            if buf is None or campos is None or sh_degree is None:
                raise RuntimeError("sh_factor_mode needs the gradients of ONE sugar_b200 backward per step "
                                   "(no accumulation) plus campos= and sh_degree=")
            self._all_reduce_factored(params, buf, campos, sh_degree)
            if scale != 1.0:
                buf.mul_(scale)
            self.reduced = buf
            return buf
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
        src = self.reduced if hasattr(self, "reduced") else self.flat
        for f in self.fields:
            o, n = self.offsets[f]
            if params[f].grad is None or params[f].grad.data_ptr() != src.data_ptr() + o * 4:
                params[f].grad = src[o:o + n].view_as(params[f]).clone()
