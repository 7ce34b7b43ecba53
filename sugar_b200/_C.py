"""`_C`-compatible shim: the three symbols the reference's pybind module exports
(diff-gaussian-rasterization/ext.cpp:15-19, rasterize_points.h:19-67), same positional
signatures and return tuples, implemented over the C ABI of libsugar_b200.so.

torch is used only for device memory (caching allocator), the current stream and the device
guard; all arithmetic happens in the hand-written sm_100a kernels.
"""
import contextlib
import ctypes as C
import os
import threading
from typing import Tuple

import torch

from . import _lib
from ._lib import SgrGaussians, SgrView, check, lib

_USE_HINT = os.environ.get("SGR_NO_CAPACITY_HINT", "0") != "1"


class Context:
    """Mutable state of the op that outlives one call.  Nothing here is process-global: a model (or a
    thread, or a view-parallel exchange) can own its context and select it with `use_context`;
    calls made outside any `use_context` block share `default_context()`.

    capacity_hint  (device, H, W) -> instance capacity.  After the first forward the binning buffer is
                   sized from the previous view's instance count (x1.25 + slack) so the whole forward is
                   enqueued without waiting for the device; the C side re-runs binning on overflow
                   (include/sugar_b200.h).
    exchange       None, or the view-parallel exchange (sugar_b200.parallel.ViewParallel) that the
                   backward hands its per-Gaussian gradients to before returning them to autograd.

    The forward stores the context it ran under in the autograd node, so the backward -- which autograd
    runs on its own engine thread -- uses the same one."""

    def __init__(self):
        self.capacity_hint = {}
        self.exchange = None
        self.lock = threading.Lock()


_default_context = Context()
_tls = threading.local()


def default_context() -> Context:
    return _default_context


def current_context() -> Context:
    return getattr(_tls, "ctx", None) or _default_context


@contextlib.contextmanager
def use_context(ctx: Context):
    """Calls made by this thread inside the block use `ctx` (nestable)."""
    prev = getattr(_tls, "ctx", None)
    _tls.ctx = ctx
    try:
        yield ctx
    finally:
        _tls.ctx = prev


def _ptr(t: torch.Tensor, name: str):
    """Device pointer of an optional tensor: empty tensor == absent == NULL (forward.cu:205,241)."""
    if t is None or t.numel() == 0:
        return None
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if not t.is_cuda:
        raise _lib.SgrError(f"{name} must be a CUDA tensor (sugar_b200 has no CPU path)")
    return t.data_ptr()


_ROUND = 64 << 20  # arena sizes are rounded up to 64 MiB so the caching allocator sees few distinct sizes

# Optional workspace cache (SGR_WORKSPACE_CACHE=1).  The per-call buffers of this op are large
# (0.5-1 GB at 3M Gaussians); a cudaMalloc inside the training loop is a 10-160 ms device-wide
# stall.  By default torch's caching allocator is enough: each call makes one rounded allocation
# for the forward state and one for the gradients, and `_Arena.release()` breaks the callback
# reference cycle so they are returned by reference counting, not at the next cyclic GC (which is
# what used to fragment the allocator; scripts/step_times.py shows the difference).  With the
# cache on, blocks are additionally pinned here and handed out again as soon as nothing references
# their storage (stream-ordered reuse, like the caching allocator's).
_ws_cache = {}
_WS_MAX_PER_KEY = 6
_USE_WS_CACHE = os.environ.get("SGR_WORKSPACE_CACHE", "0") == "1"


def _use_count(t: torch.Tensor) -> int:
    return torch._C._storage_Use_Count(t.untyped_storage()._cdata)


def _big_empty(numel: int, dtype, device) -> torch.Tensor:
    """1-D tensor of `numel` elements for a per-call arena, recycled once all its views are dead."""
    if not _USE_WS_CACHE:
        return torch.empty(numel, dtype=dtype, device=device)
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, dtype, numel)
    entries = _ws_cache.setdefault(key, [])
    for blk, baseline in entries:
        if _use_count(blk) == baseline:
            return blk
    blk = torch.empty(numel, dtype=dtype, device=device)
    if len(entries) < _WS_MAX_PER_KEY:
        entries.append((blk, _use_count(blk)))
    return blk


def clear_workspace_cache() -> None:
    _ws_cache.clear()


def _round_up(n: int, a: int) -> int:
    return (n + a - 1) // a * a


class _Arena:
    """Allocator-callback state.  One torch allocation backs all three opaque buffers of a forward
    (carved at 256-byte offsets): per-step allocations of stable, rounded sizes keep the caching
    allocator from fragmenting (and from calling cudaMalloc, a device-wide stall) in steady state."""

    def __init__(self, device, reserve_bytes: int):
        self.device = device
        self.block = _big_empty(_round_up(reserve_bytes, _ROUND), torch.uint8, device) if reserve_bytes else None
        self.used = 0
        self.parts = {}
        self.error = None   # an exception raised inside a callback (e.g. torch OOM): re-raised after the C call returns
        self.cbs = {name: _lib.ALLOC_FN(lambda _ctx, n, name=name: self._alloc(name, n)) for name in ("geom", "binning", "img")}

    def _alloc(self, name, nbytes):
        try:
            nbytes = int(nbytes)
            if self.block is not None and self.used + nbytes + 256 <= self.block.numel():
                off = _round_up(self.block.data_ptr() + self.used, 256) - self.block.data_ptr()
                t = self.block[off:off + nbytes]
                self.used = off + nbytes
            else:  # no (or too small a) reservation: plain allocation, still rounded
                t = _big_empty(_round_up(nbytes, _ROUND), torch.uint8, self.device)[:nbytes]
            self.parts[name] = t
            return t.data_ptr()
        except BaseException as e:  # never unwind through the C frame: NULL makes the library stop with an error status
            self.error = e
            return None

    def get(self, name):
        return self.parts.get(name, torch.empty(0, dtype=torch.uint8, device=self.device))

    def release(self):
        """Break the arena <-> callback reference cycle so the block is released by reference
        counting as soon as the autograd graph lets go of the views (not at the next cyclic GC)."""
        self.cbs = None
        self.block = None
        self.parts = {}


def _view_struct(bg, scale_modifier, viewmatrix, projmatrix, tan_fovx, tan_fovy, H, W, degree, campos, prefiltered,
                 debug) -> SgrView:
    v = SgrView()
    v.image_height, v.image_width = int(H), int(W)
    v.tanfovx, v.tanfovy = float(tan_fovx), float(tan_fovy)
    v.bg = _ptr(bg, "bg")
    v.scale_modifier = float(scale_modifier)
    v.viewmatrix = _ptr(viewmatrix, "viewmatrix")
    v.projmatrix = _ptr(projmatrix, "projmatrix")
    v.sh_degree = int(degree)
    v.campos = _ptr(campos, "campos")
    v.prefiltered, v.debug = int(bool(prefiltered)), int(bool(debug))
    return v


def _gauss_struct(P, M, means3D, opacity, sh, colors, scales, rotations, cov3D_precomp, sh_rest=None) -> SgrGaussians:
    g = SgrGaussians()
    g.activations = 1 if sh_rest is not None else 0   # SGR_ACT_RAW: raw SuGaR parameters, (dc, rest) SH arrays
    g.sh_rest = _ptr(sh_rest, "sh_rest") if sh_rest is not None else None
    g.P, g.M = int(P), int(M)
    g.means3D = _ptr(means3D, "means3D")
    g.opacities = _ptr(opacity, "opacity")
    g.shs = _ptr(sh, "sh")
    g.colors_precomp = _ptr(colors, "colors")
    g.scales = _ptr(scales, "scales")
    g.rotations = _ptr(rotations, "rotations")
    g.cov3D_precomp = _ptr(cov3D_precomp, "cov3D_precomp")
    return g


def _c(t):
    return t if (t is None or t.is_contiguous()) else t.contiguous()


def rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                        viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos,
                        prefiltered, debug, *, context: Context = None, sh_rest: torch.Tensor = None
                        ) -> Tuple[int, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """RasterizeGaussiansCUDA (rasterize_points.cu:36-115).  Keyword-only extras that are not part of the
    reference signature: `context` selects whose capacity hint is used (default: the calling thread's current
    one); `sh_rest` selects raw-parameter mode (sugar_b200/fused.py): `sh` is then the model's DC array [P,1,3],
    `sh_rest` [P,M-1,3], and opacity / scales / rotations are the RAW parameters, activated in the kernel."""
    context = context or current_context()
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if not means3D.is_cuda:
        raise _lib.SgrError("sugar_b200 needs CUDA tensors: there is no CPU fallback")
    dev = means3D.device
    P, H, W = means3D.size(0), int(image_height), int(image_width)
    M = sh.size(1) if (sh is not None and sh.numel() != 0) else 0
    if sh_rest is not None:
        if M != 1 or sh_rest.dim() != 3 or sh_rest.size(0) != P:
            raise RuntimeError("raw-parameter mode: sh must be [P,1,3] and sh_rest [P,M-1,3]")
        M = 1 + sh_rest.size(1)
        sh_rest = _c(sh_rest)
    with torch.cuda.device(dev):
        out_color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        key = (dev.index, H, W)
        hint = context.capacity_hint.get(key, 0) if _USE_HINT else 0
        reserve = lib.sgr_geometry_bytes(P) + lib.sgr_image_bytes(W, H) + 1024
        if hint:
            reserve += lib.sgr_binning_bytes(hint)
        arena = _Arena(dev, reserve if P else 0)
        means3D, colors, opacity, scales, rotations, cov3D_precomp, sh = map(
            _c, (means3D, colors, opacity, scales, rotations, cov3D_precomp, sh))
        background, viewmatrix, projmatrix, campos = map(_c, (background, viewmatrix, projmatrix, campos))
        view = _view_struct(background, scale_modifier, viewmatrix, projmatrix, tan_fovx, tan_fovy, H, W, degree,
                            campos, prefiltered, debug)
        g = _gauss_struct(P, M, means3D, opacity, sh, colors, scales, rotations, cov3D_precomp, sh_rest)
        rendered = C.c_int64(0)
        stream = torch.cuda.current_stream(dev).cuda_stream
        status = lib.sgr_rasterize_forward(C.byref(view), C.byref(g), arena.cbs["geom"], None, arena.cbs["binning"], None,
                                           arena.cbs["img"], None, out_color.data_ptr(), radii.data_ptr() if P else None,
                                           hint, C.byref(rendered), stream)
        if arena.error is not None:   # the allocator callback failed (e.g. out of memory): its own exception, not
            err = arena.error         # the library's generic "allocator returned NULL"
            arena.release()
            raise err
        check(status)
        R = int(rendered.value)
        if P:
            # next view's optimistic capacity: 25% headroom, rounded to 1M instances (stable sizes)
            context.capacity_hint[key] = _round_up(int(R * 1.25) + 65536, 1 << 20)
    geom_t, binning_t, img_t = arena.get("geom"), arena.get("binning"), arena.get("img")
    arena.release()
    return R, out_color, radii, geom_t, binning_t, img_t


_PART_ALIGN = 16  # floats: every gradient array starts on a 64-byte boundary (TMA bulk stores need 16)


def _alloc_backward(P: int, M: int, dev, with_records: bool, split_sh: bool = False):
    """All eight gradients (+ the 44-byte reduce records of the view-parallel step) + the accumulator scratch in
    ONE rounded allocation: a stable size for the caching allocator, one free when autograd lets go."""
    widths = [("means3D", 3), ("opacity", 1), ("scales", 3), ("rotations", 4), ("sh", 3 if split_sh else 3 * M),
              ("means2D", 3), ("colors", 3), ("cov3D", 6)]
    if with_records:
        widths.append(("records", 11))
    if split_sh:
        widths.append(("sh_rest", 3 * (M - 1)))
    n_scratch = (lib.sgr_backward_scratch_bytes(P) + 3) // 4 if P else 0
    offs, o = {}, 0
    for name, w in widths:
        offs[name] = o
        # dL_dcolors is followed by 4 spare floats: the view-parallel exchange appends the camera position there,
        # so that one all-gather carries a view's SH factors and its camera
        o = _round_up(o + P * w + (4 if name == "colors" else 0), _PART_ALIGN)
    total = o + n_scratch + 64
    flat = _big_empty(_round_up(4 * total, _ROUND) // 4, torch.float32, dev)
    skew = (-(flat.data_ptr() // 4)) % _PART_ALIGN  # the allocator aligns to 512 B; be explicit anyway
    shapes = {"means3D": (P, 3), "opacity": (P, 1), "scales": (P, 3), "rotations": (P, 4),
              "sh": (P, 1 if split_sh else M, 3), "sh_rest": (P, max(M - 1, 0), 3), "means2D": (P, 3), "colors": (P, 3),
              "cov3D": (P, 6), "records": (P, 11)}
    bufs = {name: flat[skew + offs[name]:skew + offs[name] + P * w].view(shapes[name]) for name, w in widths}
    bufs["scratch"] = flat[skew + o:skew + o + n_scratch]
    bufs["colors_block"] = flat[skew + offs["colors"]:skew + offs["colors"] + 3 * P + 4]
    return bufs


def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, scale_modifier,
                                 cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, sh,
                                 degree, campos, geomBuffer, R, binningBuffer, imageBuffer, debug, *,
                                 context: Context = None, sh_rest: torch.Tensor = None):
    """RasterizeGaussiansBackwardCUDA (rasterize_points.cu:118-196).  With a view-parallel exchange attached to
    `context` (sugar_b200/parallel.py) the gradients come back already summed over the ranks.  With `sh_rest`
    (raw-parameter mode, see rasterize_gaussians) a ninth tensor dL_dsh_rest is returned and dL_dopacity /
    dL_dscales / dL_drotations / dL_dsh are gradients of the raw parameters."""
    context = context or current_context()
    dev = means3D.device
    P, H, W = means3D.size(0), dL_dout_color.size(1), dL_dout_color.size(2)
    M = sh.size(1) if (sh is not None and sh.numel() != 0) else 0
    raw = sh_rest is not None
    if raw:
        M = 1 + sh_rest.size(1)
        sh_rest = _c(sh_rest)
    ex = context.exchange if (context.exchange is not None and context.exchange.enabled()) else None
    if raw and ex is not None:
        raise NotImplementedError("raw-parameter mode under a view-parallel exchange: activate in PyTorch instead "
                                  "(the exchange acts on the op's output gradients either way)")
    with torch.cuda.device(dev):
        has_cov = cov3D_precomp is not None and cov3D_precomp.numel() != 0
        # the peer-memory exchange keeps its records in its own peer-visible buffer
        b = _alloc_backward(P, M, dev, with_records=ex is not None and not ex.uses_peer_memory(M, has_cov), split_sh=raw)
        if P != 0:
            means3D, colors, scales, rotations, cov3D_precomp, sh, dL_dout_color = map(
                _c, (means3D, colors, scales, rotations, cov3D_precomp, sh, dL_dout_color))
            background, viewmatrix, projmatrix, campos = map(_c, (background, viewmatrix, projmatrix, campos))
            view = _view_struct(background, scale_modifier, viewmatrix, projmatrix, tan_fovx, tan_fovy, H, W, degree,
                                campos, False, debug)
            # opacities are not an input of the reference's backward; the forward stored them
            g = _gauss_struct(P, M, means3D, None, sh, colors, scales, rotations, cov3D_precomp, sh_rest)
            stream = torch.cuda.current_stream(dev).cuda_stream
            factor = ex is not None and bool(M) and ex.sh_factors
            args = (C.byref(view), C.byref(g), radii.data_ptr(), geomBuffer.data_ptr(), binningBuffer.data_ptr(),
                    imageBuffer.data_ptr(), int(R), _ptr(dL_dout_color, "dL_dout_color"), b["means2D"].data_ptr(),
                    b["colors"].data_ptr(), b["opacity"].data_ptr(), b["means3D"].data_ptr(), b["cov3D"].data_ptr(),
                    b["sh"].data_ptr() if (M and not factor) else None, b["scales"].data_ptr(),
                    b["rotations"].data_ptr(), b["scratch"].data_ptr(), stream)
            if raw:
                plan = _lib.SgrBackwardPlan(_lib.STAGE_HOOK(0), None, 1, None, b["sh_rest"].data_ptr() if M > 1 else None)
                check(lib.sgr_rasterize_backward_staged(*args, C.byref(plan)))
            elif ex is None:
                check(lib.sgr_rasterize_backward(*args))
            else:
                ex.run_backward(lib, check, _lib.STAGE_HOOK, _lib.SgrBackwardPlan, args, b, P, M, int(degree), means3D,
                                campos, has_cov_precomp=has_cov)
    out = (b["means2D"], b["colors"], b["opacity"], b["means3D"], b["cov3D"], b["sh"], b["scales"], b["rotations"])
    return out + (b["sh_rest"],) if raw else out


def mark_visible(means3D, viewmatrix, projmatrix) -> torch.Tensor:
    """markVisible (rasterize_points.cu:198-216)."""
    P = means3D.size(0)
    present = torch.zeros((P,), dtype=torch.bool, device=means3D.device)
    if P != 0:
        with torch.cuda.device(means3D.device):
            means3D, viewmatrix, projmatrix = map(_c, (means3D, viewmatrix, projmatrix))
            check(lib.sgr_mark_visible(P, _ptr(means3D, "means3D"), _ptr(viewmatrix, "viewmatrix"),
                                       _ptr(projmatrix, "projmatrix"), present.data_ptr(),
                                       torch.cuda.current_stream(means3D.device).cuda_stream))
    return present


def inspect_state(P, W, H, R, geomBuffer, binningBuffer, imageBuffer):
    """Decode the opaque buffers into the reference's named arrays (parity tests only)."""
    dev = geomBuffer.device
    T = ((W + 15) // 16) * ((H + 15) // 16)
    f = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
    out = dict(depths=f(P), means2D=f(P, 2), conic_opacity=f(P, 4), rgb=f(P, 3),
               clamped=torch.zeros((P, 3), dtype=torch.uint8, device=dev),
               tiles_touched=torch.zeros(P, dtype=torch.int32, device=dev),
               keys=torch.zeros(max(R, 1), dtype=torch.int64, device=dev),
               point_list=torch.zeros(max(R, 1), dtype=torch.int32, device=dev),
               ranges=torch.zeros((T, 2), dtype=torch.int32, device=dev),
               final_T=f(H, W), n_contrib=torch.zeros((H, W), dtype=torch.int32, device=dev),
               footprint=torch.zeros(max(R, 1), dtype=torch.uint8, device=dev))
    with torch.cuda.device(dev):
        check(lib.sgr_inspect_state(P, W, H, R, geomBuffer.data_ptr(), binningBuffer.data_ptr(), imageBuffer.data_ptr(),
                                    out["depths"].data_ptr(), out["means2D"].data_ptr(),
                                    out["conic_opacity"].data_ptr(), out["rgb"].data_ptr(), out["clamped"].data_ptr(),
                                    out["tiles_touched"].data_ptr(), out["keys"].data_ptr(),
                                    out["point_list"].data_ptr(), out["ranges"].data_ptr(), out["final_T"].data_ptr(),
                                    out["n_contrib"].data_ptr(), out["footprint"].data_ptr(),
                                    torch.cuda.current_stream(dev).cuda_stream))
    out["keys"] = out["keys"][:R]
    out["point_list"] = out["point_list"][:R]
    out["footprint"] = out["footprint"][:R]
    return out
