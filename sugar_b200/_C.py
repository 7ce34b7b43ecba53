"""`_C`-compatible shim: the three symbols the reference's pybind module exports
(diff-gaussian-rasterization/ext.cpp:15-19, rasterize_points.h:19-67), same positional
signatures and return tuples, implemented over the C ABI of libsugar_b200.so.

torch is used only for device memory (caching allocator), the current stream and the device
guard; all arithmetic happens in the hand-written sm_100a kernels.
"""
import ctypes as C
import os
from typing import Tuple

import torch

from . import _lib
from ._lib import SgrGaussians, SgrView, check, lib

# Instance-capacity hint per device: after the first forward the binning buffer is sized from the
# previous view's instance count (x1.25 + slack) so the whole forward is enqueued without
# waiting for the device; the C side re-runs binning on overflow (include/sugar_b200.h).
_capacity_hint = {}
_USE_HINT = os.environ.get("SGR_NO_CAPACITY_HINT", "0") != "1"


def _ptr(t: torch.Tensor, name: str):
    """Device pointer of an optional tensor: empty tensor == absent == NULL (forward.cu:205,241)."""
    if t is None or t.numel() == 0:
        return None
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if not t.is_cuda:
        raise _lib.SgrError(f"{name} must be a CUDA tensor (sugar_b200 has no CPU path)")
    return t.data_ptr()


class _Arena:
    """Allocator callback state: keeps the torch tensor alive and hands its pointer to C."""

    def __init__(self, device):
        self.device = device
        self.tensor = torch.empty(0, dtype=torch.uint8, device=device)
        self.cb = _lib.ALLOC_FN(self._alloc)

    def _alloc(self, _ctx, nbytes):
        self.tensor = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self.tensor.data_ptr()


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


def _gauss_struct(P, M, means3D, opacity, sh, colors, scales, rotations, cov3D_precomp) -> SgrGaussians:
    g = SgrGaussians()
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
                        prefiltered, debug) -> Tuple[int, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor,
                                                     torch.Tensor]:
    """RasterizeGaussiansCUDA (rasterize_points.cu:36-115)."""
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if not means3D.is_cuda:
        raise _lib.SgrError("sugar_b200 needs CUDA tensors: there is no CPU fallback")
    dev = means3D.device
    P, H, W = means3D.size(0), int(image_height), int(image_width)
    M = sh.size(1) if (sh is not None and sh.numel() != 0) else 0
    with torch.cuda.device(dev):
        out_color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        geom, binning, img = _Arena(dev), _Arena(dev), _Arena(dev)
        means3D, colors, opacity, scales, rotations, cov3D_precomp, sh = map(
            _c, (means3D, colors, opacity, scales, rotations, cov3D_precomp, sh))
        background, viewmatrix, projmatrix, campos = map(_c, (background, viewmatrix, projmatrix, campos))
        view = _view_struct(background, scale_modifier, viewmatrix, projmatrix, tan_fovx, tan_fovy, H, W, degree,
                            campos, prefiltered, debug)
        g = _gauss_struct(P, M, means3D, opacity, sh, colors, scales, rotations, cov3D_precomp)
        key = (dev.index, H, W)
        hint = _capacity_hint.get(key, 0) if _USE_HINT else 0
        rendered = C.c_int64(0)
        stream = torch.cuda.current_stream(dev).cuda_stream
        check(lib.sgr_rasterize_forward(C.byref(view), C.byref(g), geom.cb, None, binning.cb, None, img.cb, None,
                                        out_color.data_ptr(), radii.data_ptr() if P else None, hint,
                                        C.byref(rendered), stream))
        R = int(rendered.value)
        if P:
            _capacity_hint[key] = int(R * 1.25) + 65536
    return R, out_color, radii, geom.tensor, binning.tensor, img.tensor


def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, scale_modifier,
                                 cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, sh,
                                 degree, campos, geomBuffer, R, binningBuffer, imageBuffer, debug):
    """RasterizeGaussiansBackwardCUDA (rasterize_points.cu:118-196)."""
    dev = means3D.device
    P, H, W = means3D.size(0), dL_dout_color.size(1), dL_dout_color.size(2)
    M = sh.size(1) if (sh is not None and sh.numel() != 0) else 0
    with torch.cuda.device(dev):
        opts = dict(dtype=torch.float32, device=dev)
        dL_dmeans3D = torch.empty((P, 3), **opts)
        dL_dmeans2D = torch.empty((P, 3), **opts)
        dL_dcolors = torch.empty((P, 3), **opts)
        dL_dopacity = torch.empty((P, 1), **opts)
        dL_dcov3D = torch.empty((P, 6), **opts)
        dL_dsh = torch.empty((P, M, 3), **opts)
        dL_dscales = torch.empty((P, 3), **opts)
        dL_drotations = torch.empty((P, 4), **opts)
        if P != 0:
            means3D, colors, scales, rotations, cov3D_precomp, sh, dL_dout_color = map(
                _c, (means3D, colors, scales, rotations, cov3D_precomp, sh, dL_dout_color))
            background, viewmatrix, projmatrix, campos = map(_c, (background, viewmatrix, projmatrix, campos))
            scratch = torch.empty(lib.sgr_backward_scratch_bytes(P), dtype=torch.uint8, device=dev)
            view = _view_struct(background, scale_modifier, viewmatrix, projmatrix, tan_fovx, tan_fovy, H, W, degree,
                                campos, False, debug)
            # opacities are not an input of the reference's backward; the forward stored them
            g = _gauss_struct(P, M, means3D, None, sh, colors, scales, rotations, cov3D_precomp)
            stream = torch.cuda.current_stream(dev).cuda_stream
            check(lib.sgr_rasterize_backward(
                C.byref(view), C.byref(g), radii.data_ptr(), geomBuffer.data_ptr(), binningBuffer.data_ptr(),
                imageBuffer.data_ptr(), int(R), _ptr(dL_dout_color, "dL_dout_color"), dL_dmeans2D.data_ptr(),
                dL_dcolors.data_ptr(), dL_dopacity.data_ptr(), dL_dmeans3D.data_ptr(), dL_dcov3D.data_ptr(),
                dL_dsh.data_ptr() if M else None, dL_dscales.data_ptr(), dL_drotations.data_ptr(),
                scratch.data_ptr(), stream))
    return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations


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
               final_T=f(H, W), n_contrib=torch.zeros((H, W), dtype=torch.int32, device=dev))
    with torch.cuda.device(dev):
        check(lib.sgr_inspect_state(P, W, H, R, geomBuffer.data_ptr(), binningBuffer.data_ptr(), imageBuffer.data_ptr(),
                                    out["depths"].data_ptr(), out["means2D"].data_ptr(),
                                    out["conic_opacity"].data_ptr(), out["rgb"].data_ptr(), out["clamped"].data_ptr(),
                                    out["tiles_touched"].data_ptr(), out["keys"].data_ptr(),
                                    out["point_list"].data_ptr(), out["ranges"].data_ptr(), out["final_T"].data_ptr(),
                                    out["n_contrib"].data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    out["keys"] = out["keys"][:R]
    out["point_list"] = out["point_list"][:R]
    return out
