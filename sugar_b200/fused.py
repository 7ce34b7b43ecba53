"""Raw-parameter mode of the rasterizer: a SuGaR model's parameters go into the op as they are stored.

The reference's trainers evaluate, in PyTorch, every step and before the rasterizer is called
(sugar_scene/sugar_model.py:400-479, 839-883; sugar_utils/spherical_harmonics.py:117-172):

    strengths   = sigmoid(all_densities)          scaling = exp(_scales)         quaternions = normalize(_quaternions)
    sh          = cat(_sh_coordinates_dc, _sh_coordinates_rest)
    colors      = clamp_min(eval_sh(normalize(points - camera_center), sh) + 0.5, 0)     (compute_color_in_rasterizer=False)

~30 elementwise kernels with P x 3 / P x 48 temporaries plus their autograd.  `rasterize_raw` hands the six
parameter tensors to the kernels instead: the per-Gaussian forward pass applies the activations to the values it
has staged in shared memory anyway and evaluates the SH colour (same polynomial, the rasterizer's own evaluation
order, forward.cu:20-71), reading the two SH arrays with one bulk copy each per CTA; the per-Gaussian backward pass
folds the activations' chain rule (d sigmoid, d exp, the normalisation's projection -- dnormvdv's rule,
auxiliary.h:99-132) and writes the gradients of the RAW parameters, dL_dsh split back into the (dc, rest) arrays.

    image, radii = rasterize_raw(points, means2D, sh_dc, sh_rest, densities, scales_raw, quats_raw, raster_settings)

`raster_settings` is the drop-in module's GaussianRasterizationSettings.  Results equal "activate in PyTorch, then
call GaussianRasterizer(shs=...)" up to fp32 rounding (tests/test_gpu_fused.py), and the reference's python colour
path up to the evaluation order of the SH polynomial (~1e-7 on the colours).
"""
import torch

from . import _C


class _RasterizeRaw(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, means2D, sh_dc, sh_rest, densities, scales_raw, quats_raw, raster_settings):
        rs = raster_settings
        absent = torch.Tensor([])
        ctx.sgr_context = _C.current_context()
        num_rendered, color, radii, geom, binning, img = _C.rasterize_gaussians(
            rs.bg, points, absent, densities, scales_raw, quats_raw, rs.scale_modifier, absent, rs.viewmatrix,
            rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, sh_dc, rs.sh_degree, rs.campos,
            rs.prefiltered, rs.debug, context=ctx.sgr_context, sh_rest=sh_rest)
        ctx.raster_settings, ctx.num_rendered = rs, num_rendered
        ctx.save_for_backward(points, scales_raw, quats_raw, radii, sh_dc, sh_rest, geom, binning, img)
        ctx.mark_non_differentiable(radii)
        ctx.set_materialize_grads(False)   # no [P] zeros tensor for radii's absent gradient on every backward
        return color, radii

    @staticmethod
    def backward(ctx, grad_color, _grad_radii):
        if grad_color is None:
            return (None,) * 8
        rs = ctx.raster_settings
        points, scales_raw, quats_raw, radii, sh_dc, sh_rest, geom, binning, img = ctx.saved_tensors
        absent = torch.Tensor([])
        (g_means2D, _g_colors, g_dens, g_points, _g_cov, g_dc, g_scales, g_quats, g_rest) = _C.rasterize_gaussians_backward(
            rs.bg, points, radii, absent, scales_raw, quats_raw, rs.scale_modifier, absent, rs.viewmatrix, rs.projmatrix,
            rs.tanfovx, rs.tanfovy, grad_color, sh_dc, rs.sh_degree, rs.campos, geom, ctx.num_rendered, binning, img,
            rs.debug, context=ctx.sgr_context, sh_rest=sh_rest)
        return g_points, g_means2D, g_dc, g_rest, g_dens, g_scales, g_quats, None


def rasterize_raw(points, means2D, sh_dc, sh_rest, densities, scales_raw, quats_raw, raster_settings):
    """points [P,3], means2D [P,3] (gradient holder, as in the reference), sh_dc [P,1,3], sh_rest [P,M-1,3],
    densities [P,1] (logits), scales_raw [P,3] (logs), quats_raw [P,4] (un-normalised, w first)."""
    if sh_dc.dim() != 3 or sh_dc.size(1) != 1:
        raise RuntimeError("sh_dc must be [P,1,3]")
    return _RasterizeRaw.apply(points, means2D, sh_dc, sh_rest, densities, scales_raw, quats_raw, raster_settings)
