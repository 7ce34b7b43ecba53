"""Gaussians bound to a triangle mesh: the per-step prologue of SuGaR's refinement stage, fused.

A mesh-bound `SuGaR` model derives every Gaussian from its triangle (sugar_scene/sugar_model.py):

    points       :384-398   barycentric combination of the face's vertices (n Gaussians per face, :172-214)
    scaling      :415-441   (surface_mesh_thickness, exp(_scales[:, 0]), exp(_scales[:, 1]))
    quaternions  :443-479   R = [face normal | first edge rotated in-plane by the learned complex number |
                            their cross product] -> matrix_to_quaternion -> normalize

as ~40 PyTorch kernels with [F, n, 3, 3] temporaries and their autograd.  `bind_to_mesh` computes the three
tensors with one CUDA kernel and one for the backward (sgr_meshbind.cu, one thread per face), gradients to the
vertices, `_scales` and `_quaternions`.

    b = bind_to_mesh(verts, faces, bary, scales_raw, complex_raw, thickness)
    b.points [F*n,3], b.scaling [F*n,3], b.quaternions [F*n,4]
"""
import math
from types import SimpleNamespace

import torch

from ._lib import check, lib


def bary_coords(n_gaussians_per_surface_triangle: int, device=None) -> torch.Tensor:
    """surface_triangle_bary_coords of the reference (sugar_model.py:172-214), as [n,3]."""
    tables = {1: [[1 / 3, 1 / 3, 1 / 3]],
              3: [[1 / 2, 1 / 4, 1 / 4], [1 / 4, 1 / 2, 1 / 4], [1 / 4, 1 / 4, 1 / 2]],
              4: [[1 / 3, 1 / 3, 1 / 3], [2 / 3, 1 / 6, 1 / 6], [1 / 6, 2 / 3, 1 / 6], [1 / 6, 1 / 6, 2 / 3]],
              6: [[2 / 3, 1 / 6, 1 / 6], [1 / 6, 2 / 3, 1 / 6], [1 / 6, 1 / 6, 2 / 3], [1 / 6, 5 / 12, 5 / 12],
                  [5 / 12, 1 / 6, 5 / 12], [5 / 12, 5 / 12, 1 / 6]]}
    return torch.tensor(tables[n_gaussians_per_surface_triangle], dtype=torch.float32, device=device)


class _MeshBind(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, faces, bary, scales_raw, complex_raw, thickness):
        if not verts.is_cuda:
            raise RuntimeError("sugar_b200.meshbind needs CUDA tensors: there is no CPU fallback")
        verts, scales_raw, complex_raw = (t.contiguous().float() for t in (verts, scales_raw, complex_raw))
        faces = faces.contiguous().long()
        bary = bary.reshape(-1, 3).contiguous().float()
        F, n, V = faces.shape[0], bary.shape[0], verts.shape[0]
        P = F * n
        if scales_raw.shape != (P, 2) or complex_raw.shape != (P, 2):
            raise RuntimeError(f"_scales / _quaternions of a mesh-bound model must be [{P}, 2]")
        dev = verts.device
        with torch.cuda.device(dev):
            points = torch.empty((P, 3), device=dev)
            scaling = torch.empty((P, 3), device=dev)
            quats = torch.empty((P, 4), device=dev)
            check(lib.sgr_meshbind_forward(F, n, V, verts.data_ptr(), faces.data_ptr(), bary.data_ptr(),
                                           scales_raw.data_ptr(), complex_raw.data_ptr(), float(thickness),
                                           points.data_ptr(), scaling.data_ptr(), quats.data_ptr(),
                                           torch.cuda.current_stream(dev).cuda_stream))
        ctx.save_for_backward(verts, faces, bary, scales_raw, complex_raw)
        ctx.mark_non_differentiable(faces)
        return points, scaling, quats

    @staticmethod
    def backward(ctx, g_points, g_scaling, g_quats):
        verts, faces, bary, scales_raw, complex_raw = ctx.saved_tensors
        F, n, V = faces.shape[0], bary.shape[0], verts.shape[0]
        dev = verts.device
        z = lambda g, ref: torch.zeros_like(ref) if g is None else g.contiguous().float()
        with torch.cuda.device(dev):
            g_points = z(g_points, torch.empty((F * n, 3), device=dev))
            g_scaling = z(g_scaling, torch.empty((F * n, 3), device=dev))
            g_quats = z(g_quats, torch.empty((F * n, 4), device=dev))
            g_verts = torch.empty_like(verts)
            g_s = torch.empty_like(scales_raw)
            g_c = torch.empty_like(complex_raw)
            check(lib.sgr_meshbind_backward(F, n, V, verts.data_ptr(), faces.data_ptr(), bary.data_ptr(),
                                            scales_raw.data_ptr(), complex_raw.data_ptr(), g_points.data_ptr(),
                                            g_scaling.data_ptr(), g_quats.data_ptr(), g_verts.data_ptr(), g_s.data_ptr(),
                                            g_c.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        return g_verts, None, None, g_s, g_c, None


def bind_to_mesh(verts, faces, bary, scales_raw, complex_raw, thickness):
    """verts [V,3] (SuGaR._points of a bound model), faces [F,3] int64 (_surface_mesh_faces), bary [n,3] or the
    reference's [n,3,1], scales_raw [F*n,2] (_scales), complex_raw [F*n,2] (_quaternions), thickness (float:
    surface_mesh_thickness).  Returns namespace(points, scaling, quaternions)."""
    points, scaling, quats = _MeshBind.apply(verts, faces, bary, scales_raw, complex_raw, float(thickness))
    return SimpleNamespace(points=points, scaling=scaling, quaternions=quats)
