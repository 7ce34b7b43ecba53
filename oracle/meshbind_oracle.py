"""CPU restatement of a mesh-bound SuGaR model's derived tensors -- TEST INFRASTRUCTURE ONLY.

Follows, op for op in PyTorch:
    SuGaR.points       (bound)  sugar_scene/sugar_model.py:384-398
    SuGaR.scaling      (bound)  sugar_scene/sugar_model.py:415-441   (not editable)
    SuGaR.quaternions  (bound)  sugar_scene/sugar_model.py:443-479   (not editable)
Third-party arithmetic that is NOT under /root/reference (pytorch3d 0.7.4, environment.yml:161), restated from
its published algorithm:
    Meshes.faces_normals_list()[0]   (v1 - v0) x (v2 - v0), divided by max(norm, 1e-6)
    transforms.matrix_to_quaternion  candidates from q_abs = sqrt(max(0, 1 +- m00 +- m11 +- m22)), row of the
                                     largest q_abs over 2 max(q_abs, 0.1)
Parity pinning: tests/golden/meshbind_*.npz come from the reference's OWN property code (the three properties
called unbound on a duck-typed object, tests/golden/make_meshbind_golden.py) with only these two pytorch3d
functions substituted; tests/test_meshbind_oracle.py checks this file against them.
"""
import numpy as np
import torch


def faces_normals(verts: torch.Tensor, faces: torch.Tensor) -> torch.Tensor:
    v = verts[faces]
    n = torch.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0], dim=-1)
    return n / n.norm(dim=-1, keepdim=True).clamp(min=1e-6)


def _sqrt_positive_part(x: torch.Tensor) -> torch.Tensor:
    ret = torch.zeros_like(x)
    m = x > 0
    ret[m] = torch.sqrt(x[m])
    return ret


def matrix_to_quaternion(matrix: torch.Tensor) -> torch.Tensor:
    batch = matrix.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(matrix.reshape(batch + (9,)), dim=-1)
    q_abs = _sqrt_positive_part(torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22,
                                             1.0 - m00 - m11 + m22], dim=-1))
    quat_by_rijk = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1)], dim=-2)
    flr = torch.tensor(0.1).to(dtype=q_abs.dtype, device=q_abs.device)
    cand = quat_by_rijk / (2.0 * q_abs[..., None].max(flr))
    return cand[torch.nn.functional.one_hot(q_abs.argmax(dim=-1), num_classes=4) > 0.5, :].reshape(batch + (4,))


def bind_to_mesh_torch(verts, faces, bary, scales_raw, complex_raw, thickness):
    """bary [n,3]; returns (points [F*n,3], scaling [F*n,3], quaternions [F*n,4])."""
    F, n = faces.shape[0], bary.shape[0]
    faces_verts = verts[faces]                                                    # :391
    points = (faces_verts[:, None] * bary.reshape(n, 3, 1)[None]).sum(dim=-2).reshape(F * n, 3)   # :394-398
    plane = torch.exp(scales_raw)                                                 # :418
    scaling = torch.cat([thickness * torch.ones(len(scales_raw), 1, dtype=plane.dtype, device=plane.device), plane],
                        dim=-1)                                                      # :438-441
    R_0 = torch.nn.functional.normalize(faces_normals(verts, faces), dim=-1)       # :448
    base_R_1 = torch.nn.functional.normalize(faces_verts[:, 0] - faces_verts[:, 1], dim=-1)   # :452
    base_R_2 = torch.nn.functional.normalize(torch.cross(R_0, base_R_1, dim=-1))   # :455
    cplx = torch.nn.functional.normalize(complex_raw, dim=-1).view(F, n, 2)        # :458
    R_1 = cplx[..., 0:1] * base_R_1[:, None] + cplx[..., 1:2] * base_R_2[:, None]  # :459
    R_2 = -cplx[..., 1:2] * base_R_1[:, None] + cplx[..., 0:1] * base_R_2[:, None]
    R = torch.cat([R_0[:, None, ..., None].expand(-1, n, -1, -1).clone(), R_1[..., None], R_2[..., None]],
                  dim=-1).view(-1, 3, 3)                                            # :463-466
    q = matrix_to_quaternion(R)
    return points, scaling, torch.nn.functional.normalize(q, dim=-1)               # :479


def make_case(F=200, V=120, n_per=6, seed=0, dtype=torch.float32):
    """A seeded triangle soup over V vertices (non-degenerate faces), raw scales and in-plane rotations."""
    g = torch.Generator().manual_seed(seed)
    verts = torch.randn(V, 3, generator=g, dtype=dtype)
    faces = torch.stack([torch.randperm(V, generator=g)[:3] for _ in range(F)])
    P = F * n_per
    scales_raw = torch.randn(P, 2, generator=g, dtype=dtype) * 0.4 - 2.0
    complex_raw = torch.randn(P, 2, generator=g, dtype=dtype) * 1.5
    tables = {1: [[1 / 3, 1 / 3, 1 / 3]],
              3: [[1 / 2, 1 / 4, 1 / 4], [1 / 4, 1 / 2, 1 / 4], [1 / 4, 1 / 4, 1 / 2]],
              4: [[1 / 3, 1 / 3, 1 / 3], [2 / 3, 1 / 6, 1 / 6], [1 / 6, 2 / 3, 1 / 6], [1 / 6, 1 / 6, 2 / 3]],
              6: [[2 / 3, 1 / 6, 1 / 6], [1 / 6, 2 / 3, 1 / 6], [1 / 6, 1 / 6, 2 / 3], [1 / 6, 5 / 12, 5 / 12],
                  [5 / 12, 1 / 6, 5 / 12], [5 / 12, 5 / 12, 1 / 6]]}
    bary = torch.tensor(tables[n_per], dtype=dtype)
    return dict(verts=verts, faces=faces, bary=bary, scales_raw=scales_raw, complex_raw=complex_raw, thickness=1e-5)


def loss_weights(P, seed=0, dtype=torch.float32):
    g = torch.Generator().manual_seed(1000 + seed)  # always drawn in fp32: the same numbers for every dtype
    return tuple(torch.randn(P, w, generator=g).to(dtype) for w in (3, 3, 4))


def values_and_grads(case, seed=0):
    """Forward tensors + autograd gradients of the fixed scalar loss sum(w * output) -> numpy dict."""
    leaf = {k: case[k].clone().requires_grad_(True) for k in ("verts", "scales_raw", "complex_raw")}
    p, s, q = bind_to_mesh_torch(leaf["verts"], case["faces"], case["bary"], leaf["scales_raw"], leaf["complex_raw"],
                                 case["thickness"])
    wp, ws, wq = loss_weights(p.shape[0], seed, p.dtype)
    ((p * wp).sum() + (s * ws).sum() + (q * wq).sum()).backward()
    out = dict(points=p, scaling=s, quaternions=q)
    out.update({"g_" + k: v.grad for k, v in leaf.items()})
    return {k: v.detach().numpy() for k, v in out.items()}
