"""Exact K-nearest neighbours on the GPU (uniform grid), standing in for pytorch3d.ops.knn_points as
SuGaR uses it (sugar_scene/sugar_model.py:1013-1030 reset_neighbors, :1335-1343).

    dists, idx = knn_points(queries, points, K)     # [Q,K] squared distances (ascending), [Q,K] int64
    knn_dists, knn_idx = reset_neighbors(points, K=16)   # the tables SuGaR keeps (self is neighbour 0)
"""
import torch

from ._lib import check, lib


def knn_points(p1: torch.Tensor, p2: torch.Tensor, K: int):
    """K nearest points of p2 for every point of p1.  Accepts [N,3] or pytorch3d's batched [1,N,3]."""
    batched = p1.dim() == 3
    q = (p1[0] if batched else p1).contiguous().float()
    r = (p2[0] if p2.dim() == 3 else p2).contiguous().float()
    if not q.is_cuda or not r.is_cuda:
        raise RuntimeError("sugar_b200.knn needs CUDA tensors: there is no CPU fallback")
    Q, P, dev = q.shape[0], r.shape[0], q.device
    with torch.cuda.device(dev):
        idx = torch.empty((Q, K), dtype=torch.int64, device=dev)
        d2 = torch.empty((Q, K), dtype=torch.float32, device=dev)
        ws = torch.empty(lib.sgr_knn_workspace_bytes(P), dtype=torch.uint8, device=dev)
        check(lib.sgr_knn(P, r.data_ptr(), Q, q.data_ptr() if Q else None, K, idx.data_ptr() if Q else None,
                          d2.data_ptr() if Q else None, ws.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    return (d2[None], idx[None]) if batched else (d2, idx)


def reset_neighbors(points: torch.Tensor, K: int = 16):
    """SuGaR.reset_neighbors: (knn_dists [P,K], knn_idx [P,K]) of the cloud against itself, no grad."""
    with torch.no_grad():
        return knn_points(points, points, K)
