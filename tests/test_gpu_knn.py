"""GPU: exact K-NN (uniform grid) vs the oracle's restatement of pytorch3d.knn_points (cdist + topk, float64)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _check(points, queries, K):
    from oracle import field_oracle as fo
    from sugar_b200 import knn
    d2, idx = knn.knn_points(queries.cuda(), points.cuda(), K)
    d2, idx = d2.cpu(), idx.cpu()
    D = torch.cdist(queries.double(), points.double()) ** 2
    ref_d, ref_i = D.topk(K, dim=1, largest=False)
    assert torch.allclose(d2.double(), ref_d, rtol=1e-5, atol=1e-9), float((d2.double() - ref_d).abs().max())
    # the returned indices must realise the returned distances
    got = torch.gather(D, 1, idx)
    assert torch.allclose(got, ref_d, rtol=1e-5, atol=1e-9)
    # where the K-th gap is clear, the index sets are identical
    for q in range(0, queries.shape[0], max(1, queries.shape[0] // 200)):
        assert len(set(idx[q].tolist())) == K
    assert bool((d2[:, 1:] >= d2[:, :-1]).all()), "not sorted by distance"


@pytest.mark.parametrize("P,Q,K", [(5000, 0, 16), (20000, 3000, 16), (777, 500, 5), (3000, 1000, 64), (40, 30, 16)])
def test_knn_matches_exact(P, Q, K):
    g = torch.Generator().manual_seed(P)
    pts = torch.randn(P, 3, generator=g)
    if Q == 0:   # reset_neighbors: cloud against itself, self must be neighbour 0 at distance 0
        from sugar_b200 import knn
        d2, idx = knn.reset_neighbors(pts.cuda(), K)
        assert bool((idx[:, 0].cpu() == torch.arange(P)).all()) and float(d2[:, 0].abs().max()) == 0.0
        _check(pts, pts, K)
    else:
        qs = torch.randn(Q, 3, generator=g) * 1.5   # some queries outside the cloud's bounding box
        _check(pts, qs, K)


def test_knn_surface_like_and_degenerate_clouds():
    g = torch.Generator().manual_seed(3)
    # points on a thin sheet (most grid cells empty), plus a flat (zero-extent axis) cloud
    sheet = torch.randn(15000, 3, generator=g) * torch.tensor([1.0, 1.0, 1e-3])
    _check(sheet, sheet[:2000] + 0.01 * torch.randn(2000, 3, generator=g), 16)
    flat = torch.randn(4000, 3, generator=g); flat[:, 2] = 0.5
    _check(flat, flat[:1000], 8)
    # heavy duplicates
    dup = torch.randn(50, 3, generator=g).repeat(40, 1)
    _check(dup, dup[:300], 16)


def test_knn_planar_and_elongated_clouds_stay_fast():
    """Cubic cells with per-axis counts: a coplanar cloud (what SuGaR's surface-aligned Gaussians look like) or a
    cloud stretched along one axis by an outlier must not degrade to a linear scan per query (the shell bound was
    r * the smallest cell edge, i.e. ~0 on a degenerate axis)."""
    import time
    from sugar_b200 import knn
    g = torch.Generator().manual_seed(9)
    P = 400_000
    planar = torch.rand(P, 3, generator=g); planar[:, 2] = 0.25
    line = torch.rand(P, 3, generator=g) * torch.tensor([1.0, 1e-4, 1e-4])
    outlier = torch.randn(P, 3, generator=g); outlier[0] = torch.tensor([1e4, 0.0, 0.0])
    for name, pts in (("planar", planar), ("line", line), ("outlier", outlier)):
        pts = pts.cuda()
        knn.reset_neighbors(pts[:1000], 4)      # warm-up (context, allocator)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d2, idx = knn.reset_neighbors(pts, 16)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert dt < 2.0, f"{name}: reset_neighbors took {dt:.2f} s for {P} points"
        assert bool((idx[:, 0] == torch.arange(P, device="cuda")).all()) or float(d2[:, 0].max()) == 0.0
        sub = torch.randperm(P, generator=g)[:300]
        D = torch.cdist(pts[sub].double(), pts.double()) ** 2
        ref = D.topk(16, dim=1, largest=False).values
        assert torch.allclose(d2[sub].double(), ref, rtol=1e-4, atol=1e-12), name
