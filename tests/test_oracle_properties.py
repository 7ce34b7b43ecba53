"""CPU: size-independent properties of the rasterizer restated in the oracle -- the same invariants the GPU
tests use at sizes where no element-wise reference is affordable.

 * background:   image(bg) = image(0) + final_T * bg, and final_T / n_contrib / radii do not depend on bg
                 (forward.cu:361-367);
 * permutation:  with pairwise distinct depths the image does not depend on the order of the Gaussians in memory
                 (sort key = tile | depth, rasterizer_impl.cu:70-111), bit for bit;
 * linearity:    the backward is linear in the upstream image gradient (backward.cu is a chain of products with
                 dL_dpixel), up to fp32 rounding;
 * culling:      Gaussians behind the near plane (z <= 0.2) get radius 0 and exactly zero gradients."""
import numpy as np

import helpers as h


def _scene(P=1500, W=96, H=64, seed=5):
    from sugar_b200 import scenes
    return scenes.make_scene(P, W, H, seed=seed, camera="posed")


def test_background_enters_linearly_through_final_T():
    sc = _scene()
    fw0, _ = h.run_oracle(sc, np.zeros(3, np.float32))
    bg = np.array([0.25, 0.5, 1.0], np.float32)
    fw1, _ = h.run_oracle(sc, bg)
    assert np.array_equal(fw0["final_T"], fw1["final_T"]) and np.array_equal(fw0["n_contrib"], fw1["n_contrib"])
    assert np.array_equal(fw0["radii"], fw1["radii"])
    want = fw0["color"] + fw0["final_T"][None] * bg[:, None, None]
    assert np.abs(fw1["color"] - want).max() <= 1e-6


def test_image_is_invariant_under_permutation_of_the_gaussians():
    sc = _scene(seed=8)
    fw0, _ = h.run_oracle(sc, np.zeros(3, np.float32))
    vis = fw0["radii"] > 0
    assert len(np.unique(fw0["depths"][vis])) == int(vis.sum()), "test scene must have distinct depths"
    perm = np.random.default_rng(0).permutation(sc.means3D.shape[0])
    sc2 = sc._replace(means3D=sc.means3D[perm], scales=sc.scales[perm], rotations=sc.rotations[perm],
                      opacities=sc.opacities[perm], shs=sc.shs[perm], colors_precomp=sc.colors_precomp[perm])
    fw1, _ = h.run_oracle(sc2, np.zeros(3, np.float32))
    assert fw1["num_rendered"] == fw0["num_rendered"]
    assert np.array_equal(fw1["radii"], fw0["radii"][perm])
    assert np.array_equal(fw1["color"].view(np.uint32), fw0["color"].view(np.uint32))
    assert np.array_equal(fw1["n_contrib"], fw0["n_contrib"])


def test_backward_is_linear_in_the_upstream_gradient():
    from sugar_b200 import scenes
    sc = _scene(seed=11)
    W, H = sc.width, sc.height
    d1, d2 = scenes.upstream_grad(W, H, seed=1), scenes.upstream_grad(W, H, seed=2)
    a, b = 0.75, -1.5
    _, g1 = h.run_oracle(sc, np.zeros(3, np.float32), d1)
    _, g2 = h.run_oracle(sc, np.zeros(3, np.float32), d2)
    _, g12 = h.run_oracle(sc, np.zeros(3, np.float32), (a * d1 + b * d2).astype(np.float32))
    for k in ("dL_dmeans2D", "dL_dopacity", "dL_dmeans3D", "dL_dsh", "dL_dscales", "dL_drotations", "dL_dcov3D"):
        assert h.rel_err(g12[k], a * g1[k] + b * g2[k]) <= 2e-5, k


def test_culled_gaussians_have_zero_radius_and_zero_gradients():
    from sugar_b200 import scenes
    sc = _scene(seed=13)
    fw, bw = h.run_oracle(sc, np.zeros(3, np.float32), scenes.upstream_grad(sc.width, sc.height))
    V = sc.viewmatrix.astype(np.float64).T
    z = (sc.means3D.astype(np.float64) @ V[:3, :3].T + V[:3, 3])[:, 2]
    behind = z <= 0.2 - 1e-4
    assert behind.any() and (fw["radii"][behind] == 0).all()
    culled = fw["radii"] == 0
    for k in ("dL_dmeans2D", "dL_dopacity", "dL_dmeans3D", "dL_dsh", "dL_dscales", "dL_drotations"):
        assert not np.any(bw[k][culled]), k


def test_linear_properties_checker_on_the_oracle():
    """helpers.check_linear_properties (the checker the GPU suite runs at the headline size) on the CPU oracle."""
    sc = _scene(P=1200, W=80, H=48, seed=17)

    def run(bg, dL):
        fw, bw = h.run_oracle(sc, np.asarray(bg, np.float32), dL)
        grads = None if bw is None else {k: bw[k] for k in ("dL_dmeans3D", "dL_dmeans2D", "dL_dopacity", "dL_dsh",
                                                            "dL_dscales", "dL_drotations")}
        return fw["color"], fw["radii"], grads
    h.check_linear_properties(run, sc.width, sc.height, tol_img=1e-6, tol_grad=2e-5)
