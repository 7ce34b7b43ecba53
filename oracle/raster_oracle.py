"""numpy front-end of oracle/raster_oracle.c (CPU restatement of the reference rasterizer).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg; never by sugar_b200/.  Mirrors the reference's two entry points
(`RasterizeGaussiansCUDA`, `RasterizeGaussiansBackwardCUDA`, DGR/rasterize_points.cu:36-196)
but returns every intermediate the reference keeps in its geometry/binning/image buffers so
that tests can compare stage by stage.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libraster_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "raster_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_preprocess.restype = C.c_int64
    return _lib


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def forward(means3D, opacities, viewmatrix, projmatrix, campos, bg, W, H, tanfovx, tanfovy,
            shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None,
            sh_degree=0, scale_modifier=1.0):
    """Full forward.  Arrays are numpy fp32; viewmatrix/projmatrix are the 16 floats exactly as the
    reference receives them (row-vector convention, read column-major by the kernels)."""
    L = lib()
    means3D = _f(means3D); P = means3D.shape[0]
    opacities = _f(opacities).reshape(-1)
    shs = _f(shs); colors_precomp = _f(colors_precomp); scales = _f(scales); rotations = _f(rotations)
    cov3D_precomp = _f(cov3D_precomp)
    viewmatrix = _f(viewmatrix).reshape(-1); projmatrix = _f(projmatrix).reshape(-1)
    campos = _f(campos).reshape(-1); bg = _f(bg).reshape(-1)
    M = 0 if shs is None else shs.shape[1]
    T = ((W + 15) // 16) * ((H + 15) // 16)
    o = dict(P=P, W=W, H=H, M=M, D=sh_degree)
    o["radii"] = np.zeros(P, np.int32)
    o["means2D"] = np.zeros((P, 2), np.float32)
    o["depths"] = np.zeros(P, np.float32)
    o["cov3D"] = np.zeros((P, 6), np.float32)
    o["rgb"] = np.zeros((P, 3), np.float32)
    o["conic_opacity"] = np.zeros((P, 4), np.float32)
    o["clamped"] = np.zeros((P, 3), np.uint8)
    o["tiles_touched"] = np.zeros(P, np.uint32)
    o["point_offsets"] = np.zeros(P, np.uint32)
    R = L.oracle_preprocess(
        C.c_int(P), C.c_int(sh_degree), C.c_int(M), _p(means3D), _p(scales), C.c_float(scale_modifier),
        _p(rotations), _p(opacities), _p(shs), _p(cov3D_precomp), _p(colors_precomp), _p(viewmatrix),
        _p(projmatrix), _p(campos), C.c_int(W), C.c_int(H), C.c_float(tanfovx), C.c_float(tanfovy),
        _p(o["radii"]), _p(o["means2D"]), _p(o["depths"]), _p(o["cov3D"]), _p(o["rgb"]),
        _p(o["conic_opacity"]), _p(o["clamped"]), _p(o["tiles_touched"]), _p(o["point_offsets"]))
    o["num_rendered"] = int(R)
    o["keys"] = np.zeros(max(R, 1), np.uint64)[:R]
    o["point_list"] = np.zeros(max(R, 1), np.uint32)[:R]
    o["ranges"] = np.zeros((T, 2), np.uint32)
    keys = np.zeros(max(R, 1), np.uint64); plist = np.zeros(max(R, 1), np.uint32)
    L.oracle_binning(C.c_int(P), C.c_int(W), C.c_int(H), _p(o["radii"]), _p(o["means2D"]), _p(o["depths"]),
                     _p(o["point_offsets"]), C.c_int64(R), _p(keys), _p(plist), _p(o["ranges"]))
    o["keys"] = keys[:R]; o["point_list"] = plist[:R]
    colors = colors_precomp if colors_precomp is not None else o["rgb"]
    o["colors"] = colors
    o["final_T"] = np.zeros((H, W), np.float32)
    o["n_contrib"] = np.zeros((H, W), np.uint32)
    o["color"] = np.zeros((3, H, W), np.float32)
    L.oracle_render(C.c_int(W), C.c_int(H), _p(o["ranges"]), _p(plist), _p(o["means2D"]), _p(colors),
                    _p(o["conic_opacity"]), _p(bg), _p(o["final_T"]), _p(o["n_contrib"]), _p(o["color"]))
    o["_in"] = dict(means3D=means3D, opacities=opacities, shs=shs, colors_precomp=colors_precomp, scales=scales,
                    rotations=rotations, cov3D_precomp=cov3D_precomp, viewmatrix=viewmatrix,
                    projmatrix=projmatrix, campos=campos, bg=bg, tanfovx=tanfovx, tanfovy=tanfovy,
                    scale_modifier=scale_modifier)
    return o


def backward(fw, dL_dout_color):
    """Backward for a forward() result.  Returns the eight tensors of
    RasterizeGaussiansBackwardCUDA plus the internal dL_dconic."""
    L = lib()
    i = fw["_in"]; P, W, H, M, D = fw["P"], fw["W"], fw["H"], fw["M"], fw["D"]
    dpix = _f(dL_dout_color).reshape(3, H, W)
    g = dict(
        dL_dmeans2D=np.zeros((P, 3), np.float32), dL_dconic=np.zeros((P, 2, 2), np.float32),
        dL_dopacity=np.zeros((P, 1), np.float32), dL_dcolors=np.zeros((P, 3), np.float32),
        dL_dmeans3D=np.zeros((P, 3), np.float32), dL_dcov3D=np.zeros((P, 6), np.float32),
        dL_dsh=np.zeros((P, M, 3), np.float32), dL_dscales=np.zeros((P, 3), np.float32),
        dL_drotations=np.zeros((P, 4), np.float32))
    plist = np.ascontiguousarray(fw["point_list"])
    L.oracle_render_backward(C.c_int(P), C.c_int(W), C.c_int(H), _p(fw["ranges"]), _p(plist), _p(i["bg"]),
                             _p(fw["means2D"]), _p(fw["conic_opacity"]), _p(fw["colors"]), _p(fw["final_T"]),
                             _p(fw["n_contrib"]), _p(dpix), _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]),
                             _p(g["dL_dopacity"]), _p(g["dL_dcolors"]))
    cov3Ds = i["cov3D_precomp"] if i["cov3D_precomp"] is not None else fw["cov3D"]
    L.oracle_preprocess_backward(
        C.c_int(P), C.c_int(D), C.c_int(M), _p(i["means3D"]), _p(fw["radii"]), _p(i["shs"]), _p(fw["clamped"]),
        _p(i["scales"]), _p(i["rotations"]), C.c_float(i["scale_modifier"]), _p(cov3Ds), _p(i["viewmatrix"]),
        _p(i["projmatrix"]), C.c_int(W), C.c_int(H), C.c_float(i["tanfovx"]), C.c_float(i["tanfovy"]),
        _p(i["campos"]), _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]), _p(g["dL_dcolors"]), _p(g["dL_dmeans3D"]),
        _p(g["dL_dcov3D"]), _p(g["dL_dsh"]), _p(g["dL_dscales"]), _p(g["dL_drotations"]))
    return g


def mark_visible(means3D, viewmatrix, projmatrix):
    L = lib()
    means3D = _f(means3D); P = means3D.shape[0]
    out = np.zeros(P, np.uint8)
    L.oracle_mark_visible(C.c_int(P), _p(means3D), _p(_f(viewmatrix).reshape(-1)), _p(_f(projmatrix).reshape(-1)),
                          _p(out))
    return out.astype(bool)
