"""ORACLE — test infrastructure only (tests/, __graft_entry__.smoke(), bench.py cpu_baseline).
Nothing under foundationpose_cpp_b200/ imports this file; the product path is CUDA only.

CPU (numpy fp32) restatement of the non-network part of FoundationPose's Register/Track hot path
as implemented by zz990099/foundationpose_cpp.  Paths below are relative to
detection_6d_foundationpose/src/ of the reference; every function names the lines it follows.

PARITY STATUS: "parity unpinned" by the reference — its test-suite holds no golden vectors for this
path (simple_tests/src/test_foundationpose.cpp only CHECKs booleans).  What pins this oracle
instead: (1) the visibility pass and the per-pixel nvdiffrast kernels are compared on the GPU box
with the reference's own CUDA sources compiled in place (oracle/Makefile -> oracle/_ref,
tests/test_ref_kernels.py); (2) hand-computed known answers in tests/test_oracle.py; (3) the
warp is cross-checked against cv2.warpPerspective.  CV-CUDA (WarpPerspective/ConvertTo/Flip,
v0.12.0-beta) and TensorRT are third-party dependencies absent from the reference tree; their
arithmetic is restated from their documented behaviour.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from pathlib import Path

import numpy as np

F = np.float32
HERE = Path(__file__).resolve().parent

CROP = 160
MIN_DEPTH = F(0.001)   # foundationpose_render.hpp:26
MAX_DEPTH = F(4.0)     # foundationpose_render.hpp:27
ZNEAR, ZFAR = 0.1, 100.0  # foundationpose_render.cpp:155-156
W_AMBIENT, W_DIFFUSE = F(0.8), F(0.5)  # foundationpose_render.cpp:663
ROT_NORMALIZER = 0.349065850398865  # foundationpose.cpp:88
REFINE_CROP_RATIO, SCORE_CROP_RATIO = 1.2, 1.1  # foundationpose.cpp:84-85


# ----------------------------------------------------------------------------- hypotheses
def _normalize(v):
    v = np.asarray(v, dtype=F)
    return (v / F(np.sqrt(np.sum(v * v, dtype=F), dtype=F))).astype(F)


def icosphere_vertices(n_views: int = 40) -> np.ndarray:
    """foundationpose_sampling.cpp:56-121 (GenerateIcosphere).  Returns [42,3] f32."""
    t = F((1.0 + math.sqrt(5.0)) / 2.0)
    raw = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
           (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    verts = [_normalize(v) for v in raw]
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2),
             (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11),
             (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    cache = {}

    def mid(i, j):
        key = (min(i, j) << 32) + max(i, j)
        if key in cache:
            return cache[key]
        pm = ((verts[i] + verts[j]) / F(2.0)).astype(F)
        verts.append(_normalize(pm))
        cache[key] = len(verts) - 1
        return cache[key]

    while len(verts) < n_views:
        new_faces = []
        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            new_faces += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = new_faces
    return np.stack(verts).astype(F)


def rotation_grid(n_views: int = 40, inplane_step: int = 60) -> np.ndarray:
    """foundationpose_sampling.cpp:178-237 (SampleViewsIcosphere + MakeRotationGrid): [252,4,4] f32,
    icosphere-vertex major, in-plane angle minor; ClusterPoses' result is discarded there (:235)."""
    out = []
    for v in icosphere_vertices(n_views):
        cam = np.eye(4, dtype=F)
        cam[:3, 3] = v
        z = _normalize(-v)
        x = np.cross(np.array([0, 0, 1], dtype=F), z).astype(F)
        if not np.any(x):
            x = np.array([1, 0, 0], dtype=F)
        x = _normalize(x)
        y = _normalize(np.cross(z, x).astype(F))
        cam[:3, 0], cam[:3, 1], cam[:3, 2] = x, y, z
        ang = 0.0
        while ang < 360:
            a = F(ang * math.pi / 180.0)  # the reference multiplies in double then narrows for AngleAxisf
            c, s = F(np.cos(a, dtype=F)), F(np.sin(a, dtype=F))
            rz = np.eye(4, dtype=F)
            rz[0, 0], rz[0, 1], rz[1, 0], rz[1, 1] = c, -s, s, c
            cam_in_ob = (cam @ rz).astype(F)
            out.append(np.linalg.inv(cam_in_ob.astype(np.float64)).astype(F))
            ang += inplane_step
    return np.stack(out)


def erode_depth(depth, radius=2, diff_thres=0.001, ratio_thres=0.8, zfar=100.0):
    """foundationpose_sampling.cu:21-82."""
    d = depth.astype(F)
    h, w = d.shape
    bad_cnt = np.zeros((h, w), F)
    total = np.zeros((h, w), F)
    pad = np.pad(d, radius, constant_values=np.nan)
    for du in range(-radius, radius + 1):
        for dv in range(-radius, radius + 1):
            cur = pad[radius + dv: radius + dv + h, radius + du: radius + du + w]
            inside = ~np.isnan(cur)
            curz = np.where(inside, cur, F(0))
            total += inside.astype(F)
            bad = inside & ((curz < F(0.1)) | (curz >= F(zfar)) | (np.abs(curz - d) > F(diff_thres)))
            bad_cnt += bad.astype(F)
    out = np.where((bad_cnt / np.maximum(total, F(1))) > F(ratio_thres), F(0), d)
    out = np.where((d < F(0.1)) | (d >= F(zfar)), F(0), out)
    return out.astype(F)


def bilateral_filter_depth(depth, zfar=100.0, radius=2, sigma_d=2.0, sigma_r=100000.0):
    """foundationpose_sampling.cu:84-164."""
    d = depth.astype(F)
    h, w = d.shape
    pad = np.pad(d, radius, constant_values=np.nan)
    offs = [(du, dv) for du in range(-radius, radius + 1) for dv in range(-radius, radius + 1)]
    num_valid = np.zeros((h, w), np.int32)
    mean = np.zeros((h, w), F)
    for du, dv in offs:  # same (u outer, v inner) accumulation order as the kernel
        cur = pad[radius + dv: radius + dv + h, radius + du: radius + du + w]
        ok = ~np.isnan(cur) & (np.nan_to_num(cur) >= F(0.1)) & (np.nan_to_num(cur) < F(zfar))
        num_valid += ok
        mean = np.where(ok, mean + np.nan_to_num(cur), mean).astype(F)
    mean = (mean / np.maximum(num_valid, 1).astype(F)).astype(F)
    sum_w = np.zeros((h, w), F)
    acc = np.zeros((h, w), F)
    for du, dv in offs:
        cur = np.nan_to_num(pad[radius + dv: radius + dv + h, radius + du: radius + du + w]).astype(F)
        inside = ~np.isnan(pad[radius + dv: radius + dv + h, radius + du: radius + du + w])
        ok = inside & (cur >= F(0.1)) & (cur < F(zfar)) & (np.abs(cur - mean) < F(0.01))
        wgt = np.exp(-F(du * du + dv * dv) / F(2.0 * sigma_d * sigma_d)
                     - (d - cur) * (d - cur) / F(2.0 * sigma_r * sigma_r), dtype=F)
        sum_w = np.where(ok, sum_w + wgt, sum_w).astype(F)
        acc = np.where(ok, acc + wgt * cur, acc).astype(F)
    out = np.where((sum_w > 0) & (num_valid > 0), acc / np.where(sum_w > 0, sum_w, F(1)), F(0))
    return out.astype(F)


def guess_translation(depth, mask, K, min_depth=MIN_DEPTH):
    """foundationpose_sampling.cpp:250-298.  Returns f32[3] or None (empty mask / no valid depth)."""
    vs, us = np.nonzero(mask > 0)
    if us.size == 0:
        return None
    uc = F((int(us.min()) + int(us.max())) / 2.0)
    vc = F((int(vs.min()) + int(vs.max())) / 2.0)
    valid = (mask > 0) & (depth >= min_depth)
    if not valid.any():
        return None
    vals = np.sort(depth[valid].astype(F))
    n = vals.size
    zc = F((float(vals[n // 2 - 1]) + float(vals[n // 2])) / 2.0) if n % 2 == 0 else vals[n // 2]
    kinv = np.linalg.inv(np.asarray(K, np.float64)).astype(F)
    return np.array([F(F(F(kinv[r, 0] * uc) + F(kinv[r, 1] * vc)) + kinv[r, 2]) * zc for r in range(3)], F)


def sample_hypotheses(depth, mask, K):
    """FoundationPoseSampler::GetHypPoses, foundationpose_sampling.cpp:344-394."""
    filt = bilateral_filter_depth(erode_depth(depth))
    c = guess_translation(filt, mask, K)
    if c is None:
        return None
    poses = rotation_grid().copy()
    poses[:, :3, 3] = c
    return poses


# ----------------------------------------------------------------------------- image side
def depth_to_xyz(depth, K, min_depth=MIN_DEPTH):
    """foundationpose_utils.cu:3-32.  Invalid pixels are written as 0 here (the reference leaves
    them uninitialised) — downstream they are zeroed by z < min_depth anyway."""
    d = depth.astype(F)
    h, w = d.shape
    fx, fy, cx, cy = F(K[0][0]), F(K[1][1]), F(K[0][2]), F(K[1][2])
    u = np.arange(w, dtype=F)[None, :]
    v = np.arange(h, dtype=F)[:, None]
    x = ((u - cx) * d / fx).astype(F)
    y = ((v - cy) * d / fy).astype(F)
    xyz = np.stack([x, y, d], -1)
    xyz[d < min_depth] = 0
    return xyz.astype(F)


def crop_window_tf(poses, K, crop_ratio, diameter, out=CROP):
    """ComputeCropWindowTF + ComputeTF, foundationpose_render.cpp:25-75.
    Returns tf [N,3,3] f32 (source pixel -> crop pixel) and the rounded (left, right, top, bottom)."""
    poses = np.asarray(poses, F)
    n = poses.shape[0]
    r = F(F(diameter) * F(crop_ratio) / F(2))
    fx, k01, cx, fy, cy = F(K[0][0]), F(K[0][1]), F(K[0][2]), F(K[1][1]), F(K[1][2])
    tfs = np.zeros((n, 3, 3), F)
    boxes = np.zeros((n, 4), F)
    for i in range(n):
        t = poses[i, :3, 3]

        def proj(p):
            pu = F(F(fx * p[0] + k01 * p[1]) + cx * p[2])
            pv = F(fy * p[1] + cy * p[2])
            return F(pu / p[2]), F(pv / p[2])

        pts = [t, t + np.array([r, 0, 0], F), t + np.array([-r, 0, 0], F), t + np.array([0, r, 0], F),
               t + np.array([0, -r, 0], F)]
        uv = [proj(p.astype(F)) for p in pts]
        u0, v0 = uv[0]
        radius = F(abs(max(F(q[1] - v0) for q in uv)))
        rnd = lambda a: F(math.floor(abs(float(a)) + 0.5) * (1 if a >= 0 else -1))  # std::round
        left, right, top, bottom = rnd(F(u0 - radius)), rnd(F(u0 + radius)), rnd(F(v0 - radius)), rnd(F(v0 + radius))
        sx, sy = F(F(out) / F(right - left)), F(F(out) / F(bottom - top))
        tfs[i] = np.array([[sx, 0, F(sx * F(-left))], [0, sy, F(sy * F(-top))], [0, 0, 1]], F)
        boxes[i] = (left, right, top, bottom)
    return tfs, boxes


def bbox2d_from_tf(tfs, out=CROP):
    """ConstructBBox2D, foundationpose_render.cpp:123-149: tf^-1 applied to (0,0) and (W-1,H-1)."""
    n = tfs.shape[0]
    bb = np.zeros((n, 4), F)
    for i in range(n):
        sx, sy = tfs[i, 0, 0], tfs[i, 1, 1]
        left, top = F(-tfs[i, 0, 2] / sx), F(-tfs[i, 1, 2] / sy)
        bb[i] = (left, top, F(F(out - 1) / sx + left), F(F(out - 1) / sy + top))
    return bb


def projection_matrix(K, h, w, znear=ZNEAR, zfar=ZFAR):
    """ProjectMatrixFromIntrinsics, foundationpose_render.cpp:151-186 ("y_down")."""
    nc, fc = F(znear), F(zfar)
    depth = F(fc - nc)
    q = F(-(fc + nc) / depth)
    qn = F(F(-2) * F(fc * nc) / depth)
    k = np.asarray(K, F)
    wf, hf = F(w), F(h)
    return np.array([[F(2) * k[0, 0] / wf, F(-2) * k[0, 1] / wf, (F(-2) * k[0, 2] + wf) / wf, 0],
                     [0, F(2) * k[1, 1] / hf, (F(2) * k[1, 2] - hf) / hf, 0],
                     [0, 0, q, qn],
                     [0, 0, -1, 0]], F)


def _mat4_mul(a, b):
    """4x4 product with the k-sum evaluated left to right in fp32 (what the CUDA side does)."""
    out = np.zeros((4, 4), F)
    for i in range(4):
        for j in range(4):
            s = F(a[i, 0] * b[0, j])
            for k in range(1, 4):
                s = F(s + F(a[i, k] * b[k, j]))
            out[i, j] = s
    return out


GL_CAM_IN_CV_CAM = np.diag([1, -1, -1, 1]).astype(F)  # foundationpose_render.cpp:22-23


def clip_vertices(verts, pose, bbox, K, img_h, img_w):
    """generate_pose_clip_kernel, foundationpose_render.cu:363-398, with
    M = Proj * (kGLCamInCVCam * pose) (foundationpose_render.cpp:588-593).  verts [V,3] centred."""
    m = _mat4_mul(projection_matrix(K, img_h, img_w), _mat4_mul(GL_CAM_IN_CV_CAM, np.asarray(pose, F)))
    x, y, z = verts[:, 0].astype(F), verts[:, 1].astype(F), verts[:, 2].astype(F)

    def row(i):
        return (((m[i, 0] * x + m[i, 1] * y).astype(F) + m[i, 2] * z).astype(F) + m[i, 3]).astype(F)

    tx, ty, tz, tw = row(0), row(1), row(2), row(3)
    l, t, r, b = F(bbox[0]), F(F(img_h) - bbox[1]), F(bbox[2]), F(F(img_h) - bbox[3])
    a00, a11 = F(F(img_w) / F(r - l)), F(F(img_h) / F(t - b))
    a30, a31 = F(F(F(img_w) - r - l) / F(r - l)), F(F(F(img_h) - t - b) / F(t - b))
    return np.stack([(tx * a00 + tw * a30).astype(F), (ty * a11 + tw * a31).astype(F), tz, tw], -1).astype(F)


def transform_points(verts, pose):
    """transform_points_kernel, foundationpose_render.cu:321-341."""
    m = np.asarray(pose, F)
    x, y, z = verts[:, 0].astype(F), verts[:, 1].astype(F), verts[:, 2].astype(F)
    return np.stack([(((m[i, 0] * x + m[i, 1] * y).astype(F) + m[i, 2] * z).astype(F) + m[i, 3]).astype(F)
                     for i in range(3)], -1)


def diffuse_intensity(normals, pose):
    """transform_normals_kernel, foundationpose_render.cu:418-443."""
    m = np.asarray(pose, F)
    x, y, z = normals[:, 0].astype(F), normals[:, 1].astype(F), normals[:, 2].astype(F)
    t = [((m[i, 0] * x + m[i, 1] * y).astype(F) + m[i, 2] * z).astype(F) for i in range(3)]
    l2 = np.sqrt(((t[0] * t[0] + t[1] * t[1]).astype(F) + t[2] * t[2]).astype(F), dtype=F)
    val = np.where(l2 == 0, F(0), -t[2] / np.where(l2 == 0, F(1), l2))
    return np.clip(val, 0, 1).astype(F)


_RASTER_LIB = None


def build_c_oracle(force: bool = False) -> Path:
    """Compile oracle/raster_oracle.c -> oracle/libraster_oracle.so (gcc, no fp contraction)."""
    so = HERE / "libraster_oracle.so"
    src = HERE / "raster_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", str(so), str(src), "-lm"],
                       check=True)
    return so


_HOST_LIB = None


def host_prepost_lib():
    """oracle/host_prepost.cpp -> oracle/libhost_prepost.so: the reference's host-side functions of the path
    (GuessTranslation, crop window, bbox2d, projection, RefinePostProcess) in dependency-free C++, for the CPU
    baseline timing.  Checked against this file by tests/test_oracle.py."""
    global _HOST_LIB
    if _HOST_LIB is None:
        so, src = HERE / "libhost_prepost.so", HERE / "host_prepost.cpp"
        if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
            subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-o", str(so),
                            str(src)], check=True)
        _HOST_LIB = ctypes.CDLL(str(so))
        _HOST_LIB.host_register_work_ms.restype = ctypes.c_double
    return _HOST_LIB


def host_register_work_ms(depth, mask, K, grid_poses, diameter, refine_itr=1, threads=1, repeat=20):
    """Milliseconds of the reference's host-side work for one Register (see host_prepost.cpp)."""
    lib = host_prepost_lib()
    depth = np.ascontiguousarray(depth, F)
    mask = np.ascontiguousarray(mask, np.uint8)
    Kf = np.ascontiguousarray(np.asarray(K, F).reshape(9))
    poses = np.ascontiguousarray(np.asarray(grid_poses, F).reshape(-1, 16))
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    return float(lib.host_register_work_ms(p(depth), p(mask), depth.shape[0], depth.shape[1], p(Kf), p(poses),
                                           poses.shape[0], ctypes.c_float(diameter), refine_itr, threads, repeat))


def rasterize_ids(clip, faces, h=CROP, w=CROP):
    """cudaraster visibility (oracle/raster_oracle.c).  Returns (ids [h,w] i32 with row 0 = GL
    bottom row, number of triangles that took cudaraster's polygon clipper)."""
    global _RASTER_LIB
    if _RASTER_LIB is None:
        _RASTER_LIB = ctypes.CDLL(str(build_c_oracle()))
    clip = np.ascontiguousarray(clip, F)
    faces = np.ascontiguousarray(faces, np.int32)
    ids = np.zeros((h, w), np.int32)
    nclip = ctypes.c_int(0)
    rc = _RASTER_LIB.fpo_rasterize(clip.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(clip.shape[0]),
                                   faces.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(faces.shape[0]),
                                   ctypes.c_int(w), ctypes.c_int(h), ids.ctypes.data_as(ctypes.c_void_p),
                                   None, ctypes.byref(nclip))
    assert rc == 0
    return ids, nclip.value


def barycentrics(clip, faces, ids):
    """RasterizeCudaFwdShaderKernel, nvdiffrast/common/rasterize.cu:55-89.  Returns rast [h,w,4]."""
    h, w = ids.shape
    tri = np.clip(ids - 1, 0, None)
    valid = ids > 0
    p0, p1, p2 = clip[faces[tri, 0]], clip[faces[tri, 1]], clip[faces[tri, 2]]
    xs, xo = F(2.0) / F(w), F(F(1.0) / F(w) - F(1.0))
    ys, yo = F(2.0) / F(h), F(F(1.0) / F(h) - F(1.0))
    fx = (xs * np.arange(w, dtype=F) + xo).astype(F)[None, :]
    fy = (ys * np.arange(h, dtype=F) + yo).astype(F)[:, None]
    p0x, p0y = (p0[..., 0] - fx * p0[..., 3]).astype(F), (p0[..., 1] - fy * p0[..., 3]).astype(F)
    p1x, p1y = (p1[..., 0] - fx * p1[..., 3]).astype(F), (p1[..., 1] - fy * p1[..., 3]).astype(F)
    p2x, p2y = (p2[..., 0] - fx * p2[..., 3]).astype(F), (p2[..., 1] - fy * p2[..., 3]).astype(F)
    a0 = (p1x * p2y - p1y * p2x).astype(F)
    a1 = (p2x * p0y - p2y * p0x).astype(F)
    a2 = (p0x * p1y - p0y * p1x).astype(F)
    with np.errstate(divide="ignore", invalid="ignore"):
        iw = (F(1.0) / ((a0 + a1).astype(F) + a2).astype(F)).astype(F)
        b0 = np.clip(np.nan_to_num(a0 * iw, nan=0.0), 0, 1).astype(F)
        b1 = np.clip(np.nan_to_num(a1 * iw, nan=0.0), 0, 1).astype(F)
        z = ((p0[..., 2] * a0 + p1[..., 2] * a1).astype(F) + p2[..., 2] * a2).astype(F)
        ww = ((p0[..., 3] * a0 + p1[..., 3] * a1).astype(F) + p2[..., 3] * a2).astype(F)
        zw = np.clip(z / ww, -1, 1).astype(F)
    rast = np.stack([b0, b1, zw, ids.astype(F)], -1)
    rast[~valid] = 0
    return rast.astype(F)


def interpolate(attr, faces, rast):
    """InterpolateFwdKernel, nvdiffrast/common/interpolate.cu:46-76.  attr [V,C] -> [h,w,C]."""
    ids = rast[..., 3].astype(np.int32)
    valid = ids > 0
    tri = np.clip(ids - 1, 0, None)
    b0, b1 = rast[..., 0:1], rast[..., 1:2]
    b2 = (F(1.0) - b0 - b1).astype(F)
    a0, a1, a2 = attr[faces[tri, 0]], attr[faces[tri, 1]], attr[faces[tri, 2]]
    out = ((b0 * a0 + b1 * a1).astype(F) + b2 * a2).astype(F)
    out[~valid] = 0
    return out


def texture_linear(tex, uv):
    """TextureFwdKernelLinear1 with WRAP boundary, nvdiffrast/common/texture.cu:19-96,131-176.
    tex [th,tw,3] f32 (already /255), uv [h,w,2]."""
    th, tw, _ = tex.shape
    u, v = uv[..., 0].astype(F), uv[..., 1].astype(F)
    u = (u - np.floor(u)).astype(F)
    v = (v - np.floor(v)).astype(F)
    u = (u * F(tw) - F(0.5)).astype(F)
    v = (v * F(th) - F(0.5)).astype(F)
    iu0, iv0 = np.floor(u).astype(np.int32), np.floor(v).astype(np.int32)
    iu1, iv1 = iu0 + 1, iv0 + 1
    fu, fv = (u - iu0.astype(F)).astype(F)[..., None], (v - iv0.astype(F)).astype(F)[..., None]
    iu0 = np.where(iu0 < 0, iu0 + tw, iu0)
    iv0 = np.where(iv0 < 0, iv0 + th, iv0)
    iu1 = np.where(iu1 >= tw, iu1 - tw, iu1)
    iv1 = np.where(iv1 >= th, iv1 - th, iv1)
    a00, a10, a01, a11 = tex[iv0, iu0], tex[iv0, iu1], tex[iv1, iu0], tex[iv1, iu1]
    top = (a00 + fu * (a10 - a00)).astype(F)
    bot = (a01 + fu * (a11 - a01)).astype(F)
    return (top + fv * (bot - top)).astype(F)


def render_crop(mesh, pose, tf, K, img_h, img_w):
    """NvdiffrastRender, foundationpose_render.cpp:611-683: returns rgb [160,160,3], xyz [160,160,3]
    (both already vertically flipped to image convention) and the id map (GL row order)."""
    bbox = bbox2d_from_tf(tf[None])[0]
    clip = clip_vertices(mesh["vertices"], pose, bbox, K, img_h, img_w)
    ids, nclip = rasterize_ids(clip, mesh["faces"])
    rast = barycentrics(clip, mesh["faces"], ids)
    xyz = interpolate(transform_points(mesh["vertices"], pose), mesh["faces"], rast)
    uv = interpolate(mesh["texcoords"], mesh["faces"], rast)
    col = texture_linear(mesh["texture"], uv)
    dif = interpolate(diffuse_intensity(mesh["normals"], pose)[:, None], mesh["faces"], rast)
    fg = np.clip(rast[..., 3:4], 0, 1)
    col = np.clip((col * (W_AMBIENT + dif * W_DIFFUSE).astype(F)).astype(F) * fg, 0, 1).astype(F)
    return col[::-1].copy(), xyz[::-1].copy(), ids, nclip


def warp_crop(rgb_u8, xyz, tf, box):
    """TransfProcess, foundationpose_render.cpp:731-797 (cvcuda::WarpPerspective + ConvertTo).
    CV-CUDA semantics restated: dst(x,y) samples src at tf^-1 (x,y,1); rgb bilinear in float with
    constant-0 border then saturate-cast to u8 (round to nearest even) then * 1/255; xyz nearest
    (floor(c+0.5)), constant-0 border.  tf is a pure scale + integer shift, so its inverse is taken
    analytically: src = (x / sx + left, y / sy + top) with box = (left, right, top, bottom)."""
    hi, wi, _ = rgb_u8.shape
    inv_sx, inv_sy = F(F(1.0) / tf[0, 0]), F(F(1.0) / tf[1, 1])
    left, top = F(box[0]), F(box[2])
    xs = np.arange(CROP, dtype=F)[None, :].repeat(CROP, 0)
    ys = np.arange(CROP, dtype=F)[:, None].repeat(CROP, 1)
    sx = ((inv_sx * xs).astype(F) + left).astype(F)
    sy = ((inv_sy * ys).astype(F) + top).astype(F)
    # bilinear rgb
    x0, y0 = np.floor(sx).astype(np.int64), np.floor(sy).astype(np.int64)
    fx, fy = (sx - x0.astype(F)).astype(F), (sy - y0.astype(F)).astype(F)
    src = rgb_u8.astype(F)

    def fetch(yy, xx):
        ok = (xx >= 0) & (xx < wi) & (yy >= 0) & (yy < hi)
        return np.where(ok[..., None], src[np.clip(yy, 0, hi - 1), np.clip(xx, 0, wi - 1)], F(0))

    w00 = ((F(1) - fx) * (F(1) - fy)).astype(F)[..., None]
    w10 = (fx * (F(1) - fy)).astype(F)[..., None]
    w01 = ((F(1) - fx) * fy).astype(F)[..., None]
    w11 = (fx * fy).astype(F)[..., None]
    acc = (fetch(y0, x0) * w00).astype(F)
    acc = (acc + fetch(y0, x0 + 1) * w10).astype(F)
    acc = (acc + fetch(y0 + 1, x0) * w01).astype(F)
    acc = (acc + fetch(y0 + 1, x0 + 1) * w11).astype(F)
    q = np.clip(np.rint(acc), 0, 255).astype(F)
    rgb = (q * F(1.0 / 255.0)).astype(F)
    # nearest xyz
    xn, yn = np.floor(sx + F(0.5)).astype(np.int64), np.floor(sy + F(0.5)).astype(np.int64)
    ok = (xn >= 0) & (xn < wi) & (yn >= 0) & (yn < hi)
    xyz_c = np.where(ok[..., None], xyz[np.clip(yn, 0, hi - 1), np.clip(xn, 0, wi - 1)], F(0)).astype(F)
    return rgb, xyz_c


def normalize_pack(rgb, xyz, pose, diameter):
    """threshold_and_downscale_pointcloud + concat, foundationpose_render.cu:61-140.  [160,160,6]."""
    t = np.asarray(pose, F)[:3, 3]
    invalid = xyz[..., 2] < MIN_DEPTH
    p = ((xyz - t).astype(F) / F(F(diameter) / F(2))).astype(F)
    p = np.where(invalid[..., None] | (np.abs(p) > MAX_DEPTH), F(0), p)
    return np.concatenate([rgb.astype(F), p.astype(F)], -1)


def render_and_transform(mesh, poses, rgb_u8, xyz_map, K, crop_ratio):
    """FoundationPoseRenderer::RenderAndTransform, foundationpose_render.cpp:814-857.
    Returns render_input, transf_input [N,160,160,6] f32 plus debug info."""
    hi, wi, _ = rgb_u8.shape
    tfs, boxes = crop_window_tf(poses, K, crop_ratio, mesh["diameter"])
    ren, tra, idmaps, nclip_total = [], [], [], 0
    for i in range(len(poses)):
        col, xyz_r, ids, nclip = render_crop(mesh, poses[i], tfs[i], K, hi, wi)
        nclip_total += nclip
        ren.append(normalize_pack(col, xyz_r, poses[i], mesh["diameter"]))
        rgb_c, xyz_c = warp_crop(rgb_u8, xyz_map, tfs[i], boxes[i])
        tra.append(normalize_pack(rgb_c, xyz_c, poses[i], mesh["diameter"]))
        idmaps.append(ids)
    return np.stack(ren), np.stack(tra), {"tf": tfs, "ids": np.stack(idmaps), "clip_path": nclip_total}


# ----------------------------------------------------------------------------- post-processing
def pose_update(poses, trans, rot, diameter):
    """RefinePostProcess, foundationpose.cpp:360-406.  A zero rotation vector gives the identity
    (Eigen's normalized() of a zero vector is the zero vector -> identity matrix)."""
    out = np.array(poses, F, copy=True)
    for i in range(len(poses)):
        out[i, :3, 3] = (out[i, :3, 3] + np.asarray(trans[i], F) * F(F(diameter) / F(2))).astype(F)
        v = (np.tanh(np.asarray(rot[i], F), dtype=F) * F(ROT_NORMALIZER)).astype(F)
        ang = F(np.sqrt(np.sum(v * v, dtype=F), dtype=F))
        if ang > 0:
            ax = (v / ang).astype(F)
            c, s = F(np.cos(ang, dtype=F)), F(np.sin(ang, dtype=F))
            kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]], F)
            rm = (c * np.eye(3, dtype=F) + s * kx + (F(1) - c) * np.outer(ax, ax)).astype(F)
        else:
            rm = np.eye(3, dtype=F)
        out[i, :3, :3] = (rm.T @ out[i, :3, :3]).astype(F)
    return out


def argmax_first(scores):
    """getMaxScoreIndex, foundationpose_decoder.cu:24-35 (thrust::max_element: first maximum)."""
    return int(np.argmax(np.asarray(scores)))
