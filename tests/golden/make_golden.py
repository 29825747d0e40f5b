"""Generates tests/golden/ref_kernels_golden.npz by running the REFERENCE's own CUDA kernels
(oracle/_ref/libref_kernels.so, built from /root/reference by oracle/Makefile) on seeded inputs.
Run on a GPU box:   python tests/golden/make_golden.py gpurun_out/ref_kernels_golden.npz
                    python tests/golden/make_golden.py clip gpurun_out/ref_clip_golden.npz
                    python tests/golden/make_golden.py render gpurun_out/ref_render_golden.npz
The inputs are stored next to the outputs so the CPU oracle can be checked without a GPU."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fp_helpers as helpers  # noqa: E402
from foundationpose_cpp_b200 import synth  # noqa: E402
from oracle import fp_oracle as O  # noqa: E402


def ref_lib():
    return C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_kernels.so"))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def ptr(t):
    return C.c_void_p(t.data_ptr())


def golden_inputs():
    mesh = synth.uv_ellipsoid("g", slices=24, stacks=13, tex_size=64, seed=5)
    rgb, depth, mask, gt = synth.make_scene(mesh)
    om = helpers.oracle_mesh(mesh)
    K = synth.DEFAULT_K
    poses = np.concatenate([gt[None], O.sample_hypotheses(depth, mask, K)[[0, 77, 190]]], 0)
    tfs, _ = O.crop_window_tf(poses, K, 1.2, om["diameter"])
    bbox = O.bbox2d_from_tf(tfs)
    clip = np.stack([O.clip_vertices(om["vertices"], poses[i], bbox[i], K, 480, 640) for i in range(len(poses))])
    pts = np.stack([O.transform_points(om["vertices"], poses[i]) for i in range(len(poses))])
    patch = depth[180:308, 260:388].copy()  # 128x128 window around the object incl. noise and dropouts
    return {"clip": clip.astype(np.float32), "faces": om["faces"], "pts_cam": pts.astype(np.float32),
            "texcoords": om["texcoords"], "texture": om["texture"], "depth_patch": patch,
            "K": np.asarray(K, np.float32)}


def clip_inputs():
    """Hypotheses so close to the camera that the mesh crosses the near plane (z = 0.1 m) or reaches behind
    the camera (w < 0): every such triangle takes cudaraster's polygon clipper (TriangleSetup.inl:303-390)."""
    mesh = synth.uv_ellipsoid("g", slices=24, stacks=13, tex_size=64, seed=5)
    _, _, _, gt = synth.make_scene(mesh)
    om = helpers.oracle_mesh(mesh)
    K = synth.DEFAULT_K
    poses = []
    for tx, ty, tz in [(0.01, -0.01, 0.14), (-0.02, 0.01, 0.11), (0.0, 0.0, 0.04)]:
        p = gt.copy()
        p[:3, 3] = (tx, ty, tz)
        poses.append(p)
    poses = np.stack(poses).astype(np.float32)
    tfs, _ = O.crop_window_tf(poses, K, 1.2, om["diameter"])
    bbox = O.bbox2d_from_tf(tfs)
    clip = np.stack([O.clip_vertices(om["vertices"], poses[i], bbox[i], K, 480, 640) for i in range(len(poses))])
    return {"clip_near": clip.astype(np.float32), "faces": om["faces"], "poses_near": poses}


def main_clip(out_path):
    """tests/golden/ref_clip_golden.npz: cudaraster's visibility ids for clip_inputs()."""
    lib = ref_lib()
    g = clip_inputs()
    n, v, _ = g["clip_near"].shape
    t = g["faces"].shape[0]
    clip_d, tri_d = dev(g["clip_near"]), dev(g["faces"])  # keep the device copies alive across the call
    rast_d = torch.zeros(n, 160, 160, 4, device="cuda")
    assert lib.ref_rasterize(ptr(clip_d), ptr(tri_d), ptr(rast_d), v, t, 160, 160, n) == 0
    ids = rast_d[..., 3].cpu().numpy().astype(np.int16)
    np.savez_compressed(out_path, **g, ref_ids_near=ids)
    print("wrote", out_path, os.path.getsize(out_path), "bytes; coverage", float((ids > 0).mean()))


def main(out_path):
    lib = ref_lib()
    g = golden_inputs()
    n, v, _ = g["clip"].shape
    t = g["faces"].shape[0]
    H = W = 160
    clip_d, tri_d = dev(g["clip"]), dev(g["faces"])
    rast_d = torch.zeros(n, H, W, 4, device="cuda")
    assert lib.ref_rasterize(ptr(clip_d), ptr(tri_d), ptr(rast_d), v, t, H, W, n) == 0
    xyz_d = torch.zeros(n, H, W, 3, device="cuda")
    # every device copy is bound to a name for the duration of the call: a temporary would be released as soon
    # as its pointer has been taken, and a second temporary in the same expression may then reuse its memory
    pts_d = dev(g["pts_cam"])
    assert lib.ref_interpolate(ptr(pts_d), ptr(rast_d), ptr(tri_d), ptr(xyz_d), v, t, 3, 3, H, W, n) == 0
    uv_d = torch.zeros(n, H, W, 2, device="cuda")
    tc_d = dev(g["texcoords"])
    assert lib.ref_interpolate(ptr(tc_d), ptr(rast_d), ptr(tri_d), ptr(uv_d), v, t, 2, 2, H, W, n) == 0
    col_d = torch.zeros(n, H, W, 3, device="cuda")
    th, tw, _ = g["texture"].shape
    tex_d = dev(g["texture"])
    assert lib.ref_texture(ptr(tex_d), ptr(uv_d), ptr(col_d), th, tw, H, W, n) == 0
    dp = dev(g["depth_patch"])
    er, fl = torch.zeros_like(dp), torch.zeros_like(dp)
    assert lib.ref_depth_filters(ptr(dp), ptr(er), ptr(fl), dp.shape[0], dp.shape[1]) == 0
    xyzm = torch.zeros(*dp.shape, 3, device="cuda")
    K = g["K"]
    lib.ref_depth_to_xyz.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                     C.c_float, C.c_float]
    assert lib.ref_depth_to_xyz(ptr(dp), ptr(xyzm), dp.shape[0], dp.shape[1], float(K[0, 0]), float(K[1, 1]),
                                float(K[0, 2]), float(K[1, 2]), 0.001) == 0
    scores = np.array([0.1, 0.9, -0.3, 0.9, 0.5], np.float32)
    scores_d = dev(scores)
    best = lib.ref_argmax(ptr(scores_d), len(scores))
    rast = rast_d.cpu().numpy()
    sub = (slice(None), slice(0, H, 4), slice(0, W, 4))
    np.savez_compressed(
        out_path, **g,
        ref_ids=rast[..., 3].astype(np.int16), ref_bary_sub=rast[sub][..., :3].astype(np.float32),
        ref_xyz_sub=xyz_d.cpu().numpy()[sub], ref_uv_sub=uv_d.cpu().numpy()[sub], ref_col_sub=col_d.cpu().numpy()[sub],
        ref_eroded=er.cpu().numpy(), ref_filtered=fl.cpu().numpy(),
        ref_xyz_map=np.where(g["depth_patch"][..., None] < 0.001, 0, xyzm.cpu().numpy()).astype(np.float32),
        scores=scores, ref_argmax=np.int32(best))
    print("wrote", out_path, os.path.getsize(out_path), "bytes; coverage", float((rast[..., 3] > 0).mean()))


def render_inputs():
    """Inputs of the reference's render sequence (ref_render_sequence in oracle/ref_render_shim.cu): 4 hypotheses
    of the small golden mesh, with the host-side matrices the reference computes in
    GeneratePoseClipOnCUDA / ConstructBBox2D (foundationpose_render.cpp:123-186,570-609) taken from the oracle."""
    mesh = synth.uv_ellipsoid("g", slices=24, stacks=13, tex_size=64, seed=5)
    rgb, depth, mask, gt = synth.make_scene(mesh)
    om = helpers.oracle_mesh(mesh)
    K = synth.DEFAULT_K
    poses = np.concatenate([gt[None], O.sample_hypotheses(depth, mask, K)[[0, 77, 190]]], 0).astype(np.float32)
    tfs, _ = O.crop_window_tf(poses, K, 1.2, om["diameter"])
    bbox = O.bbox2d_from_tf(tfs)
    proj = O.projection_matrix(K, 480, 640)
    mvp = np.stack([O._mat4_mul(proj, O._mat4_mul(O.GL_CAM_IN_CV_CAM, p)) for p in poses]).astype(np.float32)
    return {"poses": poses, "mvp": mvp, "bbox2d": bbox.astype(np.float32), "vertices": om["vertices"],
            "normals": om["normals"], "texcoords": om["texcoords"], "faces": om["faces"], "texture": om["texture"],
            "diameter": np.float32(om["diameter"]), "K": np.asarray(K, np.float32)}


def run_render_sequence(g, repeat=0):
    """Runs the reference's render path on the GPU; returns its buffers as numpy arrays (+ avg ms if repeat > 0)."""
    lib = ref_lib()
    n, v, t = g["poses"].shape[0], g["vertices"].shape[0], g["faces"].shape[0]
    th, tw, _ = g["texture"].shape
    H = W = 160
    col_major = lambda m: np.ascontiguousarray(m.transpose(0, 2, 1)).reshape(n, 16)
    d = {k: dev(a) for k, a in (("pose", col_major(g["poses"])), ("mvp", col_major(g["mvp"])), ("bbox", g["bbox2d"]),
                                ("verts", g["vertices"]), ("normals", g["normals"]), ("tc", g["texcoords"]),
                                ("faces", g["faces"]), ("tex", g["texture"]))}
    out = {"pts_cam": torch.zeros(n, v, 3, device="cuda"), "pose_clip": torch.zeros(n, v, 4, device="cuda"),
           "diffuse": torch.zeros(n, v, device="cuda"), "rast": torch.zeros(n, H, W, 4, device="cuda"),
           "xyz_map": torch.zeros(n, H, W, 3, device="cuda"), "uv": torch.zeros(n, H, W, 2, device="cuda"),
           "color": torch.zeros(n, H, W, 3, device="cuda"), "diffuse_map": torch.zeros(n, H, W, device="cuda"),
           "render_input": torch.zeros(n, H, W, 6, device="cuda")}
    ms = C.c_float(0)
    lib.ref_render_sequence.argtypes = [C.c_void_p] * 17 + [C.c_int] * 9 + [C.c_float, C.c_int, C.c_void_p]
    rc = lib.ref_render_sequence(ptr(d["pose"]), ptr(d["mvp"]), ptr(d["bbox"]), ptr(d["verts"]), ptr(d["normals"]),
                                 ptr(d["tc"]), ptr(d["faces"]), ptr(d["tex"]), ptr(out["pts_cam"]), ptr(out["pose_clip"]),
                                 ptr(out["diffuse"]), ptr(out["rast"]), ptr(out["xyz_map"]), ptr(out["uv"]),
                                 ptr(out["color"]), ptr(out["diffuse_map"]), ptr(out["render_input"]), v, t, th, tw, n,
                                 H, W, 480, 640, float(g["diameter"]), repeat, C.byref(ms))
    assert rc == 0, rc
    res = {k: a.cpu().numpy() for k, a in out.items()}
    res["ms"] = ms.value
    return res


def main_render(out_path):
    """tests/golden/ref_render_golden.npz: outputs of the reference's own foundationpose_render.cu kernels
    (K4 transform_points, K5 generate_pose_clip, K6 transform_normals, K14 refine_color, K15 clamp,
    K16 threshold_and_downscale_pointcloud, K17 concat) chained as in NvdiffrastRender + RenderProcess."""
    g = render_inputs()
    r = run_render_sequence(g)
    sub = (slice(None), slice(0, 160, 4), slice(0, 160, 4))
    np.savez_compressed(out_path, **g, ref_pts_cam=r["pts_cam"], ref_pose_clip=r["pose_clip"], ref_diffuse=r["diffuse"],
                        ref_ids=r["rast"][..., 3].astype(np.int16), ref_color_sub=r["color"][sub],
                        ref_render_input_sub=r["render_input"][sub])
    print("wrote", out_path, os.path.getsize(out_path), "bytes; coverage", float((r["rast"][..., 3] > 0).mean()))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "clip":
        main_clip(sys.argv[2])
    elif len(sys.argv) > 2 and sys.argv[1] == "render":
        main_render(sys.argv[2])
    else:
        main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "ref_kernels_golden.npz"))
