"""GPU box only: the CPU oracle and the product kernels against the REFERENCE's own CUDA kernels
(oracle/_ref/libref_kernels.so = vendored nvdiffrast + the reference's depth kernels, compiled in
place from /root/reference by oracle/Makefile; the prebuilt .so travels with the snapshot).
Bar: visibility ids identical (integer work); interpolated xyz / uv within 2e-6; barycentrics and
texture colours within 2e-4 (the reference is compiled with nvcc's default FMA contraction, the
oracle rounds every operation; the edge-function cancellation amplifies that 1-ulp difference)."""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_kernels.so")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built")]


def _mk():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_oracle_matches_reference_kernels(tmp_path):
    from oracle import fp_oracle as O
    mk = _mk()
    out = str(tmp_path / "g.npz")
    mk.main(out)
    g = np.load(out)
    n = g["clip"].shape[0]
    total_diff = 0
    for i in range(n):
        ids, nclip = O.rasterize_ids(g["clip"][i], g["faces"])
        assert nclip == 0
        total_diff += int((ids != g["ref_ids"][i]).sum())
        rast = O.barycentrics(g["clip"][i], g["faces"], ids)
        same = ids == g["ref_ids"][i]
        sub = (slice(0, 160, 4), slice(0, 160, 4))
        m = same[sub]
        np.testing.assert_allclose(rast[sub][..., :3][m], g["ref_bary_sub"][i][m], atol=2e-4)
        xyz = O.interpolate(g["pts_cam"][i], g["faces"], rast)
        np.testing.assert_allclose(xyz[sub][m], g["ref_xyz_sub"][i][m], atol=2e-6)
        uv = O.interpolate(g["texcoords"], g["faces"], rast)
        np.testing.assert_allclose(uv[sub][m], g["ref_uv_sub"][i][m], atol=2e-6)
        col = O.texture_linear(g["texture"], uv)
        # background pixels sample texel (0,0) in the reference as well (uv = 0)
        np.testing.assert_allclose(col[sub][m], g["ref_col_sub"][i][m], atol=2e-4)
    assert total_diff == 0, f"{total_diff} visibility ids differ from cudaraster over {n} images"
    assert (g["ref_ids"] > 0).mean() > 0.1
    np.testing.assert_array_equal(O.erode_depth(g["depth_patch"]), g["ref_eroded"])
    np.testing.assert_allclose(O.bilateral_filter_depth(g["ref_eroded"]), g["ref_filtered"], rtol=2e-6, atol=1e-7)
    np.testing.assert_array_equal(O.depth_to_xyz(g["depth_patch"], g["K"]), g["ref_xyz_map"])
    assert O.argmax_first(g["scores"]) == int(g["ref_argmax"])


def test_product_raster_matches_reference_cudaraster(tmp_path):
    """The product's fused render kernel against cudaraster on a full-size mesh at 12 hypotheses."""
    import fp_helpers as helpers
    from foundationpose_cpp_b200 import api, synth
    from oracle import fp_oracle as O
    mk = _mk()
    lib = mk.ref_lib()
    mesh = synth.uv_ellipsoid("m", slices=100, stacks=51, tex_size=128, seed=1)  # 10 000 triangles
    rgb, depth, mask, gt = synth.make_scene(mesh)
    rp, sp = helpers.write_seeded_weights(str(tmp_path))
    model = api.CreateFoundationPoseModel(rp, sp, [api.ArrayMeshLoader.from_mesh_data(mesh)], synth.DEFAULT_K, 480, 640)
    model.debug_options(rast_ids=True)
    model.upload(rgb, depth)
    poses = np.concatenate([gt[None], O.sample_hypotheses(depth, mask, synth.DEFAULT_K)[::23]], 0)
    n = len(poses)
    model.set_hypotheses(poses, "m")
    model.render_and_transform(1.2)
    got = model.read_buffer("rast_id", np.int32, n * 160 * 160).reshape(n, 160, 160)
    om = helpers.oracle_mesh(mesh)
    tfs, _ = O.crop_window_tf(poses, synth.DEFAULT_K, 1.2, om["diameter"])
    bbox = O.bbox2d_from_tf(tfs)
    clip = np.stack([O.clip_vertices(om["vertices"], poses[i], bbox[i], synth.DEFAULT_K, 480, 640) for i in range(n)])
    clip_d, tri_d = mk.dev(clip.astype(np.float32)), mk.dev(om["faces"])
    rast_d = torch.zeros(n, 160, 160, 4, device="cuda")
    assert lib.ref_rasterize(mk.ptr(clip_d), mk.ptr(tri_d), mk.ptr(rast_d), clip.shape[1], om["faces"].shape[0], 160,
                             160, n) == 0
    ref_ids = rast_d[..., 3].cpu().numpy().astype(np.int32)
    diff = int((got != ref_ids).sum())
    assert diff == 0, f"{diff} of {got.size} visibility ids differ from the reference cudaraster"
    model.close()


def test_polygon_clipper_matches_reference_cudaraster(tmp_path):
    """Hypotheses whose mesh crosses the near plane / reaches behind the camera: the oracle's and the product's
    restatement of cudaraster's polygon clipper against cudaraster itself.  Bar: identical visibility ids."""
    import fp_helpers as helpers
    from foundationpose_cpp_b200 import api, synth
    from oracle import fp_oracle as O
    mk = _mk()
    out = str(tmp_path / "c.npz")
    mk.main_clip(out)
    g = np.load(out)
    n = g["clip_near"].shape[0]
    ref = g["ref_ids_near"].astype(np.int32)
    assert (ref > 0).mean() > 0.2
    took = 0
    for i in range(n):
        ids, nclip = O.rasterize_ids(g["clip_near"][i], g["faces"])
        took += nclip
        d = int((ids != ref[i]).sum())
        assert d == 0, f"oracle: image {i}: {d} visibility ids differ from cudaraster ({nclip} clipped triangles)"
    assert took > 50, "the scene must exercise the clipper"
    # product kernels (both render paths) on the same hypotheses
    mesh = synth.uv_ellipsoid("g", slices=24, stacks=13, tex_size=64, seed=5)
    rgb, depth, mask, gt = synth.make_scene(mesh)
    rp, sp = helpers.write_seeded_weights(str(tmp_path))
    model = api.CreateFoundationPoseModel(rp, sp, [api.ArrayMeshLoader.from_mesh_data(mesh)], synth.DEFAULT_K, 480, 640)
    model.upload(rgb, depth)
    for split in (0, 1):
        model.debug_options(rast_ids=True, render_split=split)
        model.set_hypotheses(g["poses_near"], "g")
        model.render_and_transform(1.2)
        got = model.read_buffer("rast_id", np.int32, n * 160 * 160).reshape(n, 160, 160)
        d = int((got != ref).sum())
        assert d == 0, f"product (split={split}): {d} visibility ids differ from cudaraster"
        assert model.read_buffer("status", np.int32, 1)[0] & 4, "the clip-path flag must be raised"
    model.close()
