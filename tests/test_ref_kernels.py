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


def test_render_sequence_matches_reference_render_kernels(tmp_path):
    """The reference's own foundationpose_render.cu kernels (K4-K6, K14-K17; compiled in place into oracle/_ref)
    chained as in NvdiffrastRender + RenderProcess, against (a) the oracle's restatement of each of them and
    (b) the product's fused render kernel, on the same hypotheses.  Bars: vertex stage bit-exact (the same fp32
    expressions; nvcc contracts a*x + b*y + c*z + d into FMAs in the reference build, so <= 2 ulp is allowed and the
    maximum is recorded), ids identical, packed render_input within fp16 rounding of the reference's fp32 output."""
    import fp_helpers as helpers
    from conftest import record_err
    from foundationpose_cpp_b200 import api, synth
    from oracle import fp_oracle as O
    mk = _mk()
    g = mk.render_inputs()
    r = mk.run_render_sequence(g)
    n = g["poses"].shape[0]
    # (a) oracle restatement of the vertex-stage kernels
    for i in range(n):
        pts = O.transform_points(g["vertices"], g["poses"][i])
        clip = O.clip_vertices(g["vertices"], g["poses"][i], g["bbox2d"][i], g["K"], 480, 640)
        dif = O.diffuse_intensity(g["normals"], g["poses"][i])
        for name, got, ref in (("K4 transform_points", pts, r["pts_cam"][i]), ("K5 generate_pose_clip", clip, r["pose_clip"][i]),
                               ("K6 transform_normals", dif, r["diffuse"][i])):
            err = float(np.abs(got - ref).max())
            tol = 4 * np.finfo(np.float32).eps * max(1.0, float(np.abs(ref).max()))
            record_err(f"ref render kernels/oracle {name} [{i}]", err, tol)
            assert err <= tol, (name, err)
    # (b) the product on the same hypotheses
    mesh = synth.uv_ellipsoid("g", slices=24, stacks=13, tex_size=64, seed=5)
    rgb, depth, mask, gt = synth.make_scene(mesh)
    rp, sp = helpers.write_seeded_weights(str(tmp_path))
    model = api.CreateFoundationPoseModel(rp, sp, [api.ArrayMeshLoader.from_mesh_data(mesh)], synth.DEFAULT_K, 480, 640)
    model.debug_options(rast_ids=True)
    model.upload(rgb, depth)
    model.set_hypotheses(g["poses"], "g")
    model.render_and_transform(1.2)
    ids = model.read_buffer("rast_id", np.int32, n * 160 * 160).reshape(n, 160, 160)
    ref_ids = r["rast"][..., 3].astype(np.int32)
    assert int((ids != ref_ids).sum()) == 0
    got = model.read_packed("render_input", n)
    ref = r["render_input"][:, ::-1]  # the two cvcuda::Flip calls of the reference (CV-CUDA is absent)
    ref16 = ref.astype(np.float16).astype(np.float32)
    diff = np.abs(got - ref16)
    ulp = np.maximum(np.abs(ref16), 2.0 ** -14) * 2.0 ** -10
    # the reference build contracts fp32 expressions into FMAs, the product (like the oracle) rounds every operation:
    # barycentrics / texture colours then differ by up to ~4e-4 (edge-function cancellation), i.e. up to 2 fp16 steps
    record_err("ref render kernels/product render_input vs reference fp32 output (fp16 ulps)", float((diff / ulp).max()), 2.5)
    record_err("ref render kernels/product render_input: fraction of elements that differ", float((diff > 0).mean()), 1e-2)
    assert np.all(diff <= 2.5 * ulp), float((diff / ulp).max())
    assert float((diff > 0).mean()) <= 1e-2
    assert float((diff > 1.001 * ulp).mean()) <= 1e-4
    # the oracle's full render (what every other test compares the product with) against the same reference output
    xyz_map = O.depth_to_xyz(depth, synth.DEFAULT_K)
    ren, _, _ = O.render_and_transform(helpers.oracle_mesh(mesh), g["poses"], rgb, xyz_map, synth.DEFAULT_K, 1.2)
    err = float(np.abs(ren - ref).max())
    record_err("ref render kernels/oracle render_input vs reference", err, 2e-4)
    assert err <= 2e-4
    model.close()
