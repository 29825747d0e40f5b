"""CPU: the oracle against golden vectors produced by the REFERENCE's own CUDA kernels
(tests/golden/ref_kernels_golden.npz, generated on a B200 by tests/golden/make_golden.py from
oracle/_ref = vendored nvdiffrast/cudaraster + the reference's depth kernels compiled in place).
This is what pins the oracle: the reference's test-suite itself holds no vectors for this path."""
import os

import numpy as np

from oracle import fp_oracle as O

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_kernels_golden.npz"))
SUB = (slice(0, 160, 4), slice(0, 160, 4))


def test_visibility_ids_identical_to_cudaraster():
    for i in range(G["clip"].shape[0]):
        ids, nclip = O.rasterize_ids(G["clip"][i], G["faces"])
        assert nclip == 0
        assert np.array_equal(ids, G["ref_ids"][i].astype(np.int32)), f"image {i}"
    assert (G["ref_ids"] > 0).mean() > 0.2


def test_polygon_clipper_identical_to_cudaraster():
    """tests/golden/ref_clip_golden.npz (make_golden.py clip): hypotheses so close that the mesh crosses the
    near plane / reaches behind the camera, rasterised by the reference's cudaraster on a B200.  The oracle's
    restatement of its polygon clipper (TriangleSetup.inl:303-390, Util.inl:101-160) must give the same ids."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_clip_golden.npz"))
    took = 0
    for i in range(g["clip_near"].shape[0]):
        ids, nclip = O.rasterize_ids(g["clip_near"][i], g["faces"])
        took += nclip
        assert np.array_equal(ids, g["ref_ids_near"][i].astype(np.int32)), f"image {i}"
    assert took > 50 and (g["ref_ids_near"] > 0).mean() > 0.2


def test_shader_interpolate_texture_match_nvdiffrast():
    for i in range(G["clip"].shape[0]):
        ids = G["ref_ids"][i].astype(np.int32)
        rast = O.barycentrics(G["clip"][i], G["faces"], ids)
        np.testing.assert_allclose(rast[SUB][..., :3], G["ref_bary_sub"][i], atol=2e-4)
        xyz = O.interpolate(G["pts_cam"][i], G["faces"], rast)
        np.testing.assert_allclose(xyz[SUB], G["ref_xyz_sub"][i], atol=2e-6)
        uv = O.interpolate(G["texcoords"], G["faces"], rast)
        np.testing.assert_allclose(uv[SUB], G["ref_uv_sub"][i], atol=2e-6)
        col = O.texture_linear(G["texture"], uv)
        np.testing.assert_allclose(col[SUB], G["ref_col_sub"][i], atol=2e-4)


def test_vertex_stage_and_shading_match_reference_render_kernels():
    """tests/golden/ref_render_golden.npz (make_golden.py render): outputs of the reference's own
    foundationpose_render.cu kernels K4 / K5 / K6 (vertex stage) and K14-K17 (shading, clamp, normalise, concat)
    chained as in NvdiffrastRender, generated on a B200.  The oracle restates each of them
    (detection_6d_foundationpose/src/foundationpose_render.cu:61-140,321-501)."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_render_golden.npz")
    g = np.load(path)  # committed; regenerate with `python tests/golden/make_golden.py render <out>` on a GPU box
    eps = np.finfo(np.float32).eps
    mesh = {"vertices": g["vertices"], "normals": g["normals"], "texcoords": g["texcoords"], "faces": g["faces"],
            "texture": g["texture"], "diameter": float(g["diameter"])}
    for i in range(g["poses"].shape[0]):
        pts = O.transform_points(g["vertices"], g["poses"][i])
        clip = O.clip_vertices(g["vertices"], g["poses"][i], g["bbox2d"][i], g["K"], 480, 640)
        dif = O.diffuse_intensity(g["normals"], g["poses"][i])
        # the reference build contracts the row sums into FMAs; the oracle rounds every product: <= a few ulp
        assert np.abs(pts - g["ref_pts_cam"][i]).max() <= 4 * eps * max(1.0, np.abs(g["ref_pts_cam"][i]).max())
        assert np.abs(clip - g["ref_pose_clip"][i]).max() <= 4 * eps * max(1.0, np.abs(g["ref_pose_clip"][i]).max())
        assert np.abs(dif - g["ref_diffuse"][i]).max() <= 4 * eps
        ids, _ = O.rasterize_ids(clip, g["faces"])
        assert np.array_equal(ids, g["ref_ids"][i].astype(np.int32))
        tfs, _ = O.crop_window_tf(g["poses"][i:i + 1], g["K"], 1.2, mesh["diameter"])
        col, xyz, _, _ = O.render_crop(mesh, g["poses"][i], tfs[0], g["K"], 480, 640)
        packed = O.normalize_pack(col, xyz, g["poses"][i], mesh["diameter"])
        # reference rows are in GL order (its two cvcuda::Flip calls are CV-CUDA, absent here): un-flip the oracle
        np.testing.assert_allclose(packed[::-1][SUB], g["ref_render_input_sub"][i], atol=2e-4)


def test_depth_kernels_match_reference():
    np.testing.assert_array_equal(O.erode_depth(G["depth_patch"]), G["ref_eroded"])
    np.testing.assert_allclose(O.bilateral_filter_depth(G["ref_eroded"]), G["ref_filtered"], rtol=2e-6, atol=1e-7)
    np.testing.assert_array_equal(O.depth_to_xyz(G["depth_patch"], G["K"]), G["ref_xyz_map"])
    assert (G["ref_eroded"] == 0).sum() > (G["depth_patch"] == 0).sum()  # the patch exercises the erosion


def test_argmax_matches_thrust_max_element():
    assert O.argmax_first(G["scores"]) == int(G["ref_argmax"]) == 1  # first of the two equal maxima


def test_network_oracle_reproduces_committed_outputs():
    """tests/golden/nets_golden.npz (make_nets_golden.py): the fp32 PyTorch oracle of the refine / score
    networks on seeded weights and inputs.  Pins the oracle against drift; the CUDA path is checked
    against the same file in test_pipeline_gpu.py."""
    import sys
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import fp_helpers as helpers
    import make_nets_golden as mk
    from oracle import nets
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nets_golden.npz"))
    a, b = mk.golden_inputs()
    assert abs(float(a.sum() + 2 * b.sum()) - float(g["input_checksum"])) < 1e-3 * abs(float(g["input_checksum"]))
    refiner, scorer = helpers.seeded_nets()
    n = 2  # two of the three hypotheses keep the CPU suite short; the scorer needs all of them (cross-attention)
    with torch.no_grad():
        trans, rot = refiner(nets.nhwc_to_inputs(torch.from_numpy(a[:n])), nets.nhwc_to_inputs(torch.from_numpy(b[:n])))
        scores = scorer(nets.nhwc_to_inputs(torch.from_numpy(a)), nets.nhwc_to_inputs(torch.from_numpy(b)))
    np.testing.assert_allclose(trans.numpy(), g["trans"][:n], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(rot.numpy(), g["rot"][:n], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(scores.numpy().reshape(-1), g["scores"], rtol=1e-3, atol=1e-5)
