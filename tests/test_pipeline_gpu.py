"""GPU parity of the hot path through the C ABI against the CPU oracle (oracle/), same seeded inputs.

Tolerances (stated per test):
  * integer work (visibility ids, crop-window rounding) ............ bit-exact
  * packed network inputs (fp16 storage of fp32 arithmetic) ........ equal to the oracle value rounded
    to fp16, except <= 0.1 % of elements that may differ by one fp16 ulp (fp32 contraction in numpy vs CUDA)
  * network outputs (fp16 tensor-core path vs fp32 PyTorch) ......... |err| <= 3e-2 * max|ref| + 3e-3
  * refined pose ................................................... rotation <= 0.2 deg, translation <= 0.5 mm
"""
import math

import os

import numpy as np
import pytest
import torch

from foundationpose_cpp_b200 import api, synth
from oracle import fp_oracle as O
from oracle import nets
import fp_helpers as helpers

pytestmark = pytest.mark.gpu
F = np.float32


@pytest.fixture(scope="module")
def env(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("weights"))
    rp, sp = helpers.write_seeded_weights(d)
    mesh = synth.uv_ellipsoid("mustard_proxy", slices=64, stacks=33, tex_size=256, seed=0)
    rgb, depth, mask, gt = synth.make_scene(mesh)
    model = api.CreateFoundationPoseModel(rp, sp, [api.ArrayMeshLoader.from_mesh_data(mesh)], synth.DEFAULT_K,
                                          max_h=480, max_w=640)
    yield {"model": model, "mesh": mesh, "omesh": helpers.oracle_mesh(mesh), "rgb": rgb, "depth": depth,
           "mask": mask, "gt": gt, "K": synth.DEFAULT_K}
    model.close()


def test_depth_to_xyz_exact(env):
    m = env["model"]
    m.upload(env["rgb"], env["depth"])
    got = m.read_buffer("xyz_map", np.float32, 480 * 640 * 3).reshape(480, 640, 3)
    ref = O.depth_to_xyz(env["depth"], env["K"])
    assert np.array_equal(got, ref)


def test_sampler_matches_oracle(env):
    m = env["model"]
    m.upload(env["rgb"], env["depth"])
    m.sample_hypotheses(env["mask"], "mustard_proxy")
    got = m.get_hypotheses()
    ref = O.sample_hypotheses(env["depth"], env["mask"], env["K"])
    assert got.shape == (252, 4, 4)
    np.testing.assert_allclose(got[:, :3, :3], ref[:, :3, :3], atol=2e-6)
    np.testing.assert_allclose(got[:, :3, 3], ref[:, :3, 3], rtol=2e-6, atol=1e-7)
    filt = m.read_buffer("depth_filtered", np.float32, 480 * 640).reshape(480, 640)
    np.testing.assert_allclose(filt, O.bilateral_filter_depth(O.erode_depth(env["depth"])), rtol=1e-5, atol=1e-6)


def test_sampler_failure_modes(env):
    m = env["model"]
    m.upload(env["rgb"], env["depth"])
    with pytest.raises(Exception) as e:
        m.sample_hypotheses(np.zeros_like(env["mask"]), "mustard_proxy")
    assert "Mask is all zero" in str(e.value)
    m.upload(env["rgb"], np.zeros_like(env["depth"]))
    with pytest.raises(Exception) as e:
        m.sample_hypotheses(env["mask"], "mustard_proxy")
    assert "No valid value" in str(e.value)
    with pytest.raises(Exception):
        m.sample_hypotheses(env["mask"], "no_such_object")


def _fp16_close(got, ref, what):
    ref16 = ref.astype(np.float16).astype(np.float32)
    diff = np.abs(got - ref16)
    ulp = np.maximum(np.abs(ref16), 2.0 ** -14) * 2.0 ** -10
    assert np.all(diff <= 1.001 * ulp), f"{what}: max diff {diff.max()} beyond one fp16 ulp"
    frac = float((diff > 0).mean())
    assert frac <= 1e-3, f"{what}: {frac:.2e} of elements differ from the fp16-rounded oracle"


@pytest.mark.parametrize("split", [0, 1])
@pytest.mark.parametrize("crop_ratio", [1.2, 1.1])
def test_render_and_transform_parity(env, crop_ratio, split):
    # split = "1": the few-hypotheses path (triangles spread over all SMs on global z-buffers, then a shading
    # kernel: what Track uses); "0": one CTA per (band, hypothesis) on a shared-memory z-buffer (Register)
    m = env["model"]
    m.debug_options(rast_ids=True, render_split=split)
    m.upload(env["rgb"], env["depth"])
    poses = O.sample_hypotheses(env["depth"], env["mask"], env["K"])[[0, 7, 50, 101, 200, 251]]
    poses = np.concatenate([poses, env["gt"][None]], 0)
    m.set_hypotheses(poses, "mustard_proxy")
    m.render_and_transform(crop_ratio)
    n = len(poses)
    xyz_map = O.depth_to_xyz(env["depth"], env["K"])
    ren, tra, info = O.render_and_transform(env["omesh"], poses, env["rgb"], xyz_map, env["K"], crop_ratio)
    assert info["clip_path"] == 0
    ids = m.read_buffer("rast_id", np.int32, n * 160 * 160).reshape(n, 160, 160)
    assert np.array_equal(ids, info["ids"]), f"{int((ids != info['ids']).sum())} visibility ids differ"
    assert (ids > 0).mean() > 0.1
    _fp16_close(m.read_packed("transf_input", n), tra, "transf_input")
    _fp16_close(m.read_packed("render_input", n), ren, "render_input")
    status = m.read_buffer("status", np.int32, 1)[0]
    assert status == 0
    # the crop transform / bbox2d the C ABI serves == ComputeCropWindowTF / ConstructBBox2D of the oracle
    np.testing.assert_array_equal(m.read_buffer("crop_tf", np.float32, n * 9).reshape(n, 3, 3), info["tf"].astype(F))
    np.testing.assert_array_equal(m.read_buffer("bbox2d", np.float32, n * 4).reshape(n, 4),
                                  O.bbox2d_from_tf(info["tf"]).astype(F))
    m.debug_options()  # back to the product defaults (no id writes, automatic render path)


def _net_close(got, ref, what):
    tol = 3e-2 * np.abs(ref).max() + 3e-3
    err = np.abs(got - ref).max()
    assert err <= tol, f"{what}: max err {err:.4g} > tol {tol:.4g}"


def _rot_err_deg(a, b):
    c = (np.trace(a[:3, :3].astype(np.float64) @ b[:3, :3].astype(np.float64).T) - 1) / 2
    return math.degrees(math.acos(min(1.0, max(-1.0, c))))


def test_refiner_and_pose_update_parity(env):
    m = env["model"]
    refiner, _ = helpers.seeded_nets()
    m.upload(env["rgb"], env["depth"])
    poses = O.sample_hypotheses(env["depth"], env["mask"], env["K"])[:5]
    m.set_hypotheses(poses, "mustard_proxy")
    m.refine_step()
    n = len(poses)
    a = torch.from_numpy(m.read_packed("render_input", n))
    b = torch.from_numpy(m.read_packed("transf_input", n))
    with torch.no_grad():
        t_ref, r_ref = refiner(nets.nhwc_to_inputs(a), nets.nhwc_to_inputs(b))
    _net_close(m.read_buffer("trans", np.float32, n * 3).reshape(n, 3), t_ref.numpy(), "trans")
    _net_close(m.read_buffer("rot", np.float32, n * 3).reshape(n, 3), r_ref.numpy(), "rot")
    got = m.get_hypotheses()
    ref = O.pose_update(poses, t_ref.numpy(), r_ref.numpy(), env["mesh"].diameter)
    for i in range(n):
        assert _rot_err_deg(got[i], ref[i]) <= 0.2
        assert np.abs(got[i][:3, 3] - ref[i][:3, 3]).max() <= 5e-4


def test_refiner_with_second_attention_kernel(env, tmp_path, monkeypatch):
    """attention2.cu (FPB200_ATTN_V2=1, read when a session builds its plans) as a drop-in inside the refiner: same
    outputs as the fp32 modules (same tolerance as the default kernel) on the same hypotheses."""
    monkeypatch.setenv("FPB200_ATTN_V2", "1")
    rp, sp = helpers.write_seeded_weights(str(tmp_path))
    m = api.CreateFoundationPoseModel(rp, sp, [api.ArrayMeshLoader.from_mesh_data(env["mesh"])], env["K"], max_h=480,
                                      max_w=640)
    try:
        refiner, _ = helpers.seeded_nets()
        m.upload(env["rgb"], env["depth"])
        poses = O.sample_hypotheses(env["depth"], env["mask"], env["K"])[:5]
        m.set_hypotheses(poses, "mustard_proxy")
        m.refine_step()
        n = len(poses)
        a = torch.from_numpy(m.read_packed("render_input", n))
        b = torch.from_numpy(m.read_packed("transf_input", n))
        with torch.no_grad():
            t_ref, r_ref = refiner(nets.nhwc_to_inputs(a), nets.nhwc_to_inputs(b))
        _net_close(m.read_buffer("trans", np.float32, n * 3).reshape(n, 3), t_ref.numpy(), "trans (attention2)")
        _net_close(m.read_buffer("rot", np.float32, n * 3).reshape(n, 3), r_ref.numpy(), "rot (attention2)")
    finally:
        m.close()


def test_scorer_parity_and_argmax(env):
    m = env["model"]
    _, scorer = helpers.seeded_nets()
    m.upload(env["rgb"], env["depth"])
    poses = O.sample_hypotheses(env["depth"], env["mask"], env["K"])[::42]
    m.set_hypotheses(poses, "mustard_proxy")
    m.score_features()
    n = len(poses)
    a = torch.from_numpy(m.read_packed("render_input", n))
    b = torch.from_numpy(m.read_packed("transf_input", n))
    with torch.no_grad():
        feats = scorer.features(nets.nhwc_to_inputs(a), nets.nhwc_to_inputs(b))
        scores = scorer.tail(feats)
    idx, pose = m.score_finish()
    _net_close(m.read_buffer("features", np.float32, n * 512).reshape(n, 512), feats.numpy(), "features")
    got_scores = m.read_buffer("scores", np.float32, n)
    _net_close(got_scores, scores.numpy(), "scores")
    ref_idx = O.argmax_first(scores.numpy())
    srt = np.sort(scores.numpy())
    assert idx == ref_idx or (srt[-1] - srt[-2]) < 2 * np.abs(got_scores - scores.numpy()).max()
    assert idx == int(np.argmax(got_scores))
    np.testing.assert_allclose(pose, m.get_hypotheses()[idx], atol=0)


def test_register_and_track_end_to_end(env):
    """Config 1 of BASELINE.json at reduced N is covered above; here the public calls themselves:
    Register over all 252 hypotheses returns one of the refined hypotheses, Track refines a pose."""
    m = env["model"]
    ok, pose = m.Register(env["rgb"], env["depth"], env["mask"], "mustard_proxy", refine_itr=1)
    assert ok, m.last_error
    assert pose.shape == (4, 4) and np.isfinite(pose).all()
    r = pose[:3, :3].astype(np.float64)
    np.testing.assert_allclose(r @ r.T, np.eye(3), atol=1e-4)
    assert abs(pose[2, 3] - 0.6) < 0.1  # translation guess + one refine step stays near the object
    assert m.last_launch_count() > 40
    ok2, pose2 = m.Track(env["rgb"], env["depth"], pose, "mustard_proxy", refine_itr=2)
    assert ok2 and np.isfinite(pose2).all()
    ok3, _ = m.Register(env["rgb"], env["depth"], np.zeros_like(env["mask"]), "mustard_proxy")
    assert not ok3 and "Mask is all zero" in m.last_error
    ok4, _ = m.Register(env["rgb"], env["depth"], env["mask"], "unknown_object")
    assert not ok4
    ok5, _ = m.Register(env["rgb"][:100], env["depth"], env["mask"], "mustard_proxy")
    assert not ok5


def test_cuda_graph_replay_matches_eager(env):
    """Register / Track run eagerly on the first call, are captured into a CUDA graph on the second and
    replayed afterwards: all three must return the same pose bit for bit."""
    m = env["model"]
    poses = [m.Register(env["rgb"], env["depth"], env["mask"], "mustard_proxy", refine_itr=2)[1] for _ in range(4)]
    for p in poses[1:]:
        assert np.array_equal(p, poses[0])
    tr = [m.Track(env["rgb"], env["depth"], poses[0], "mustard_proxy", refine_itr=1)[1] for _ in range(4)]
    for p in tr[1:]:
        assert np.array_equal(p, tr[0])
    # a different image size or refine count is a different graph
    ok, p2 = m.Register(env["rgb"], env["depth"], env["mask"], "mustard_proxy", refine_itr=1)
    assert ok and np.isfinite(p2).all()


def test_register_async_matches_register(env):
    """fpb200_register_async (frame already resident, no host sync until wait) == the public Register."""
    m = env["model"]
    ok, ref = m.Register(env["rgb"], env["depth"], env["mask"], "mustard_proxy", refine_itr=1)
    assert ok
    m.upload(env["rgb"], env["depth"])
    m.sample_hypotheses(env["mask"], "mustard_proxy")
    for _ in range(3):  # eager, capture, replay
        m.register_async("mustard_proxy", 1)
        idx, pose = m.wait()
        assert np.array_equal(pose, ref)
        assert 0 <= idx < 252


def test_track_async_matches_track(env):
    """fpb200_track_async + fpb200_track_wait (pipelining several sessions from one thread) == Track."""
    m = env["model"]
    hyp = synth.perturb_pose(env["gt"], 5.0, 0.01, 3)
    ok, ref = m.Track(env["rgb"], env["depth"], hyp, "mustard_proxy", refine_itr=2)
    assert ok
    rgb, depth = np.ascontiguousarray(env["rgb"]), np.ascontiguousarray(env["depth"], np.float32)
    for _ in range(3):
        m.track_async(rgb, depth, hyp, "mustard_proxy", 2)
        assert np.array_equal(m.track_wait(), ref)


def test_networks_match_committed_golden_outputs(tmp_path):
    """CUDA refine / score networks (through the infer-core C ABI) against tests/golden/nets_golden.npz,
    the committed outputs of the fp32 oracle on seeded weights and inputs."""
    import ctypes as C
    import sys
    from foundationpose_cpp_b200 import capi
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_nets_golden as mk
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nets_golden.npz"))
    a, b = mk.golden_inputs()
    n = a.shape[0]
    rp, sp = helpers.write_seeded_weights(str(tmp_path))
    lib = capi.lib()
    ad, bd = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    core = C.c_void_p()
    capi.check(lib.fpb200_infer_create(rp.encode(), 0, 8, 0, C.byref(core)))
    trans, rot = torch.zeros(n, 3, device="cuda"), torch.zeros(n, 3, device="cuda")
    capi.check(lib.fpb200_infer_refiner(core, C.c_void_p(ad.data_ptr()), C.c_void_p(bd.data_ptr()), n,
                                        C.c_void_p(trans.data_ptr()), C.c_void_p(rot.data_ptr())))
    lib.fpb200_infer_destroy(core)
    _net_close(trans.cpu().numpy(), g["trans"], "golden trans")
    _net_close(rot.cpu().numpy(), g["rot"], "golden rot")
    core = C.c_void_p()
    capi.check(lib.fpb200_infer_create(sp.encode(), 1, 8, 0, C.byref(core)))
    scores = torch.zeros(n, device="cuda")
    capi.check(lib.fpb200_infer_scorer(core, C.c_void_p(ad.data_ptr()), C.c_void_p(bd.data_ptr()), n,
                                       C.c_void_p(scores.data_ptr())))
    lib.fpb200_infer_destroy(core)
    np.testing.assert_allclose(scores.cpu().numpy(), g["scores"], atol=5e-3)


def test_infer_core_boundary(tmp_path):
    """The BaseInferCore-shaped entry points on the reference's own blob layout (f32 NHWC [n,160,160,6])."""
    import ctypes as C
    from foundationpose_cpp_b200 import capi
    refiner, scorer = helpers.seeded_nets()
    rp, sp = helpers.write_seeded_weights(str(tmp_path))
    lib = capi.lib()
    n = 4
    torch.manual_seed(5)
    a = torch.rand(n, 160, 160, 6)
    b = torch.rand(n, 160, 160, 6)
    a[..., 3:] = a[..., 3:] * 2 - 1
    b[..., 3:] = b[..., 3:] * 2 - 1
    ad, bd = a.cuda(), b.cuda()
    core = C.c_void_p()
    capi.check(lib.fpb200_infer_create(rp.encode(), 0, 8, 0, C.byref(core)))
    trans = torch.zeros(n, 3, device="cuda")
    rot = torch.zeros(n, 3, device="cuda")
    capi.check(lib.fpb200_infer_refiner(core, C.c_void_p(ad.data_ptr()), C.c_void_p(bd.data_ptr()), n,
                                        C.c_void_p(trans.data_ptr()), C.c_void_p(rot.data_ptr())))
    with torch.no_grad():
        t_ref, r_ref = refiner(nets.nhwc_to_inputs(a.half().float()), nets.nhwc_to_inputs(b.half().float()))
    _net_close(trans.cpu().numpy(), t_ref.numpy(), "infer-core trans")
    _net_close(rot.cpu().numpy(), r_ref.numpy(), "infer-core rot")
    assert lib.fpb200_infer_refiner(core, C.c_void_p(ad.data_ptr()), C.c_void_p(bd.data_ptr()), 9,
                                    C.c_void_p(trans.data_ptr()), C.c_void_p(rot.data_ptr())) == capi.ERR_INVALID
    lib.fpb200_infer_destroy(core)
    core = C.c_void_p()
    capi.check(lib.fpb200_infer_create(sp.encode(), 1, 8, 0, C.byref(core)))
    scores = torch.zeros(n, device="cuda")
    capi.check(lib.fpb200_infer_scorer(core, C.c_void_p(ad.data_ptr()), C.c_void_p(bd.data_ptr()), n,
                                       C.c_void_p(scores.data_ptr())))
    with torch.no_grad():
        s_ref = scorer(nets.nhwc_to_inputs(a.half().float()), nets.nhwc_to_inputs(b.half().float()))
    _net_close(scores.cpu().numpy(), s_ref.numpy(), "infer-core scores")
    lib.fpb200_infer_destroy(core)
    bad = C.c_void_p()
    assert lib.fpb200_infer_create(rp.encode(), 1, 8, 0, C.byref(bad)) != 0  # refiner file given as scorer


def test_mesh_diameter_on_device_matches_all_pairs_maximum(tmp_path):
    """CalcMeshDiameter (assimp_mesh_loader.cpp:47-60, maximum over ALL vertex pairs) on the device.
    Float32 all-pairs maximum in numpy (the reference's arithmetic: differences, squares, sum, sqrt) on 2 210 and on
    5 252 vertices (ragged last tile); a session created with diameter <= 0 computes the same value itself."""
    from foundationpose_cpp_b200 import mesh_loader
    for mesh in (synth.uv_ellipsoid("a", slices=64, stacks=33, tex_size=32, seed=1),
                 synth.uv_ellipsoid("b", slices=100, stacks=51, axes=(0.08, 0.08, 0.1), tex_size=32, seed=2)):
        v = np.ascontiguousarray(mesh.vertices, F)
        best = F(0)
        for i in range(0, len(v), 1024):
            d = v[i:i + 1024, None, :] - v[None, :, :]
            best = max(best, (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2]).max())
        want = float(np.sqrt(F(best)))
        got = api.mesh_diameter_gpu(v)
        assert got == want, (got, want)  # bit-exact: same fp32 operations, max is order-independent
        assert abs(got - mesh_loader.max_pairwise_distance(v)) < 1e-6
    rp, sp = helpers.write_seeded_weights(str(tmp_path))
    ld = api.ArrayMeshLoader.from_mesh_data(mesh)
    ld._d = 0.0  # "not supplied": fpb200_create computes it
    m0 = api.CreateFoundationPoseModel(rp, sp, [ld], synth.DEFAULT_K, max_h=480, max_w=640, max_hyps=4)
    ld._d = got
    m1 = api.CreateFoundationPoseModel(rp, sp, [ld], synth.DEFAULT_K, max_h=480, max_w=640, max_hyps=4)
    rgb, depth, mask, gt = synth.make_scene(mesh)
    try:
        outs = []
        for m in (m0, m1):
            ok, pose = m.Track(rgb, depth, gt, mesh.name, refine_itr=1)
            assert ok, m.last_error
            outs.append(pose)
        assert np.array_equal(outs[0], outs[1])
    finally:
        m0.close()
        m1.close()


def test_two_meshes_in_one_session_match_single_mesh_sessions(tmp_path):
    """one session holding two meshes (map_name2renderer_ of the reference,
    foundationpose.cpp:141-151).  Register of either target - alternating, so the per-mesh CUDA graphs and the shared
    vertex scratch are reused across targets - must give exactly what a session holding only that mesh gives."""
    rp, sp = helpers.write_seeded_weights(str(tmp_path))
    mesh_a = synth.uv_ellipsoid("a", slices=48, stacks=25, tex_size=128, seed=0)
    mesh_b = synth.uv_ellipsoid("b", slices=64, stacks=33, axes=(0.07, 0.05, 0.06), tex_size=64, seed=5)
    scene_a = synth.make_scene(mesh_a)
    scene_b = synth.make_scene(mesh_b, t=(-0.03, 0.02, 0.7), rot_seed=4)
    want = {}
    for mesh, scene in ((mesh_a, scene_a), (mesh_b, scene_b)):
        m1 = api.CreateFoundationPoseModel(rp, sp, [api.ArrayMeshLoader.from_mesh_data(mesh)], synth.DEFAULT_K, max_h=480,
                                           max_w=640, max_hyps=40)
        ok, pose = m1.Register(scene[0], scene[1], scene[2], mesh.name, refine_itr=1)
        assert ok, m1.last_error
        want[mesh.name] = (pose, m1.read_buffer("scores", np.float32, 40).copy())
        m1.close()
    both = api.CreateFoundationPoseModel(rp, sp, [api.ArrayMeshLoader.from_mesh_data(mesh_a),
                                                  api.ArrayMeshLoader.from_mesh_data(mesh_b)], synth.DEFAULT_K,
                                         max_h=480, max_w=640, max_hyps=40)
    try:
        for name, scene in (("a", scene_a), ("b", scene_b), ("a", scene_a), ("b", scene_b), ("b", scene_b), ("a", scene_a)):
            ok, pose = both.Register(scene[0], scene[1], scene[2], name, refine_itr=1)
            assert ok, both.last_error
            assert np.array_equal(pose, want[name][0]), name
            assert np.array_equal(both.read_buffer("scores", np.float32, 40), want[name][1]), name
        ok, _ = both.Register(scene_a[0], scene_a[1], scene_a[2], "c", refine_itr=1)
        assert not ok and "unknown target" in both.last_error
    finally:
        both.close()


def test_skewed_intrinsics_match_oracle(tmp_path):
    """K[0][1] != 0: the crop-window projection and the GL projection matrix use the full K
    (foundationpose_render.cpp:44-75,151-186) while depth -> xyz ignores the skew like the reference kernel
    (foundationpose_utils.cu:3-32).  Sampler, crop transform, visibility ids and packed inputs against the oracle."""
    rp, sp = helpers.write_seeded_weights(str(tmp_path))
    K = synth.DEFAULT_K.copy()
    K[0, 1] = 2.5
    mesh = synth.uv_ellipsoid("obj", slices=48, stacks=25, tex_size=128, seed=2)
    rgb, depth, mask, gt = synth.make_scene(mesh, K=K)
    m = api.CreateFoundationPoseModel(rp, sp, [api.ArrayMeshLoader.from_mesh_data(mesh)], K, max_h=480, max_w=640, max_hyps=252)
    try:
        m.debug_options(rast_ids=True)
        m.upload(rgb, depth)
        m.sample_hypotheses(mask, "obj")
        got = m.get_hypotheses()
        ref = O.sample_hypotheses(depth, mask, K)
        np.testing.assert_allclose(got[:, :3, 3], ref[:, :3, 3], rtol=2e-6, atol=1e-7)
        assert abs(ref[0, 0, 3] - O.sample_hypotheses(depth, mask, synth.DEFAULT_K)[0, 0, 3]) > 1e-6  # the skew matters
        poses = np.concatenate([ref[[3, 77, 160]], gt[None]], 0)
        m.set_hypotheses(poses, "obj")
        m.render_and_transform(1.2)
        n = len(poses)
        ren, tra, info = O.render_and_transform(helpers.oracle_mesh(mesh), poses, rgb, O.depth_to_xyz(depth, K), K, 1.2)
        np.testing.assert_array_equal(m.read_buffer("crop_tf", np.float32, n * 9).reshape(n, 3, 3), info["tf"].astype(F))
        ids = m.read_buffer("rast_id", np.int32, n * 160 * 160).reshape(n, 160, 160)
        assert np.array_equal(ids, info["ids"]), f"{int((ids != info['ids']).sum())} visibility ids differ"
        _fp16_close(m.read_packed("transf_input", n), tra, "transf_input (skewed K)")
        _fp16_close(m.read_packed("render_input", n), ren, "render_input (skewed K)")
    finally:
        m.close()


def test_full_hd_frame_with_default_capacity(tmp_path):
    """1080 x 1920 frames, the reference's default max_input_image_height / width (foundationpose.hpp:99-105): the
    session is created with the defaults; sampler (whole-image depth filters, mask scan, median) and the packed network
    inputs against the oracle, then Register / Track through the public call."""
    rp, sp = helpers.write_seeded_weights(str(tmp_path))
    K = np.array([[1400.0, 0, 960.0], [0, 1400.0, 540.0], [0, 0, 1]], F)
    mesh = synth.uv_ellipsoid("obj", slices=48, stacks=25, tex_size=128, seed=1)
    rgb, depth, mask, gt = synth.make_scene(mesh, K=K, h=1080, w=1920, t=(0.15, -0.08, 0.55))
    m = api.CreateFoundationPoseModel(rp, sp, [api.ArrayMeshLoader.from_mesh_data(mesh)], K, max_hyps=252)
    try:
        m.debug_options(rast_ids=True)
        m.upload(rgb, depth)
        m.sample_hypotheses(mask, "obj")
        ref = O.sample_hypotheses(depth, mask, K)
        np.testing.assert_allclose(m.get_hypotheses()[:, :3, 3], ref[:, :3, 3], rtol=2e-6, atol=1e-7)
        poses = np.concatenate([ref[[10, 200]], gt[None]], 0)
        m.set_hypotheses(poses, "obj")
        m.render_and_transform(1.2)
        ren, tra, info = O.render_and_transform(helpers.oracle_mesh(mesh), poses, rgb, O.depth_to_xyz(depth, K), K, 1.2)
        ids = m.read_buffer("rast_id", np.int32, 3 * 160 * 160).reshape(3, 160, 160)
        assert np.array_equal(ids, info["ids"])
        _fp16_close(m.read_packed("transf_input", 3), tra, "transf_input (1080p)")
        _fp16_close(m.read_packed("render_input", 3), ren, "render_input (1080p)")
        m.debug_options()
        ok, pose = m.Register(rgb, depth, mask, "obj", refine_itr=1)
        assert ok and np.isfinite(pose).all(), m.last_error
        ok, pose2 = m.Track(rgb, depth, gt, "obj", refine_itr=2)
        assert ok and np.isfinite(pose2).all(), m.last_error
        # a frame larger than the capacity is refused, not truncated
        big = np.zeros((1081, 1920, 3), np.uint8)
        ok, _ = m.Track(big, np.zeros((1081, 1920), F), gt, "obj")
        assert not ok
    finally:
        m.close()


@pytest.mark.parametrize("arch", [
    # one residual block less in encodeA, one more in front of the down-sampling layer, one less after it
    ((0, 0, 64, 2), (0, 1, 128, 2), (0, 2, 128, 1), (1, 2, 256, 1), (1, 2, 256, 1), (1, 2, 256, 1), (1, 1, 512, 2), (1, 2, 512, 1)),
    # wider first stage, the channel concat straight after a ConvBNReLU, down-sampling as the first encodeAB layer
    ((0, 0, 64, 2), (0, 1, 256, 2), (1, 1, 512, 2), (1, 2, 512, 1), (1, 2, 512, 1)),
], ids=["shuffled_blocks", "short_wide"])
def test_backbone_layer_list_is_data(tmp_path, arch):
    """SURVEY section 7, hard part (a): the backbone is recalled, not given, so its layer list is DATA - the "arch"
    table of the weight file, from which csrc/network.cu builds its launch plan (buffers, phase-split / concat output
    layouts, tile shapes).  Two architectures that differ from the default one, exported from the same-arch PyTorch
    modules, against those modules in fp32 (refiner and scorer, through the inference-core boundary)."""
    import ctypes as C
    from foundationpose_cpp_b200 import capi, weights
    lib = capi.lib()
    n = 3
    torch.manual_seed(9)
    a, b = torch.rand(n, 160, 160, 6), torch.rand(n, 160, 160, 6)
    a[..., 3:] = a[..., 3:] * 2 - 1
    b[..., 3:] = b[..., 3:] * 2 - 1
    ad, bd = a.cuda(), b.cuda()
    ins = (nets.nhwc_to_inputs(a.half().float()), nets.nhwc_to_inputs(b.half().float()))
    refiner = nets.seeded_init(nets.RefineNet(arch), 12)
    scorer = nets.seeded_init(nets.ScoreNet(arch), 13)
    rp, sp = str(tmp_path / "r.fpb2w"), str(tmp_path / "s.fpb2w")
    weights.export_refiner(rp, refiner.state_dict())
    weights.export_scorer(sp, scorer.state_dict())
    assert weights.read_weight_file(rp)["arch"].tolist() == [list(map(float, r)) for r in arch]
    core = C.c_void_p()
    capi.check(lib.fpb200_infer_create(rp.encode(), 0, 4, 0, C.byref(core)))
    trans, rot = torch.zeros(n, 3, device="cuda"), torch.zeros(n, 3, device="cuda")
    capi.check(lib.fpb200_infer_refiner(core, C.c_void_p(ad.data_ptr()), C.c_void_p(bd.data_ptr()), n,
                                        C.c_void_p(trans.data_ptr()), C.c_void_p(rot.data_ptr())))
    with torch.no_grad():
        t_ref, r_ref = refiner(*ins)
    _net_close(trans.cpu().numpy(), t_ref.numpy(), "variant arch trans")
    _net_close(rot.cpu().numpy(), r_ref.numpy(), "variant arch rot")
    lib.fpb200_infer_destroy(core)
    core = C.c_void_p()
    capi.check(lib.fpb200_infer_create(sp.encode(), 1, 4, 0, C.byref(core)))
    scores = torch.zeros(n, device="cuda")
    capi.check(lib.fpb200_infer_scorer(core, C.c_void_p(ad.data_ptr()), C.c_void_p(bd.data_ptr()), n, C.c_void_p(scores.data_ptr())))
    with torch.no_grad():
        s_ref = scorer(*ins)
    _net_close(scores.cpu().numpy(), s_ref.numpy(), "variant arch scores")
    lib.fpb200_infer_destroy(core)
    # a layer list the kernels cannot run is refused at load time with a message, not run wrongly
    bad = dict(weights.refiner_tensors(refiner.state_dict()))
    bad["arch"] = torch.tensor([[0.0, 0, 64, 2], [0, 1, 128, 2], [1, 1, 512, 2], [1, 1, 512, 2]])  # ends in 10 x 10 tokens
    bp = str(tmp_path / "bad.fpb2w")
    weights.write_weight_file(bp, weights.KIND_REFINER, bad)
    core = C.c_void_p()
    assert lib.fpb200_infer_create(bp.encode(), 0, 4, 0, C.byref(core)) != 0
    assert b"arch" in lib.fpb200_last_error() or b"tensor" in lib.fpb200_last_error()


def test_track_stream_equals_consecutive_track_calls(env):
    """fpb200_track_stream_* - frame k + 1 uploaded on a copy stream into the second buffer set
    while frame k is refined, the refined pose chained in device memory.  10 frames (noisy copies of the scene, so that
    consecutive uploads differ), two in flight: every pose must equal, bit for bit, consecutive Track calls that feed
    each result back in through the host; eager, captured and replayed calls of all four graph variants are covered.
    Then an ordinary Track on the same session, and the error paths."""
    m = env["model"]
    rng = np.random.default_rng(5)
    frames = []
    for k in range(10):
        rgb = np.clip(env["rgb"].astype(np.int16) + rng.integers(-3, 4, env["rgb"].shape), 0, 255).astype(np.uint8)
        depth = (env["depth"] + (env["depth"] > 0) * rng.normal(0, 2e-4, env["depth"].shape)).astype(F)
        frames.append((np.ascontiguousarray(rgb), np.ascontiguousarray(depth)))
    start = synth.perturb_pose(env["gt"], 3.0, 0.005, 7)
    want, pose = [], start
    for rgb, depth in frames:
        ok, pose = m.Track(rgb, depth, pose, "mustard_proxy", refine_itr=2)
        assert ok, m.last_error
        want.append(pose)
    assert any(not np.array_equal(want[i], want[i + 1]) for i in range(9))
    for rep in range(2):  # the second round replays the CUDA graphs captured by the first
        m.track_stream_begin(start, "mustard_proxy", 2)
        got = []
        m.track_stream_push(*frames[0])
        for k in range(1, 10):
            m.track_stream_push(*frames[k])  # two in flight
            got.append(m.track_stream_pop())
        with pytest.raises(api.capi.FpbError):
            m.track_stream_push(*frames[0]); m.track_stream_push(*frames[1])  # a third frame in flight is refused
        got.append(m.track_stream_pop())
        got_extra = m.track_stream_pop()  # the frame pushed inside the raises block above
        assert np.isfinite(got_extra).all()
        with pytest.raises(api.capi.FpbError):
            m.track_stream_pop()
        for k in range(10):
            assert np.array_equal(got[k], want[k]), (rep, k)
    ok, pose = m.Track(frames[0][0], frames[0][1], start, "mustard_proxy", refine_itr=2)
    assert ok and np.array_equal(pose, want[0])
